"""Times the codec kernels on one geometry (CUDA events, 2 GiB of fp16 per call by default) and
checks the result against the library's own first output (regression guard between variants):
python profiles/quick_time.py [G] [n_groups] [iters]     (SPECKV_LIB selects a variant build)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cxl_speckv_b200 import codec

G = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 30
torch.manual_seed(1234)
x = torch.randn(n * G, device="cuda").half()
c = codec.compress(x, G)
y = codec.decompress(c)
torch.cuda.synchronize()
cb = c.comp_bytes.to(torch.int64)
digest = (int(cb.sum().item()), int(c.payload[:, :4096].to(torch.int64).sum().item()),
          int(y.view(torch.int16).to(torch.int64).sum().item()))
res = {}
for name, fn in (("compress", lambda: codec.compress(x, G, out=c)), ("decompress", lambda: codec.decompress(c, out=y))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    tot = 0.0
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b)
        best = min(best, t)
        tot += t
    res[name] = (tot / iters, best)
alg = n * (2 * G + 12) + int(cb.sum().item())
tag = os.path.basename(os.environ.get("SPECKV_LIB", "default"))
print(f"{tag:28s} G={G} n={n} compress {res['compress'][0]*1e3:7.1f} us (best {res['compress'][1]*1e3:7.1f}) "
      f"{alg/res['compress'][0]/1e6:6.0f} GB/s | decompress {res['decompress'][0]*1e3:7.1f} us (best {res['decompress'][1]*1e3:7.1f}) "
      f"{alg/res['decompress'][0]/1e6:6.0f} GB/s | digest {digest}")
