"""Ad-hoc GPU check + timing of the run-expansion decode path against the library's generic kernel and the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from cxl_speckv_b200 import codec
from oracle.oracle import Port

dev = "cuda:0"
torch.manual_seed(0)
G = 131072
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for name, mk in (("runs300", lambda n: torch.randn((n * G + 299) // 300, device=dev).half().repeat_interleave(300)[: n * G].contiguous()),
                 ("smooth", lambda n: torch.cumsum(torch.randn(n, G, device=dev) * 0.01, 1).half().view(-1)),
                 ("mixed", lambda n: torch.where(torch.rand(n * G, device=dev) < 0.5, torch.randn(n * G, device=dev), torch.zeros(n * G, device=dev)).half()),
                 ("tail60", lambda n: torch.where((torch.arange(G, device=dev) < int(G * 0.4)).repeat(n), torch.randn(n * G, device=dev), torch.zeros(n * G, device=dev)).half()),
                 ("one700", lambda n: torch.randn(n * G, device=dev).half().index_fill_(0, torch.arange(G // 2, G // 2 + 700, device=dev), 0.5))):
    for n in (4, 512):
        x = mk(n)
        for scheme in (2, 3):
            c = codec.compress(x, G, scheme=scheme)
            oel = torch.zeros(n, dtype=torch.int32, device=dev)
            y = codec.decompress(c, out_elems=oel)
            torch.cuda.synchronize()
            k = min(n, 6)
            raw = x[: k * G].cpu().numpy()
            p, s, cb = Port.compress_batch(raw, G, threads=8, scheme=scheme)
            want, wn = Port.decompress_batch(p, s, cb, G, 0, threads=8, scheme=scheme)
            ok = np.array_equal(y[:k].view(torch.int16).cpu().numpy().view(np.uint16), want.view(np.uint16)) and (oel == G).all().item()
            # the compressed side: sizes, scales and payload bytes against the oracle
            gcb = c.comp_bytes[:k].cpu().numpy()
            okc = np.array_equal(gcb, cb) and np.array_equal(c.scales[:k].cpu().numpy().view(np.uint32), np.asarray(s).view(np.uint32))
            gp = c.payload[:k].cpu().numpy()
            pp = np.asarray(p).reshape(k, -1)
            for i in range(k):
                okc = okc and np.array_equal(gp[i, : int(cb[i])], pp[i, : int(cb[i])])
            ok = ok and okc
            td = t(lambda: codec.decompress(c, out=y))
            tc = t(lambda: codec.compress(x, G, scheme=scheme, out=c))
            cbs = float(c.comp_bytes.to(torch.int64).sum())
            print(f"{name:8s} n={n:4d} scheme {scheme}: {'ok ' if ok else 'BAD'} ratio {n*G*2/cbs:7.2f} decompress {td*1e3:8.1f} us "
                  f"{n*G*2/td/1e6:7.0f} KV GB/s  alg {(n*G*2+cbs)/td/1e6/6553:5.3f} of peak | compress {tc*1e3:8.1f} us "
                  f"alg {(n*G*2+cbs)/tc/1e6/6553:5.3f} of peak", flush=True)
print(codec.stats())
