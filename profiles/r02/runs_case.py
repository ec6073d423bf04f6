"""One decompress call on a runs-of-300 batch (for ncu): python profiles/r02/runs_case.py [runs300|smooth] [n_groups] [scheme]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cxl_speckv_b200 import codec
kind = sys.argv[1] if len(sys.argv) > 1 else "runs300"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
scheme = int(sys.argv[3]) if len(sys.argv) > 3 else 2
G = 131072
torch.manual_seed(0)
if kind == "runs300":
    x = torch.randn((n * G + 299) // 300, device="cuda").half().repeat_interleave(300)[: n * G].contiguous()
else:
    x = torch.cumsum(torch.randn(n, G, device="cuda") * 0.01, 1).half().view(-1)
c = codec.compress(x, G, scheme=scheme)
y = codec.decompress(c)
for _ in range(3):
    codec.compress(x, G, scheme=scheme, out=c)
    codec.decompress(c, out=y)
torch.cuda.synchronize()
