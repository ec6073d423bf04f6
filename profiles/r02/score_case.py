"""Scoring calls for ncu / timing: python profiles/r02/score_case.py [batch] [k]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from cxl_speckv_b200 import prefetch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rng = np.random.default_rng(1)
emb = ((rng.random((32000, 64), dtype=np.float32) - 0.5) * 0.1).astype(np.float32)
wout = ((rng.random((32000, 128), dtype=np.float32) - 0.5) * 0.1).astype(np.float32)
prefetch.load_predictor(emb, wout)
toks = torch.from_numpy(np.random.default_rng(7).integers(0, 32000, (B, 16)).astype(np.int32)).cuda()
table = torch.zeros((prefetch.table_records(B, k), 32), dtype=torch.uint8, device="cuda")
for _ in range(5):
    prefetch.emit(toks, k=k, table=table)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    prefetch.emit(toks, k=k, table=table)
b.record(); torch.cuda.synchronize()
print(f"emit B={B} k={k}: {a.elapsed_time(b) / 50 * 1e3:.1f} us per call")
