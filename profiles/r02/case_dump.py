"""Replays one case of tests/fuzz_codec.py (seed, case index) and prints where the compressed payload differs from the oracle:
python profiles/r02/case_dump.py SEED CASE"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tests.fuzz_codec import make
from tests.helpers import BF16, F16, bf16_from_f32
from cxl_speckv_b200 import codec
from oracle.oracle import Port
seed, target = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(seed)
for case in range(target + 1):
    tuned = rng.random() < 0.75
    G = 2048 * int(2 ** rng.integers(0, 8)) if tuned else int(rng.integers(1, 20000))
    n_groups = int(max(1, min(64, rng.integers(1, max(2, (3 << 20) // G)))))
    dtype = F16 if rng.random() < 0.6 else BF16
    x = np.concatenate([make(rng, G) for _ in range(n_groups)])
    if dtype == BF16 and rng.random() < 0.3:
        with np.errstate(over="ignore"):
            x *= np.float32(np.exp(rng.uniform(-70, 80)))
if dtype == F16:
    with np.errstate(over="ignore"):
        raw = x.astype(np.float16)
    xd = torch.from_numpy(raw).cuda()
else:
    raw = bf16_from_f32(x)
    xd = torch.from_numpy(raw.astype(np.int16)).cuda().view(torch.bfloat16)
print(f"case {target}: G={G} n={n_groups} dtype={dtype}")
payload, scales, comp = Port.compress_batch(raw, G, dtype=dtype, threads=8)
for rep in range(3):
    c = codec.compress(xd, G)
    torch.cuda.synchronize()
    gp = c.payload.cpu().numpy()
    gc = c.comp_bytes.cpu().numpy().view(np.uint32)
    for g in range(n_groups):
        if gc[g] != comp[g]:
            print(f"rep {rep} group {g}: comp {gc[g]} want {comp[g]}")
            continue
        a, b = gp[g, :comp[g]].reshape(-1, 2), payload[g, :comp[g]].reshape(-1, 2)
        bad = np.nonzero((a != b).any(axis=1))[0]
        if len(bad):
            pos = np.concatenate([[0], np.cumsum(b[:, 1].astype(np.int64))])
            print(f"rep {rep} group {g}: {len(bad)} bad pairs of {len(a)}; first {bad[:12].tolist()}")
            for i in bad[:6]:
                print(f"   pair {i}: got {a[i].tolist()} want {b[i].tolist()}  starts at element {pos[i]} (region {pos[i] // 2048}, rel {pos[i] % 2048}, iter {pos[i] % 2048 // 256}, lane {pos[i] % 256 // 8})"
                      f"  neighbours got {a[max(0,i-2):i+3].tolist()} want {b[max(0,i-2):i+3].tolist()}")
try:
    import ctypes
    from cxl_speckv_b200._lib import lib
    buf = (ctypes.c_uint32 * 1024)()
    if lib().speckv_dbg_read(buf) == 0:
        for b in (18, 19, 20):
            print("dbg block", b, [(buf[b * 16 + w] >> 16, buf[b * 16 + w] & 0xffff) for w in range(16)])
        print("loop trace (st.z, n|0x100, inc|0x10000, pidx, am@31, am@17, -, -) per k:")
        for k in range(8):
            print("  k", k, [hex(buf[512 + k * 8 + j]) for j in range(8)])
except AttributeError:
    pass
print("done")
