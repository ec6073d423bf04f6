"""Randomised differential run of the CUDA codec against the CPU oracle (test infrastructure, not collected by
pytest):  python tests/fuzz_codec.py [seconds] [seed]
Random geometries (tuned and generic kernels), dtypes and value distributions -- mixtures of noise, constant and
zero stretches of random lengths (around the 8-element lane chunk and the 255 cap), NaN / Inf / denormal
sprinkles, tiny and huge magnitudes -- compressed and decompressed on the GPU; sizes, scales, payload bytes,
decoded lengths and decoded bits must equal the oracle's."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cxl_speckv_b200 import codec
from oracle.oracle import Port
from tests.helpers import BF16, F16, bf16_from_f32


def make(rng, n):
    x = rng.standard_normal(n).astype(np.float32) * np.float32(np.exp(rng.uniform(-3, 3)))
    kind = rng.integers(0, 6)
    if kind >= 1:                                   # stretches: constant / zero / slowly varying
        pos = 0
        while pos < n:
            seg = int(rng.choice([1, 2, 3, 7, 8, 9, 15, 16, 17, 254, 255, 256, 300, 511, 2047, 2048, 2049, 5000]))
            seg = max(1, int(seg * rng.uniform(0.5, 1.5))) if rng.random() < 0.3 else seg
            mode = rng.integers(0, 4)
            if mode == 0:
                x[pos:pos + seg] = x[pos]
            elif mode == 1:
                x[pos:pos + seg] = 0.0
            elif mode == 2 and kind >= 3:
                x[pos:pos + seg] = np.linspace(x[pos], x[pos] * 1.5, min(seg, n - pos), dtype=np.float32)
            pos += seg + int(rng.integers(0, 40 if kind < 5 else 4000))
    if rng.random() < 0.3:
        idx = rng.integers(0, n, max(1, n // 3000))
        x[idx] = rng.choice(np.array([np.nan, np.inf, -np.inf, 1e-8, -1e-8, 6e-8, 65504.0, -0.0], dtype=np.float32), idx.size)
    return x


def main(seconds=120.0, seed=0):
    rng = np.random.default_rng(seed)
    t_end = time.time() + seconds
    cases = 0
    while time.time() < t_end:
        tuned = rng.random() < 0.75
        if tuned:
            G = 2048 * int(2 ** rng.integers(0, 8))
        else:
            G = int(rng.integers(1, 20000))
        n_groups = int(max(1, min(64, rng.integers(1, max(2, (3 << 20) // G)))))
        dtype = F16 if rng.random() < 0.6 else BF16
        x = np.concatenate([make(rng, G) for _ in range(n_groups)])
        if dtype == BF16 and rng.random() < 0.3:
            x *= np.float32(np.exp(rng.uniform(-70, 80)))
        if dtype == F16:
            with np.errstate(over="ignore"):
                raw = x.astype(np.float16)
            xd = torch.from_numpy(raw).cuda()
        else:
            raw = bf16_from_f32(x)
            xd = torch.from_numpy(raw.astype(np.int16)).cuda().view(torch.bfloat16)
        c = codec.compress(xd, G)
        oel = torch.zeros(n_groups, dtype=torch.int32, device="cuda")
        y = codec.decompress(c, out_elems=oel)
        torch.cuda.synchronize()
        payload, scales, comp = Port.compress_batch(raw, G, dtype=dtype, threads=8)
        tag = f"case {cases}: G={G} n={n_groups} dtype={dtype} seed={seed}"
        assert np.array_equal(c.scales.cpu().numpy().view(np.uint32), scales.view(np.uint32)), tag
        assert np.array_equal(c.comp_bytes.cpu().numpy().view(np.uint32), comp), tag
        gp = c.payload.cpu().numpy()
        for g in range(n_groups):
            assert np.array_equal(gp[g, :comp[g]], payload[g, :comp[g]]), (tag, g)
        want, want_n = Port.decompress_batch(payload, scales, comp, G, dtype, threads=8)
        assert np.array_equal(oel.cpu().numpy().view(np.uint32), want_n), tag
        got = y.view(torch.int16).cpu().numpy().view(np.uint16)
        wantb = want.view(np.uint16).reshape(n_groups, G)
        if not np.array_equal(got, wantb):
            bad = np.argwhere(got != wantb)
            g0, e0 = bad[0]
            print(tag, "mismatches:", len(bad), "first at group", g0, "elem", e0, "got", hex(got[g0, e0]), "want", hex(wantb[g0, e0]),
                  "scale bits", hex(int(scales.view(np.uint32)[g0])), "comp", comp[g0], "groups hit", sorted(set(bad[:, 0].tolist()))[:10],
                  "input around", raw.reshape(n_groups, G)[g0, max(0, e0 - 3):e0 + 3])
            raise AssertionError(tag)
        cases += 1
    print(f"fuzz ok: {cases} cases in {seconds:.0f} s (seed {seed}); {codec.stats()}")


if __name__ == "__main__":
    main(float(sys.argv[1]) if len(sys.argv) > 1 else 120.0, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
