"""GPU: the other compression schemes through the C ABI.

Scheme 1 (INT8, host/include/speckv.h:61) is the reference's quantiser alone (cache_engine.cpp:186-196, :275-284) and
is checked against the oracle restatement of exactly those lines.  Schemes 3 / 4 are this library's explicit extension
ids (the non-wrapping quantiser SURVEY.md section 8f-4 asks for) -- NOT reference behaviour; their oracle is the
restatement in oracle/speckv_oracle.c marked as such.  Bit-exact either way: scales, sizes, payload bytes, decoded bits.
Also: speckv_set_compression_scheme as a real switch (host/include/speckv.h:59-66) over the tier and the frozen API,
pools of bf16 elements, and the host-buffer API shipping payload bytes rather than slots."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle.oracle import Port
from tests.helpers import BF16, F16, bf16_from_f32, f32_bits

pytestmark = pytest.mark.gpu

import cxl_speckv_b200 as pkg  # noqa: E402
from cxl_speckv_b200 import (COMP_FP16, COMP_INT8, COMP_INT8_CLAMP, COMP_INT8_CLAMP_DELTA_RLE, COMP_INT8_DELTA_RLE,  # noqa: E402
                             codec)
from cxl_speckv_b200.tier import HostTier  # noqa: E402

DEV = "cuda:0"


def _inputs(rng, n_groups, G, bf):
    x = rng.standard_normal(n_groups * G).astype(np.float32).reshape(n_groups, G)
    x *= np.exp(rng.uniform(-8, 8, n_groups)).astype(np.float32)[:, None]
    x[1 % n_groups] = 0.0                                                 # zero group: scale 1
    x[2 % n_groups, : G // 2] = 0.375                                     # constant stretch
    if n_groups > 4:
        x[3, 5] = np.inf                                                  # non-finite group max
        x[4, 7] = np.nan                                                  # NaN is skipped by the max and codes as 0
    if n_groups > 6 and bf:
        x[5] *= 1e-25                                                     # bf16 max below 2^-60: exact division path
    x = x.reshape(-1)
    if bf:
        bits = bf16_from_f32(x)
        return bits, torch.from_numpy(bits.astype(np.int16)).to(DEV).view(torch.bfloat16), BF16
    with np.errstate(over="ignore"):
        h = x.astype(np.float16)
    return h, torch.from_numpy(h).to(DEV), F16


@pytest.mark.parametrize("scheme", [COMP_INT8, COMP_INT8_CLAMP_DELTA_RLE, COMP_INT8_CLAMP])
@pytest.mark.parametrize("G,n_groups", [(131072, 6), (2048, 70), (32768, 9), (1000, 33), (65536, 5)])
@pytest.mark.parametrize("bf", [False, True])
def test_scheme_bit_exact_vs_oracle(scheme, G, n_groups, bf):
    rng = np.random.default_rng(G + scheme)
    raw, xd, dcode = _inputs(rng, n_groups, G, bf)
    c = codec.compress(xd, G, scheme=scheme)
    y = codec.decompress(c)
    torch.cuda.synchronize()
    slot = c.payload.shape[1]
    payload, scales, comp = Port.compress_batch(raw, G, dtype=dcode, threads=8, scheme=scheme, slot_bytes=max(slot, 2 * G))
    assert np.array_equal(c.comp_bytes.cpu().numpy().view(np.uint32), comp)
    assert np.array_equal(f32_bits(c.scales.cpu().numpy()), f32_bits(scales))
    gp = c.payload.cpu().numpy()
    for g in range(n_groups):
        assert np.array_equal(gp[g, :comp[g]], payload[g, :comp[g]]), g
    want, _ = Port.decompress_batch(payload, scales, comp, G, dcode, threads=8, scheme=scheme)
    assert np.array_equal(y.view(torch.int16).cpu().numpy().view(np.uint16), want.view(np.uint16))
    if scheme == COMP_INT8_CLAMP_DELTA_RLE:   # same container, same decoder as the reference scheme
        c2 = codec.CompressedKV(c.payload, c.scales, c.comp_bytes, G, c.dtype, COMP_INT8_DELTA_RLE)
        assert torch.equal(codec.decompress(c2).view(torch.int16), y.view(torch.int16))


def test_clamped_scheme_reconstructs_where_the_reference_scheme_wraps():
    """What the extension is for: on N(0,1) KV the reference's double x127 leaves an error of the order of the data
    (MSE ~ 1); the clamped quantiser's error is the usual step^2 / 12."""
    torch.manual_seed(3)
    G, n = 131072, 16
    x = torch.randn(n * G, device=DEV).half()
    mse = {}
    for sch in (COMP_INT8_DELTA_RLE, COMP_INT8_CLAMP_DELTA_RLE, COMP_INT8, COMP_INT8_CLAMP):
        y = codec.decompress(codec.compress(x, G, scheme=sch))
        mse[sch] = ((y.float() - x.view(n, G).float()) ** 2).mean().item()
    assert 0.9 < mse[COMP_INT8_DELTA_RLE] < 1.1 and 0.9 < mse[COMP_INT8] < 1.1
    step = (x.view(n, G).float().abs().amax(1) / 127).pow(2).mean().item() / 12
    assert mse[COMP_INT8_CLAMP_DELTA_RLE] < 1.5 * step and mse[COMP_INT8_CLAMP] == mse[COMP_INT8_CLAMP_DELTA_RLE]


@pytest.mark.parametrize("G", [131072, 2048, 1000])
def test_passthrough_scheme_roundtrip(G):
    x = torch.randn(40 * G, device=DEV).half()
    c = codec.compress(x, G, scheme=COMP_FP16)
    assert (c.comp_bytes == 2 * G).all() and (c.scales == 1.0).all()
    assert torch.equal(codec.decompress(c).view(torch.int16).view(-1), x.view(torch.int16))
    idx = torch.tensor([3, 0, 39, 3], dtype=torch.int32, device=DEV)
    assert torch.equal(codec.decompress_indexed(c, idx).view(torch.int16), x.view(40, G)[idx.long()].view(torch.int16))


def test_tier_stores_blocks_under_the_scheme_in_force():
    """speckv_ext_tier_set_scheme: new offloads use the current scheme, stored blocks keep theirs; one restore call
    may name blocks of several schemes (it is cut into runs)."""
    G, n = 2048, 600
    torch.manual_seed(9)
    x = torch.randn(n * G, device=DEV).half()
    ids = (np.arange(n, dtype=np.uint64) + 1) << np.uint64(12)
    tier = HostTier(64 << 20)
    try:
        parts = [(0, 200, COMP_INT8_DELTA_RLE), (200, 350, COMP_INT8_CLAMP), (350, 500, COMP_INT8_CLAMP_DELTA_RLE),
                 (500, 560, COMP_INT8), (560, 600, COMP_FP16)]
        stored = 0
        for a, b, sch in parts:
            tier.set_scheme(sch)
            before = tier.stats()["used_bytes"]
            tier.offload(x[a * G:b * G], G, ids[a:b])
            used = tier.stats()["used_bytes"] - before
            if sch in (COMP_INT8, COMP_INT8_CLAMP):
                assert used == (b - a) * G                          # one byte per element
            if sch == COMP_FP16:
                assert used == (b - a) * G * 2
            stored += used
        tier.set_scheme(COMP_INT8_DELTA_RLE)
        order = np.random.default_rng(1).permutation(n)
        out = tier.restore(ids[order], G, torch.float16, device=DEV)
        want = torch.empty((n, G), dtype=torch.float16, device=DEV)
        for a, b, sch in parts:
            want[a:b] = codec.decompress(codec.compress(x[a * G:b * G], G, scheme=sch))
        assert torch.equal(out.view(torch.int16), want[torch.from_numpy(order).to(DEV)].view(torch.int16))
    finally:
        tier.close()


def test_set_compression_scheme_drives_the_frozen_api_and_bf16_pools():
    """speckv_set_compression_scheme (host/include/speckv.h:59-66) / SET_PARAM key 2 select the scheme pages are
    offloaded under; a bf16 pool is quantised as bf16 (speckv_ext_set_pool_dtype), not as reinterpreted fp16."""
    from cxl_speckv_b200 import CxlSpeckvKVAllocator
    alloc = CxlSpeckvKVAllocator(pkg.lib_path(), "cuda:0")
    L = pkg.lib()
    tier = HostTier(32 << 20)
    try:
        tokens, layers, heads, hd = 128, 2, 8, 128
        alloc.allocate(tokens, layers, heads, hd, 2)
        n_pages = tokens * layers * heads * hd * 2 * 2 // 4096
        torch.manual_seed(11)
        pool = (torch.randn(n_pages * 2048, device=DEV) * 3).to(torch.bfloat16)
        orig = pool.clone()
        alloc.bind_pool(pool, tier)
        half = n_pages // 2
        assert L.speckv_set_compression_scheme(COMP_INT8_CLAMP) == 0
        sch = C.c_int(-1)
        assert L.speckv_ext_get_compression_scheme(C.byref(sch)) == 0 and sch.value == COMP_INT8_CLAMP
        alloc.offload_pages(0, half)
        assert tier.stats()["used_bytes"] == half * 2048                  # codes only: the switch reached the tier
        assert L.speckv_ext_set_param(2, COMP_INT8_DELTA_RLE) == 0
        alloc.offload_pages(half, n_pages - half)
        assert L.speckv_ext_set_param(2, 9) != 0 and L.speckv_ext_set_param(7, 1) == -4
        pool.zero_()
        alloc.fetch_pages(0, n_pages)
        want = torch.empty_like(orig).view(n_pages, 2048)
        want[:half] = codec.decompress(codec.compress(orig[: half * 2048], 2048, scheme=COMP_INT8_CLAMP))
        want[half:] = codec.decompress(codec.compress(orig[half * 2048:], 2048, scheme=COMP_INT8_DELTA_RLE))
        assert torch.equal(pool.view(torch.int16), want.view(-1).view(torch.int16))
        # the clamped half is a faithful reconstruction of the bf16 values (a pool read as fp16 would not be)
        err = (pool[: half * 2048].float() - orig[: half * 2048].float()).abs().max().item()
        assert err < 0.1 * orig.float().abs().max().item()
    finally:
        L.speckv_finalize()
        tier.close()


def test_host_api_ships_payload_bytes_not_slots():
    """speckv_ext_compress_host / _decompress_host: on compressible input the PCIe traffic follows comp_bytes."""
    L = pkg.lib()
    G, n = 131072, 64
    x = np.zeros(n * G, dtype=np.float16)
    x.reshape(n, G)[:, ::1000] = 1.5                                   # long constant stretches: tiny payloads
    x.reshape(n, G)[5] = np.random.default_rng(0).standard_normal(G).astype(np.float16)   # one incompressible group
    sb = codec.slot_bytes(G)
    payload = np.zeros(n * sb, dtype=np.uint8)
    scales, comp = np.zeros(n, np.float32), np.zeros(n, np.uint32)
    s0 = codec.stats()
    assert L.speckv_ext_compress_host(x.ctypes.data, 0, G, n, payload.ctypes.data, sb, scales.ctypes.data, comp.ctypes.data, 2) == 0
    s1 = codec.stats()
    want_p, want_s, want_c = Port.compress_batch(x, G, threads=8)
    assert np.array_equal(comp, want_c) and np.array_equal(f32_bits(scales), f32_bits(want_s))
    for g in range(n):
        assert np.array_equal(payload.reshape(n, sb)[g, :comp[g]], want_p[g, :comp[g]]), g
    moved = s1["host_api_d2h_bytes"] - s0["host_api_d2h_bytes"]
    assert moved <= int(((comp.astype(np.int64) + 15) & ~15).sum()) + 8 * n < n * sb // 20
    out = np.zeros(n * G, dtype=np.float16)
    assert L.speckv_ext_decompress_host(payload.ctypes.data, sb, scales.ctypes.data, comp.ctypes.data, G, n, 0, out.ctypes.data, None, 2) == 0
    s2 = codec.stats()
    want_y, _ = Port.decompress_batch(want_p, want_s, want_c, G, F16, threads=8)
    assert np.array_equal(out.view(np.uint16), want_y.view(np.uint16).ravel())
    assert s2["host_api_h2d_bytes"] - s1["host_api_h2d_bytes"] == moved
