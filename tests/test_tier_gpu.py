"""GPU: offload to / restore from the pinned host tier (BASELINE config 4 geometry: 4 KiB page
groups).  The restored pages must equal a direct compress -> decompress on the device bit for
bit, which in turn is bit-exact against the oracle (tests/test_codec_gpu.py)."""
import numpy as np
import pytest
import torch

from oracle.oracle import Port

pytestmark = pytest.mark.gpu

from cxl_speckv_b200 import codec  # noqa: E402
from cxl_speckv_b200.tier import HostTier  # noqa: E402
from cxl_speckv_b200 import SpeckvError  # noqa: E402

DEV = "cuda:0"


@pytest.mark.parametrize("G,n_groups", [(2048, 50000), (131072, 1200), (1000, 333)])
def test_offload_restore_roundtrip(G, n_groups):
    torch.manual_seed(G)
    x = torch.randn(n_groups * G, device=DEV).half()
    x[3 * G:5 * G] = 0                                    # highly compressible pages (tiny payload)
    x[7 * G:8 * G] = 1.5
    tier = HostTier(pool_bytes=int(n_groups * G * 2 * 1.6) + (1 << 20))   # chunks need contiguous extents
    try:
        ids = (np.arange(n_groups, dtype=np.uint64) << np.uint64(12)) | (np.uint64(1) << np.uint64(32))   # virt_page_id style
        tier.offload(x, G, ids)
        st = tier.stats()
        assert st["blocks"] == n_groups and st["bytes_offloaded_raw"] == n_groups * G * 2
        c = codec.compress(x, G)
        comp = c.comp_bytes.cpu().numpy().view(np.uint32).astype(np.uint64)
        assert st["used_bytes"] == int(((comp + 15) // 16 * 16).sum()) == st["bytes_offloaded_stored"]
        want = codec.decompress(c)
        y = tier.restore(ids, G, torch.float16)
        assert torch.equal(y.view(torch.int16), want.view(torch.int16))
        # a scattered subset, in another order (prefetch / page-fault pattern)
        rng = np.random.default_rng(0)
        sel = rng.permutation(n_groups)[:257]
        y2 = tier.restore(ids[sel], G, torch.float16)
        assert torch.equal(y2.view(torch.int16), want[torch.from_numpy(sel).to(DEV)].view(torch.int16))
        # oracle spot check on the restored bytes
        for g in (0, 3, 7, n_groups - 1):
            s, p = Port.compress(x[g * G:(g + 1) * G].cpu().numpy().astype(np.float32))
            ref = Port.decompress(s, p, G).astype(np.float16)
            assert np.array_equal(y[g].cpu().numpy().view(np.uint16), ref.view(np.uint16))
        # drop half, re-offload new content under the same ids: space is reused
        tier.drop(ids[: n_groups // 2])
        assert tier.stats()["blocks"] == n_groups - n_groups // 2
        with pytest.raises(SpeckvError):
            tier.restore(ids[:4], G, torch.float16)       # unknown after drop -> SPECKV_ERR_GENERAL
        x2 = torch.randn((n_groups // 2) * G, device=DEV).half()
        tier.offload(x2, G, ids[: n_groups // 2])
        y3 = tier.restore(ids[: n_groups // 2], G, torch.float16)
        assert torch.equal(y3.view(torch.int16), codec.decompress(codec.compress(x2, G)).view(torch.int16))
        assert tier.stats()["used_bytes"] <= tier.stats()["pool_bytes"]
    finally:
        tier.close()


@pytest.mark.parametrize("tdt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("scheme", [2, 3])
def test_offload_page_groups_packed_emission(tdt, scheme):
    """4 KiB page groups: the compress kernel places the payloads in the packed stream itself (look-back over per-CTA
    totals, no pack pass).  Pages of every kind the kernel sizes differently -- noise, zero pages (closed form), constant
    and long-run pages (long path), pages with +-inf / tiny magnitudes (left to the generic kernel: they keep a whole
    slot inside the stream) -- over enough pages for look-back chains of hundreds of CTAs; every restored page must be
    what the device codec gives, and the stored bytes must be the 16 B-rounded payloads plus those slots' slack."""
    G, n_groups = 2048, 40000 + 7
    g = torch.Generator(device=DEV)
    g.manual_seed(99)
    x = torch.randn(n_groups, G, device=DEV, generator=g)
    kind = torch.randint(0, 10, (n_groups,), device=DEV, generator=g)
    x[kind == 1] = 0.0
    x[kind == 2] = 0.75
    runs = (kind == 3).nonzero().flatten()
    x[runs] = x[runs][:, ::300].repeat_interleave(300, dim=1)[:, :G]
    half = (kind == 4).nonzero().flatten()
    x[half, : G // 2 + 5] = 0.0
    x = x.to(tdt)
    special = (kind == 5).nonzero().flatten()
    x[special[0::2], 17] = float("inf")
    if tdt == torch.bfloat16:
        x[special[1::2]] = (x[special[1::2]].float() * 1e-38).to(tdt)      # scales outside the short quantiser's domain
    x = x.contiguous().view(-1)
    tier = HostTier(pool_bytes=int(n_groups * G * 2 * 1.3) + (1 << 20))
    try:
        ids = np.arange(n_groups, dtype=np.uint64) + np.uint64(1000)
        tier.set_scheme(scheme)
        tier.offload(x, G, ids)
        c = codec.compress(x, G, scheme=scheme)
        want = codec.decompress(c)
        y = tier.restore(ids, G, tdt)
        assert torch.equal(y.view(torch.int16), want.view(torch.int16))
        comp = c.comp_bytes.cpu().numpy().view(np.uint32).astype(np.int64)
        rounded = int(((comp + 15) // 16 * 16).sum())
        st = tier.stats()
        assert rounded <= st["bytes_offloaded_stored"] <= rounded + int(special.numel()) * 4096
        sel = np.random.default_rng(3).permutation(n_groups)[:513]
        y2 = tier.restore(ids[sel], G, tdt)
        assert torch.equal(y2.view(torch.int16), want[torch.from_numpy(sel).to(DEV)].view(torch.int16))
    finally:
        tier.close()


def test_pool_exhaustion_is_reported():
    G, n_groups = 2048, 4096
    x = torch.randn(n_groups * G, device=DEV).half()
    tier = HostTier(pool_bytes=1 << 20)                   # far too small
    try:
        with pytest.raises(SpeckvError) as ei:
            tier.offload(x, G, np.arange(n_groups, dtype=np.uint64))
        assert ei.value.status == -3                       # SPECKV_ERR_NOMEM
    finally:
        tier.close()


def test_tier_destroyed_before_the_pool_handle():
    """A tier closed while pools are still bound to it must leave no dangling pointer behind: access to a page whose
    only copy went with the tier reports an error, resident pages keep working, speckv_free does not touch the
    freed tier (advisor finding, c_api.cu PoolBinding::tier)."""
    import ctypes as C

    import cxl_speckv_b200 as pkg
    from cxl_speckv_b200 import CxlSpeckvKVAllocator

    alloc = CxlSpeckvKVAllocator(pkg.lib_path(), "cuda:0")
    L = pkg.lib()
    tier = HostTier(8 << 20)
    try:
        tokens, layers, heads, hd = 64, 1, 8, 128
        h = alloc.allocate(tokens, layers, heads, hd, 2)
        total = tokens * layers * heads * hd * 2 * 2
        n_pages = total // 4096
        pool = torch.randn(total // 2, device=DEV).half()
        alloc.bind_pool(pool, tier)
        alloc.offload_pages(0, n_pages // 2)             # the first half lives only in the tier now
        tier.close()                                     # ... and the tier goes away first
        ptr = C.c_void_p()
        assert L.speckv_access(h, (n_pages - 1) * 4096, 16, C.byref(ptr)) == 0      # resident page: served
        assert ptr.value == pool.data_ptr() + (n_pages - 1) * 4096
        assert L.speckv_access(h, 0, 16, C.byref(ptr)) != 0                          # its copy is gone: an error, not a crash
        alloc.prefetch_step(req_id=0, layer=0, cur_pos=1, recent_tokens=list(range(8)), depth_k=2)   # status discarded
        assert L.speckv_free(h) == 0
    finally:
        alloc._speckv.finalize()
        tier.close()


def test_frozen_api_serves_real_pointers_with_page_faults():
    """speckv_alloc/access/prefetch over a bound HBM pool + host tier: get_kv_ptr returns addresses
    inside the pool, demoted pages come back (decompressed) on access and on prefetch_step."""
    import ctypes as C

    import cxl_speckv_b200 as pkg
    from cxl_speckv_b200 import CxlSpeckvKVAllocator

    alloc = CxlSpeckvKVAllocator(pkg.lib_path(), "cuda:0")
    L = pkg.lib()
    tier = HostTier(16 << 20)
    try:
        tokens, layers, heads, hd = 256, 2, 8, 128
        h = alloc.allocate(tokens, layers, heads, hd, 2)
        total = tokens * layers * heads * hd * 2 * 2                   # bytes (K+V)
        n_pages = total // 4096
        torch.manual_seed(5)
        pool = torch.randn(total // 2, device=DEV).half()
        want = codec.decompress(codec.compress(pool, 2048)).clone()    # what a page looks like after a round trip
        orig = pool.clone()
        alloc.bind_pool(pool, tier)
        entry = hd * 2
        off = alloc._calc_offset(0, 1, 3, 7, 1, entry)
        assert alloc.get_kv_ptr(0, 1, 3, 7, 1, entry) == pool.data_ptr() + off      # a real device pointer
        assert torch.equal(pool, orig)                                               # resident pages are untouched

        alloc.offload_pages(0, n_pages)                                              # demote everything
        assert tier.stats()["blocks"] == n_pages
        pool.zero_()                                                                 # the HBM copy is gone
        p = alloc.get_kv_ptr(0, 1, 3, 7, 1, entry)                                   # page fault -> restore
        page = off // 4096
        assert p == pool.data_ptr() + off
        got = pool.view(n_pages, 2048)
        assert torch.equal(got[page].view(torch.int16), want[page].view(torch.int16))
        others = torch.ones(n_pages, dtype=torch.bool, device=DEV)
        others[page] = False
        assert (got[others] == 0).all()                                              # only that page came back

        # prefetch_step makes positions cur+1 .. cur+k of (req, layer) resident for K and V
        alloc.prefetch_step(req_id=0, layer=0, cur_pos=10, recent_tokens=list(range(16)), depth_k=4)
        for kind in (0, 1):
            o0 = alloc._calc_offset(0, 0, 0, 11, kind, entry)
            o1 = alloc._calc_offset(0, 0, 0, 15, kind, entry)
            for pg in range(o0 // 4096, (o1 + 4095) // 4096):
                assert torch.equal(got[pg].view(torch.int16), want[pg].view(torch.int16)), (kind, pg)
        # page-table flags: demoted = compressed only (4); restored = L2 | compressed (6)
        tbl = torch.zeros(n_pages * 3, dtype=torch.int64, device=DEV)
        cnt = C.c_size_t()
        assert L.speckv_ext_page_table_export(h, tbl.data_ptr(), n_pages, C.byref(cnt), None) == 0
        flags = (tbl.cpu().numpy().view(np.uint64).reshape(n_pages, 3)[:, 2] >> np.uint64(32)).astype(np.int64)
        assert flags[page] == 6 and set(np.unique(flags).tolist()) == {4, 6}
        # explicit promote of a range, then a span crossing pages
        assert L.speckv_ext_fetch_pages(h, 100, 8, None) == 0
        assert torch.equal(got[100:108].view(torch.int16), want[100:108].view(torch.int16))
        ptr = C.c_void_p()
        assert L.speckv_access(h, 200 * 4096 + 4000, 5000, C.byref(ptr)) == 0        # touches pages 200..202
        assert torch.equal(got[200:203].view(torch.int16), want[200:203].view(torch.int16))
        assert L.speckv_access(h, total, 1, C.byref(ptr)) == -1                       # past the end: as the reference
        assert L.speckv_free(h) == 0
        assert tier.stats()["blocks"] == 0                                            # the handle's blocks left the tier
    finally:
        alloc._speckv.finalize()
        tier.close()


def test_descriptor_interface_like_the_reference_ioctl_client():
    """speckv_ioctl_dma_desc batches (tests/test_dma.c), SET_PARAM with an invalid key
    (tests/test_params.c:68-84) and POLL_DONE against the CUDA backend."""
    import ctypes as C

    import cxl_speckv_b200 as pkg
    L = pkg.lib()

    class Desc(C.Structure):
        _fields_ = [("fpga_addr", C.c_uint64), ("gpu_addr", C.c_uint64), ("bytes", C.c_uint32), ("flags", C.c_uint32)]

    assert C.sizeof(Desc) == 24
    tier = HostTier(8 << 20)
    try:
        src = torch.randn(6 * 2048, device=DEV).half()
        dst = torch.zeros_like(src)
        base = src.data_ptr()
        # the shape of tests/test_dma.c: two single pages, one 2-page write, one compressed page
        wr = (Desc * 4)(Desc(0x4000000000, base, 4096, 1), Desc(0x4000001000, base + 4096, 4096, 1),
                        Desc(0x4000002000, base + 8192, 8192, 1), Desc(0x4000004000, base + 16384, 4096, 1 | 2))
        assert L.speckv_ext_poll_complete() == 0
        assert L.speckv_ext_submit_dma_batch(tier._h, wr, 4, None) == 0
        assert L.speckv_ext_poll_complete() == 4 and L.speckv_ext_poll_complete() == 0
        assert tier.stats()["blocks"] == 5
        d0 = dst.data_ptr()
        rd = (Desc * 4)(Desc(0x4000000000, d0, 4096, 0), Desc(0x4000001000, d0 + 4096, 4096, 0),
                        Desc(0x4000002000, d0 + 8192, 8192, 0), Desc(0x4000004000, d0 + 16384, 4096, 2))
        assert L.speckv_ext_submit_dma_batch(tier._h, rd, 4, None) == 0
        torch.cuda.synchronize()
        assert torch.equal(dst[:4 * 2048], src[:4 * 2048])                       # raw pages: bit copies
        want = codec.decompress(codec.compress(src[4 * 2048:5 * 2048], 2048)).view(-1)
        assert torch.equal(dst[4 * 2048:5 * 2048].view(torch.int16), want.view(torch.int16))   # codec page
        # reading a raw page through the codec (or an unknown page) fails like a bad descriptor
        bad = (Desc * 1)(Desc(0x4000000000, d0, 4096, 2))
        assert L.speckv_ext_submit_dma_batch(tier._h, bad, 1, None) == -1
        bad = (Desc * 1)(Desc(0x4999999000, d0, 4096, 0))
        assert L.speckv_ext_submit_dma_batch(tier._h, bad, 1, None) == -1
        assert L.speckv_ext_submit_dma_batch(tier._h, wr, 4097, None) == -4      # batch too large: -EINVAL
        # parameters
        assert L.speckv_ext_set_param(1, 6) == -4                                 # before speckv_init
        assert L.speckv_init(b"cuda:0") == 0
        try:
            assert L.speckv_ext_set_param(1, 6) == 0 and L.speckv_ext_set_param(2, 1) == 0
            d = C.c_uint32()
            assert L.speckv_ext_get_prefetch_depth(C.byref(d)) == 0 and d.value == 6
            assert L.speckv_ext_set_param(999, 123) == -4                         # invalid key is rejected
        finally:
            L.speckv_finalize()
    finally:
        tier.close()


@pytest.mark.parametrize("G,num_blocks,n_sel", [(2048, 70000, 40000), (32768, 900, 555), (1000, 400, 123)])
def test_paged_offload_restore_through_block_tables(G, num_blocks, n_sel):
    """speckv_ext_tier_offload_paged / _restore_paged: blocks of a paged KV cache leave and come back through
    block tables (vLLM layout), several chunks per call; the result equals the staged path bit for bit and
    cache blocks that were not named stay untouched."""
    torch.manual_seed(G + 1)
    rng = np.random.default_rng(G)
    cache = torch.randn(num_blocks, G, device=DEV).half()
    cache[5] = 0
    table = rng.permutation(num_blocks)[:n_sel].astype(np.int32)
    if 5 not in table:
        table[0] = 5
    td = torch.from_numpy(table).to(DEV)
    ids = (np.arange(n_sel, dtype=np.uint64) + np.uint64(7)) << np.uint64(12)
    tier = HostTier(pool_bytes=int(n_sel * G * 2 * 1.6) + (1 << 20))
    try:
        tier.offload_blocks(cache, td, ids)
        assert tier.stats()["blocks"] == n_sel
        want = codec.decompress(codec.compress(cache[td.long()].contiguous(), G))   # staged path on the same blocks
        # come back into another cache, through another table, in another order
        dst = torch.full((num_blocks, G), 0x1234, dtype=torch.int16, device=DEV).view(torch.float16)
        order = rng.permutation(n_sel)
        dst_table = rng.permutation(num_blocks)[:n_sel].astype(np.int32)
        tier.restore_blocks(ids[order], dst, torch.from_numpy(dst_table).to(DEV))
        got = dst.view(torch.int16)
        assert torch.equal(got[torch.from_numpy(dst_table).to(DEV).long()],
                           want.view(torch.int16)[torch.from_numpy(order).to(DEV)])
        untouched = np.setdiff1d(np.arange(num_blocks), dst_table)
        assert (got[torch.from_numpy(untouched).to(DEV)] == 0x1234).all()
        # the contiguous entry points see the same stored blocks
        y = tier.restore(ids[:16], G, torch.float16)
        assert torch.equal(y.view(torch.int16), want[:16].view(torch.int16))
    finally:
        tier.close()


def test_allocator_block_table_helpers():
    """CxlSpeckvKVAllocator.offload_kv_blocks / restore_kv_blocks on a vLLM-shaped cache
    ([num_blocks, block_size 16, kv_heads 8, head_dim 128] fp16: one block = one 16384-element group)."""
    from cxl_speckv_b200 import CxlSpeckvKVAllocator
    torch.manual_seed(3)
    cache = torch.randn(300, 16, 8, 128, device=DEV).half()
    keep = cache.clone()
    table = torch.tensor([7, 3, 250, 11, 12, 13, 299, 0], dtype=torch.int32, device=DEV)
    tier = HostTier(pool_bytes=64 << 20)
    try:
        CxlSpeckvKVAllocator.offload_kv_blocks(cache, table, tier)
        cache[table.long()] = 0                                      # the blocks are gone from HBM
        CxlSpeckvKVAllocator.restore_kv_blocks(cache, table, tier)
        want = codec.decompress(codec.compress(keep[table.long()].reshape(-1), 16 * 8 * 128)).view(8, 16, 8, 128)
        assert torch.equal(cache[table.long()].view(torch.int16), want.view(torch.int16))
        others = torch.ones(300, dtype=torch.bool, device=DEV)
        others[table.long()] = False
        assert torch.equal(cache[others].view(torch.int16), keep[others].view(torch.int16))
    finally:
        tier.close()
