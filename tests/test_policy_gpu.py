"""GPU: tier residency policy (speckv_ext_policy_*) against the oracle restatement of
CXLMemoryManager's policy (oracle_policy_*, itself pinned against the reference in
tests/test_oracle_golden.py) on random traces, including L1 capacities small enough to force
evictions -- the path the reference cannot run (its evict_l1_lru re-locks a held mutex)."""
import random

import numpy as np
import pytest
import torch

from oracle.oracle import run_policy_trace

pytestmark = pytest.mark.gpu

from cxl_speckv_b200.tier import TierPolicy  # noqa: E402

DEV = "cuda:0"


def make_trace(seed, n_pages, n_batches, max_batch, release=True):
    """A trace as a list of batches (op, [pages], tier) -- one C-ABI call each."""
    rng = random.Random(seed)
    batches = [("place", [g for g in range(n_pages) if rng.random() < 0.9], None)]
    batches[0] = ("place_mixed", [(g, rng.choice([0, 1, 2, 2])) for g in batches[0][1]], None)
    for _ in range(n_batches):
        r = rng.random()
        k = rng.randint(1, max_batch)
        pages = [rng.randrange(n_pages + 2) for _ in range(k)]        # a few ids out of range
        if r < 0.45:
            batches.append(("touch", pages, None))
        elif r < 0.55:
            batches.append(("hot", pages, None))
        elif r < 0.75:
            batches.append(("promote", pages, None))
        elif r < 0.88:
            batches.append(("demote", pages, None))
        elif r < 0.93 and release:
            batches.append(("release", pages[:3], None))
        else:
            batches.append(("place", pages[:4], rng.choice([0, 1, 2])))
    return batches


def run_gpu(batches, n_pages, caps, touch_on_device):
    pol = TierPolicy(n_pages, *caps)
    results = []
    try:
        for op, pages, tier in batches:
            if op == "place_mixed":
                for t in (0, 1, 2):          # same order as the flattened trace below: by tier
                    ids = [g for g, tt in pages if tt == t]
                    results += [int(x) if x != 255 else -1 for x in pol.place(ids, t)]
            elif op == "place":
                results += [int(x) if x != 255 else -1 for x in pol.place(pages, tier)]
            elif op == "touch":
                if touch_on_device:
                    pol.touch(torch.tensor(pages, dtype=torch.int64, device=DEV))
                else:
                    pol.touch(pages)
                results += [0] * len(pages)
            elif op == "hot":
                results += pol.is_hot(torch.tensor(pages, dtype=torch.int64, device=DEV)).cpu().tolist()
            elif op == "promote":
                ok, ev = pol.promote(pages)
                results.append(("promote", ok.tolist(), ev.tolist()))
            elif op == "demote":
                results += pol.demote(pages).tolist()
            elif op == "release":
                pol.release(pages)
                results += [0] * len(pages)
        return {"results": results, "tiers": pol.tiers().tolist(), "lru": pol.lru_order().tolist(), "stats": pol.stats()}
    finally:
        pol.close()


def run_oracle(batches, n_pages, caps):
    """Flattens the batches into the oracle's one-page-at-a-time calls."""
    ops, shape = [], []
    for op, pages, tier in batches:
        if op == "place_mixed":
            for t in (0, 1, 2):
                for g, tt in pages:
                    if tt == t:
                        ops.append(("place", g, t)); shape.append("scalar")
        elif op == "promote":
            shape.append(("promote", len(pages)))
            ops += [("promote", g) for g in pages]
        else:
            for g in pages:
                ops.append((op, g, tier) if op == "place" else (op, g)); shape.append("scalar")
    # pages >= n_pages: the oracle bounds-checks like the library (place -> -1, others no-ops)
    r = run_policy_trace(ops, n_pages, caps, impl="port")
    flat, i = [], 0
    for s in shape:
        if s == "scalar":
            flat.append(r["results"][i]); i += 1
        else:
            chunk = r["results"][i:i + s[1]]; i += s[1]
            flat.append(("promote", [ok for ok, _ in chunk], [ev for _, ev in chunk if ev is not None]))
    r["results"] = flat
    return r


@pytest.mark.parametrize("seed,n_pages,caps,on_dev", [
    (0, 64, (1 << 40, 1 << 40, 1 << 40), False),      # nothing ever full
    (1, 64, (8, 1 << 40, 1 << 40), True),             # constant eviction pressure
    (2, 300, (40, 1 << 40, 1 << 40), True),
    (3, 17, (1, 1 << 40, 1 << 40), False),            # single-slot L1
    (4, 33, (0, 1 << 40, 1 << 40), True),             # no L1 at all: every promotion evicts the previous one
    (5, 5000, (100, 1 << 40, 1 << 40), True),
])
def test_policy_trace_matches_oracle(seed, n_pages, caps, on_dev):
    batches = make_trace(seed, n_pages, n_batches=400, max_batch=min(64, 2 * n_pages))
    got = run_gpu(batches, n_pages, caps, on_dev)
    want = run_oracle(batches, n_pages, caps)
    assert got["results"] == want["results"]
    assert got["tiers"] == want["tiers"]
    assert got["lru"] == want["lru"]
    for k, v in want["stats"].items():
        assert got["stats"][k] == v, k


def test_policy_large_batches():
    """One decode step of a 32K-context model touches every page: a million ids in one call;
    then a promotion batch larger than L1 has room for."""
    n = 1 << 20
    pol = TierPolicy(n, l1_pages=4096)
    try:
        ids = np.arange(n, dtype=np.uint64)
        assert (pol.place(ids[:4096], 0) == 0).all()
        assert (pol.place(ids[4096:], 0) == 2).all()          # L1 full -> L3 (cxl_memory_manager.cpp:37-40)
        g = torch.Generator(device=DEV); g.manual_seed(5)
        perm = torch.randperm(n, device=DEV, generator=g)
        pol.touch(perm)                                        # every page once, random order
        pol.touch(perm[:12345])                                # some twice
        st = pol.stats()
        assert st["l1_hits"] + st["l3_accesses"] == n + 12345
        order = pol.lru_order()
        p = perm.cpu().numpy().astype(np.uint64)
        want = np.concatenate([p[12345:], p[:12345]])          # re-touched pages moved to the back
        assert np.array_equal(order, want)
        # promote 1000 L3 pages: L1 is full, so the 1000 least recently used L1 pages leave
        l3 = want[np.isin(want, ids[4096:])][:1000]
        ok, ev = pol.promote(l3)
        assert ok.all() and ev.size == 1000
        l1_in_order = want[np.isin(want, ids[:4096])]
        assert np.array_equal(ev, l1_in_order[:1000])
        t = pol.tiers()
        assert (t[l3.astype(np.int64)] == 0).all() and (t[ev.astype(np.int64)] == 2).all()
        assert pol.stats()["l1_pages"] == 4096
        hot = pol.is_hot(perm[:100])
        assert int(hot.sum()) == 0                              # two touches at most so far
        for _ in range(10):
            pol.touch(perm[:50])
        hot = pol.is_hot(perm[:100]).cpu().numpy()
        assert hot[:50].all() and not hot[50:].any()            # > 10 accesses (:253)
    finally:
        pol.close()


def test_policy_drives_the_pool():
    """residency_step(): pages the policy promotes are restored into the bound pool, the pages it
    evicts to make room are compressed out to the host tier (page-table flags follow)."""
    import ctypes as C

    import cxl_speckv_b200 as pkg
    from cxl_speckv_b200 import CxlSpeckvKVAllocator, codec
    from cxl_speckv_b200.tier import HostTier

    alloc = CxlSpeckvKVAllocator(pkg.lib_path(), "cuda:0")
    L = pkg.lib()
    tier = HostTier(16 << 20)
    pol = None
    try:
        tokens, layers, heads, hd = 128, 2, 8, 128
        h = alloc.allocate(tokens, layers, heads, hd, 2)
        total = tokens * layers * heads * hd * 2 * 2
        n_pages = total // 4096                                         # 256
        torch.manual_seed(9)
        pool = torch.randn(total // 2, device=DEV).half()
        want = codec.decompress(codec.compress(pool, 2048)).clone().view(n_pages, 2048)
        alloc.bind_pool(pool, tier)
        pol = TierPolicy(n_pages, l1_pages=64)
        alloc.attach_policy(pol)
        ids = np.arange(n_pages)
        assert (pol.place(ids[:64], 0) == 0).all() and (pol.place(ids[64:], 2) == 2).all()
        alloc.offload_pages(64, n_pages - 64)                           # L3 pages live in the host tier only
        pool.view(n_pages, 2048)[64:] = 0
        alloc.residency_step(touched=torch.arange(64, device=DEV))      # L1 pages used in order 0..63
        ok, ev = alloc.residency_step(touched=torch.tensor([0, 1, 2], device=DEV), promote=[100, 101, 102, 200])
        assert list(ok) == [1, 1, 1, 1] and list(ev) == [3, 4, 5, 6]    # 0..2 were re-touched: 3..6 are the LRU
        got = pool.view(n_pages, 2048)
        for pg in (100, 101, 102, 200):
            assert torch.equal(got[pg].view(torch.int16), want[pg].view(torch.int16))
        assert (got[103] == 0).all()                                    # not promoted: still only in the tier
        tbl = torch.zeros(n_pages * 3, dtype=torch.int64, device=DEV)
        cnt = C.c_size_t()
        assert L.speckv_ext_page_table_export(h, tbl.data_ptr(), n_pages, C.byref(cnt), None) == 0
        flags = (tbl.cpu().numpy().view(np.uint64).reshape(n_pages, 3)[:, 2] >> np.uint64(32)).astype(np.int64)
        assert all(flags[pg] == 4 for pg in (3, 4, 5, 6))               # evicted: compressed copy only
        assert all(flags[pg] == 6 for pg in (100, 101, 102, 200))       # restored: L2 | compressed
        assert pol.stats()["l1_pages"] == 64 and pol.stats()["migrations_l1_to_l3"] == 4
        # an evicted page comes back bit-exact when it is promoted again
        ok, ev = alloc.residency_step(promote=[4])
        assert list(ok) == [1] and list(ev) == [7]
        assert torch.equal(got[4].view(torch.int16), want[4].view(torch.int16))
        assert L.speckv_free(h) == 0
    finally:
        if pol is not None:
            pol.close()
        alloc._speckv.finalize()
        tier.close()
