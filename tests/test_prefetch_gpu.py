"""GPU: batched LSTM prefetch scoring + decompress of the predicted blocks against the reference's
recorded predictions (tests/golden, LSTMPredictor / SpeculativePrefetcher with srand(1) weights)
and the CPU oracle.  Bar: same top-k ids (up to permutations inside an
exact confidence tie -- the reference's std::sort is unstable), |confidence difference| <= 1e-7,
request addresses bit-exact."""
import numpy as np
import pytest
import torch

from oracle.oracle import Port
from tests.test_oracle_golden import same_topk

pytestmark = pytest.mark.gpu

from cxl_speckv_b200 import codec, prefetch  # noqa: E402

DEV = "cuda:0"
CONF_TOL = 1e-7


@pytest.fixture(scope="module")
def weights():
    emb, wout = Port.lstm_weights(1)
    prefetch.load_predictor(emb, wout, layers=2, history_len=16)
    return emb, wout


def test_golden_predictions(golden, weights):
    m = golden["meta"]["lstm"]
    hists = [p["hist"] for p in m["predictions"]]
    k = len(m["predictions"][0]["ids"])
    toks = torch.from_numpy(prefetch.window_history(hists).astype(np.int32)).to(DEV)
    ids, conf, va = prefetch.score(toks, k=k, layer_id=5)
    ids, conf, va = ids.cpu().numpy(), conf.cpu().numpy(), va.cpu().numpy()
    worst = 0.0
    for i, p in enumerate(m["predictions"]):
        want_conf = np.array(p["conf_bits"], dtype=np.uint32).view(np.float32)
        worst = max(worst, float(np.abs(conf[i] - want_conf).max()))
        assert np.abs(conf[i] - want_conf).max() <= CONF_TOL, p["hist"]
        # ids: identical wherever the reference's confidences are distinct
        assert same_topk(ids[i].tolist(), p["ids"], p["conf_bits"]), (ids[i].tolist(), p["ids"])
        assert va[i].tolist() == [(5 << 16) | (j + 1) for j in range(k)]
    pf = m["prefetch"]
    ids4, conf4, va4 = prefetch.score(toks[:1], k=pf["depth"], layer_id=pf["layer"])
    assert va4.cpu().numpy()[0].tolist() == pf["va"]
    assert same_topk(ids4.cpu().numpy()[0].tolist(), pf["tok"], pf["conf_bits"])
    print("max |conf - reference| =", worst)


def test_batch256_vs_oracle(weights):
    emb, wout = weights
    rng = np.random.default_rng(7)
    hists = rng.integers(0, 32000, (256, 16)).astype(np.uint32)       # BASELINE config 5 geometry
    hists[3, :5] = 40000                                              # out-of-vocabulary ids embed to zero
    ids, conf, va = prefetch.score(torch.from_numpy(hists.astype(np.int32)).to(DEV), k=4, layer_id=2)
    ids, conf = ids.cpu().numpy(), conf.cpu().numpy()
    for b in list(range(0, 256, 23)) + [3]:
        oi, oc, _ = Port.lstm_predict(emb, wout, hists[b], k=4)
        assert np.abs(conf[b] - oc).max() <= CONF_TOL
        assert same_topk(ids[b].tolist(), oi.tolist(), oc.view(np.uint32).tolist()), b
    assert (np.diff(conf, axis=1) <= 0).all()                          # sorted by confidence
    assert (va.cpu().numpy() == np.array([(2 << 16) | (i + 1) for i in range(4)])).all()


def test_predicted_blocks_are_decompressed(weights):
    """config 5: predicted tokens select KV blocks; only those blocks are decoded."""
    rng = np.random.default_rng(11)
    G, n_blocks = 131072, 96
    x = torch.randn(n_blocks * G, device=DEV).half()
    c = codec.compress(x, G)
    hists = rng.integers(0, 32000, (64, 16)).astype(np.int32)
    ids, conf, va = prefetch.score(torch.from_numpy(hists).to(DEV), k=4)
    block_index = (ids.view(-1) % n_blocks).to(torch.int32)
    y = codec.decompress_indexed(c, block_index)
    full = codec.decompress(c)
    assert torch.equal(y.view(torch.int16), full[block_index.long()].view(torch.int16))
    # paged geometry + generic path (odd group size)
    for G2 in (2048, 1000):
        x2 = torch.randn(200 * G2, device=DEV).half()
        x2[5 * G2:6 * G2] = 0.25
        c2 = codec.compress(x2, G2)
        idx = torch.tensor([5, 0, 199, 5, 17], dtype=torch.int32, device=DEV)
        y2 = codec.decompress_indexed(c2, idx)
        assert torch.equal(y2.view(torch.int16), codec.decompress(c2)[idx.long()].view(torch.int16))


def test_rank_one_scoring_equals_the_exact_kernels(weights, monkeypatch):
    """The rank-one candidate pass + exact re-scoring (default) against the same kernel with exact logits everywhere
    (SPECKV_SCORE_EXACT=1) and the tiled exact kernel (=2): identical ids, confidences within the tolerance."""
    rng = np.random.default_rng(41)
    for B, k in ((256, 4), (33, 7), (5, 12), (3, 1)):
        toks = torch.from_numpy(rng.integers(0, 32000, (B, 16)).astype(np.int32)).to(DEV)
        out = {}
        for mode in ("0", "1", "2"):
            monkeypatch.setenv("SPECKV_SCORE_EXACT", mode)
            ids, conf, _ = prefetch.score(toks, k=k)
            out[mode] = (ids.cpu().numpy().copy(), conf.cpu().numpy().copy())
        monkeypatch.delenv("SPECKV_SCORE_EXACT")
        for mode in ("1", "2"):
            assert np.array_equal(out["0"][0], out[mode][0]), (B, k, mode)
            assert np.abs(out["0"][1] - out[mode][1]).max() <= CONF_TOL


def test_rank_one_scoring_falls_back_on_wide_ties(weights):
    """Forty identical best rows (an exact tie wider than the candidate list) and an all-out-of-vocabulary window (h = 0,
    every logit 0): the candidate pass cannot separate them, the CTA repeats the pass with exact logits -- ids are the
    lowest row ids, as the oracle's."""
    emb, wout = weights
    w2 = wout.copy()
    big = np.abs(w2[123]) + 0.07
    rows = np.arange(500, 540)
    w2[rows] = big
    try:
        prefetch.load_predictor(emb, w2, layers=2, history_len=16)
        rng = np.random.default_rng(43)
        hists = rng.integers(0, 32000, (6, 16)).astype(np.uint32)
        hists[2, :] = 50000                                       # embeds to zeros: h = 0
        ids, conf, _ = prefetch.score(torch.from_numpy(hists.astype(np.int32)).to(DEV), k=4)
        ids, conf = ids.cpu().numpy(), conf.cpu().numpy()
        for b in range(6):
            oi, oc, _ = Port.lstm_predict(emb, w2, hists[b], k=4)
            assert np.abs(conf[b] - oc).max() <= CONF_TOL
            assert same_topk(ids[b].tolist(), oi.tolist(), oc.view(np.uint32).tolist()), (b, ids[b], oi)
        assert ids[2].tolist() == [0, 1, 2, 3]
    finally:
        prefetch.load_predictor(emb, wout, layers=2, history_len=16)


# ---- SpeculativePrefetcher::prefetch on the device: residency filter, request records, statistics --------------
def _ref_prefetcher():
    from oracle.oracle import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref was not built")
    import ctypes as C
    L = Ref.lib()
    return L, C.c_void_p(L.ref_prefetcher_new(1, 4, 16)), C


def _ref_prefetch(L, pf, C, hist, layer, depth):
    h = np.ascontiguousarray(hist, dtype=np.uint32)
    va, lay, tok = np.zeros(64, np.uint64), np.zeros(64, np.uint32), np.zeros(64, np.uint32)
    conf = np.zeros(64, np.float32)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    n = L.ref_prefetcher_prefetch(pf, p(h, C.c_uint32), h.size, layer, depth, p(va, C.c_uint64), p(lay, C.c_uint32),
                                  p(tok, C.c_uint32), p(conf, C.c_float))
    return n, va[:n], lay[:n], tok[:n], conf[:n]


def test_emit_with_populated_page_table_matches_reference(weights):
    """speculative_prefetcher.cpp:41-67 against the reference's own SpeculativePrefetcher holding a populated
    CXLMemoryManager: requests whose page sits in L1 or L2 are skipped, the rest become PrefetchRequest records.
    The reference's request address (req 0) only reaches its allocator's range (from 0x1_0000_0000) for layer ids
    >= 0x10000: layer 0x10000 + j names the page at va_base + j * 65536."""
    from cxl_speckv_b200.tier import CxlAddressMap, L1, L2, L3
    L, pf, C = _ref_prefetcher()
    amap = CxlAddressMap()
    try:
        # 16-page allocations: allocation j covers the address of layer 0x10000 + j; tiers L1, L3, L2, L3, L1 ...
        tiers = [L1, L3, L2, L3, L1, L3, L3, L2]
        for j, t in enumerate(tiers):
            va_ref = L.ref_prefetcher_mm_allocate(pf, 16 * 4096, j, t)
            va_own, used = amap.allocate(16 * 4096, j, t)
            assert va_ref == va_own == CxlAddressMap.VA_BASE + j * 65536 and used == t
        table = amap.export(DEV)
        rng = np.random.default_rng(23)
        hists = rng.integers(0, 32000, (len(tiers) + 2, 16)).astype(np.uint32)
        total = 0
        for j in range(len(tiers) + 2):                         # the last two layers lie beyond every allocation
            layer = 0x10000 + j
            n, rva, rlay, rtok, rconf = _ref_prefetch(L, pf, C, hists[j], layer, 4)
            toks = torch.from_numpy(hists[j:j + 1].astype(np.int32)).to(DEV)
            tab = prefetch.emit(toks, k=4, layer_id=layer, page_table=table, va_base=CxlAddressMap.VA_BASE, timestamp=9)
            m, va, lay, tok, conf, ts = prefetch.unpack_table(tab)
            assert m == n == (0 if j < len(tiers) and tiers[j] in (L1, L2) else 4), (j, m, n)
            assert np.array_equal(va, rva) and np.array_equal(lay, rlay)
            assert np.abs(conf - rconf).max(initial=0.0) <= CONF_TOL
            assert n == 0 or same_topk(tok.tolist(), rtok.tolist(), rconf.view(np.uint32).tolist())
            total += n
        # a batch with per-sequence request ids: sequences whose id moves the address out of the table are kept
        toks = torch.from_numpy(hists[:4].astype(np.int32)).to(DEV)
        req_ids = torch.tensor([0, 2, 0, 6], dtype=torch.int32, device=DEV)
        tab = prefetch.emit(toks, k=4, layer_id=0x10000, req_ids=req_ids, page_table=table, va_base=CxlAddressMap.VA_BASE)
        m, va, _, _, _, _ = prefetch.unpack_table(tab)
        assert m == 8 and sorted(set((va >> np.uint64(32)).tolist())) == [3, 7]   # (req << 32) | (0x10000 << 16): req 0 stays in page 0 (L1)
        # statistics follow the reference's counters (total_prefetches :72-79; mispredictions :84-97)
        cnt, rates = np.zeros(3, np.uint64), np.zeros(3, np.float64)
        L.ref_prefetcher_stats(pf, cnt.ctypes.data_as(C.POINTER(C.c_uint64)), rates.ctypes.data_as(C.POINTER(C.c_double)), 0)
        assert int(cnt[0]) == total
    finally:
        amap.close()
        L.ref_prefetcher_free(pf)


def test_prefetch_statistics_and_queue_follow_reference(weights):
    """handle_misprediction (:84-97), get/reset_statistics (:126-142), the 16-deep outstanding queue (:162-185)."""
    L, pf, C = _ref_prefetcher()
    try:
        prefetch.statistics(reset=True)
        rng = np.random.default_rng(5)
        hists = rng.integers(0, 32000, (7, 16)).astype(np.uint32)
        emitted = 0
        for j in range(7):                                          # 7 calls x 4 requests: the queue keeps the last 16
            n, rva, _, rtok, _ = _ref_prefetch(L, pf, C, hists[j], 3 + j, 4)
            tab, ids, _ = prefetch.emit(torch.from_numpy(hists[j:j + 1].astype(np.int32)).to(DEV), k=4, layer_id=3 + j,
                                        want_predictions=True)
            emitted += n
            pred = ids.cpu().numpy().view(np.uint32)[0]
            for actual in (int(pred[1]), 31999 - j):                # one hit, one miss
                L.ref_prefetcher_mispredict(pf, actual, rtok.ctypes.data_as(C.POINTER(C.c_uint32)), rtok.size)
                ok = prefetch.handle_misprediction(actual, pred)
                assert ok == (actual in pred.tolist())
        cnt, rates = np.zeros(3, np.uint64), np.zeros(3, np.float64)
        L.ref_prefetcher_stats(pf, cnt.ctypes.data_as(C.POINTER(C.c_uint64)), rates.ctypes.data_as(C.POINTER(C.c_double)), 0)
        st = prefetch.statistics()
        assert st["total_prefetches"] == int(cnt[0]) == emitted == 28
        assert st["mispredictions"] == int(cnt[2]) == 7
        assert st["successful_prefetches"] == int(cnt[1]) == 0 and st["hit_rate"] == rates[0] == 0.0
        assert st["avg_prediction_latency_us"] > 0.0
        qva, qtok = np.zeros(32, np.uint64), np.zeros(32, np.uint32)
        qn = L.ref_prefetcher_queue(pf, qva.ctypes.data_as(C.POINTER(C.c_uint64)), qtok.ctypes.data_as(C.POINTER(C.c_uint32)), 32)
        probe = [int(v) for v in qva[:qn]] + [(3 << 16) | 1, (4 << 16) | 4, 12345]      # the oldest requests fell out
        found, queue = prefetch.outstanding(probe)
        want = [bool(L.ref_prefetcher_is_outstanding(pf, v)) for v in probe]
        assert found.tolist() == want and want[-3:] == [False, False, False] and all(want[:qn])
        assert [q["virtual_addr"] for q in queue] == [int(v) for v in qva[:qn]] and len(queue) == qn == 16
        prefetch.statistics(reset=True)
        L.ref_prefetcher_stats(pf, cnt.ctypes.data_as(C.POINTER(C.c_uint64)), rates.ctypes.data_as(C.POINTER(C.c_double)), 1)
        L.ref_prefetcher_stats(pf, cnt.ctypes.data_as(C.POINTER(C.c_uint64)), rates.ctypes.data_as(C.POINTER(C.c_double)), 0)
        st = prefetch.statistics()
        assert st["total_prefetches"] == int(cnt[0]) == 0 and st["mispredictions"] == int(cnt[2]) == 0
    finally:
        L.ref_prefetcher_free(pf)


def test_route_and_routed_decode(weights):
    """Device-side request routing (contiguous block ownership) + decode with the request count on the device:
    the same blocks as the host-side selection, for every (world, rank); nothing beyond the count is written."""
    rng = np.random.default_rng(31)
    B, k, G, nb = 64, 4, 131072, 96
    toks = torch.from_numpy(rng.integers(0, 32000, (B, 16)).astype(np.int32)).to(DEV)
    table = prefetch.emit(toks, k=k, layer_id=1, timestamp=5)
    n, va, layer, tok, conf, ts = prefetch.unpack_table(table)
    assert n == B * k and (ts == 5).all() and (layer == 1).all()
    blocks = tok.astype(np.int64) % 4096
    for world in (1, 2, 4, 8):
        per = (4096 + world - 1) // world
        # all-gathered layout: `world` tables in a row (here the same table repeated)
        tables = table.unsqueeze(0).repeat(world, 1, 1).contiguous()
        for rank in range(world):
            ri = torch.zeros(world * B * k, dtype=torch.int32, device=DEV)
            bi, cnt = prefetch.route(tables, world, B * k, 4096, world, rank, request_index=ri)
            want = np.tile(blocks, world)
            keep = want // per == rank
            got = bi.cpu().numpy()[:int(cnt.item())]
            assert np.array_equal(got.astype(np.int64), want[keep] - rank * per), (world, rank)
            assert np.array_equal(ri.cpu().numpy()[:got.size], np.nonzero(keep)[0])
    x = torch.randn(nb * G, device=DEV).half()
    c = codec.compress(x, G)
    full = codec.decompress(c)
    bi, cnt = prefetch.route(table, 1, B * k, nb, 1, 0)
    out = torch.zeros((B * k, G), dtype=torch.float16, device=DEV)
    codec.decompress_routed(c, bi, cnt, out)
    assert torch.equal(out.view(torch.int16), full[bi.long()].view(torch.int16))
    short = torch.tensor([37], dtype=torch.int32, device=DEV)
    out.zero_()
    codec.decompress_routed(c, bi, short, out)
    assert torch.equal(out[:37].view(torch.int16), full[bi[:37].long()].view(torch.int16)) and not out[37:].any()
    # an empty request list decodes nothing
    out.zero_()
    codec.decompress_routed(c, bi, torch.zeros(1, dtype=torch.int32, device=DEV), out)
    assert not out.any()


def test_engine_latency_statistics():
    """EngineStatistics (cache_engine.cpp:65-79,103-112,150-158): counts per group, running means per call."""
    G = 131072
    x = torch.randn(32 * G, device=DEV).half()
    s0 = codec.engine_stats()
    c = codec.compress(x, G)
    codec.decompress(c)
    codec.decompress(c)
    s1 = codec.engine_stats()
    assert s1["total_compressions"] - s0["total_compressions"] == 32
    assert s1["total_decompressions"] - s0["total_decompressions"] == 64
    assert s1["compress_calls_timed"] - s0["compress_calls_timed"] == 1
    assert s1["decompress_calls_timed"] - s0["decompress_calls_timed"] == 2
    assert s1["avg_compression_latency_ns"] > 0 and s1["avg_decompression_latency_ns"] > 0 and s1["throughput_gbps"] > 0
