"""Ad-hoc GPU check of the round-2 entry points (scoring/emit/route/routed decode) + timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from cxl_speckv_b200 import codec, prefetch
from oracle.oracle import Port
from tests.test_oracle_golden import same_topk

dev = "cuda:0"
emb, wout = Port.lstm_weights(1)
prefetch.load_predictor(emb, wout)
rng = np.random.default_rng(7)
for B in (256, 32, 5, 700):
    hists = rng.integers(0, 32000, (B, 16)).astype(np.uint32)
    toks = torch.from_numpy(hists.astype(np.int32)).to(dev)
    for k in (4, 8, 16, 3):
        ids, conf, va = prefetch.score(toks, k=k, layer_id=2)
        torch.cuda.synchronize()
        ids, conf = ids.cpu().numpy(), conf.cpu().numpy()
        bad = 0
        for b in range(0, B, max(1, B // 6)):
            oi, oc, _ = Port.lstm_predict(emb, wout, hists[b], k=k)
            if np.abs(conf[b] - oc).max() > 1e-7 or not same_topk(ids[b].tolist(), oi.tolist(), oc.view(np.uint32).tolist()):
                bad += 1
                print("MISMATCH", B, k, b, ids[b], oi, conf[b], oc)
        print(f"score B={B} k={k}: {'ok' if not bad else 'BAD'}")
# timing
for B in (256, 32):
    toks = torch.from_numpy(rng.integers(0, 32000, (B, 16)).astype(np.int32)).to(dev)
    for _ in range(3): prefetch.score(toks, k=4)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): prefetch.score(toks, k=4)
    b.record(); torch.cuda.synchronize()
    print(f"score B={B}: {a.elapsed_time(b)/20*1e3:.1f} us per call")
# emit + route + routed decode
B, k = 256, 4
toks = torch.from_numpy(rng.integers(0, 32000, (B, 16)).astype(np.int32)).to(dev)
table, ids, conf = prefetch.emit(toks, k=k, layer_id=3, timestamp=77, want_predictions=True)
n, va, layer, tok, cf, ts = prefetch.unpack_table(table)
assert n == B * k and (tok.reshape(B, k) == ids.cpu().numpy().view(np.uint32)).all() and (ts == 77).all() and (layer == 3).all()
assert (va.reshape(B, k) == np.array([(3 << 16) | (i + 1) for i in range(k)], dtype=np.uint64)).all()
print("emit ok", n, prefetch.statistics(), prefetch.outstanding([int(va[-1]), 12345])[0])
G, nb = 131072, 512
x = torch.randn(nb * G, device=dev).half()
c = codec.compress(x, G)
full = codec.decompress(c)
for world in (1, 2, 8):
    for rank in range(world):
        per = (4096 + world - 1) // world
        bi, cnt = prefetch.route(table, 1, B * k, 4096, world, rank)
        blocks = tok.astype(np.int64) % 4096
        want = (blocks[blocks // per == rank] - rank * per)
        got = bi.cpu().numpy()[:int(cnt.item())]
        assert np.array_equal(got.astype(np.int64), want), (world, rank)
bi, cnt = prefetch.route(table, 1, B * k, nb, 1, 0)
out = torch.zeros((B * k, G), dtype=torch.float16, device=dev)
codec.decompress_routed(c, bi, cnt, out)
torch.cuda.synchronize()
assert torch.equal(out.view(torch.int16), full[bi.long()].view(torch.int16))
cnt2 = torch.tensor([100], dtype=torch.int32, device=dev)
out.zero_()
codec.decompress_routed(c, bi, cnt2, out)
assert torch.equal(out[:100].view(torch.int16), full[bi[:100].long()].view(torch.int16)) and not out[100:].any()
print("route + routed decode ok;", codec.engine_stats())
