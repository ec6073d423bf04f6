"""CPU: the in-place staging invariant of the tuned kernels, simulated on byte addresses.

Both kernels build their output in the shared-memory tile they are still reading: the output stream trails the read
position, and nothing may be written onto a slot that has not been read yet (kv_codec_fast.cu: "staged IN PLACE").
* compress: pair i of a region is staged at tile - 16 + 2 * (i - (p0 & ~7)); position e of the region is read at
  tile + 2 * e; every position emits at most one pair.
* decompress, one-region groups (single pass): element i is staged at tile - 512 + 2 * i, pair j is read at tile + 2 * j,
  and the kernel gives the region up (generic kernel) as soon as an iteration would end beyond pk + 512 - 8 elements.
The simulation replays those rules on random inputs and checks, per warp iteration of 256 positions, that every byte
written lies below the first unread slot and inside the tile's own pad + body."""
import numpy as np
import pytest

PAD, TILE, IT = 512, 4096, 256


@pytest.mark.parametrize("seed", range(4))
def test_compress_pairs_never_reach_unread_slots(seed):
    rng = np.random.default_rng(seed)
    for _ in range(300):
        p0 = int(rng.integers(0, 5000))                      # pairs emitted by lower regions
        p_hole = float(rng.choice([0.0, 1 / 256, 0.05, 0.5, 0.95]))
        heads = rng.random(2048) >= p_hole                   # position emits a pair (natural or forced head)
        base = -16 - 2 * (p0 & ~7)                           # byte offset of pair 0 relative to the tile
        idx = p0
        for k in range(8):
            n = int(heads[k * IT:(k + 1) * IT].sum())
            lo, hi = base + 2 * idx, base + 2 * (idx + n)    # bytes written in this iteration
            assert lo >= -16 and hi <= 512 * (k + 1), (p0, k, lo, hi)      # below slot k + 1, not before the 16-byte lead
            idx += n
        assert base + 2 * idx <= TILE and base + 2 * p0 >= -PAD


def single_pass_decode(counts):
    """the kernel's bookkeeping for a one-region group: returns (accepted, list of (k, first byte, end byte))"""
    ecur, writes = 0, []
    for pk in range(0, counts.size, IT):
        c = counts[pk:pk + IT]
        lanes = c.reshape(-1, 8) if c.size % 8 == 0 else np.pad(c, (0, IT - c.size)).reshape(-1, 8)
        small = np.isin(lanes, (1, 2)).all(axis=1) & ((lanes == 2).sum(axis=1) <= 1)
        tot = int(c.sum())
        if small.all():
            if ecur + 256 + int(((lanes == 2).sum(axis=1) > 0).sum()) > pk + 512 - 8:
                return False, writes
        elif ecur + tot > pk + 512 - 8:
            return False, writes
        writes.append((pk // IT, -PAD + 2 * ecur, -PAD + 2 * (ecur + tot)))
        ecur += tot
    return ecur <= 2048, writes


@pytest.mark.parametrize("seed", range(4))
def test_single_pass_decode_bound_keeps_writes_behind_reads(seed):
    rng = np.random.default_rng(100 + seed)
    accepted = 0
    for _ in range(600):
        npairs = int(rng.integers(65, 2049))
        style = rng.integers(0, 4)
        if style == 0:
            counts = np.ones(npairs, np.int64)
            counts[rng.random(npairs) < rng.choice([0.0, 1 / 256, 0.03, 0.12])] = 2
        elif style == 1:
            counts = rng.choice(np.array([0, 1, 1, 1, 2, 3]), npairs)
        elif style == 2:
            counts = np.ones(npairs, np.int64)
            counts[rng.integers(0, npairs, 3)] = rng.integers(2, 256, 3)
        else:
            counts = rng.integers(0, 4, npairs)
        ok, writes = single_pass_decode(counts)
        accepted += ok
        for k, lo, hi in writes:
            assert lo >= -PAD and hi <= 512 * (k + 1), (k, lo, hi)          # below the first unread slot
            assert hi <= TILE                                               # and inside the tile
    assert accepted > 50
