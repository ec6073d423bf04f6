"""CPU: the arithmetic behind the tuned compress kernel's long-run path (kv_codec_fast.cu: forced_head_in,
long_region_scan, long_reduce, long_emit), restated in a few lines of Python and checked against the oracle's
sequential run-length encoder (oracle_rle_encode = cache_engine.cpp:213-239, itself pinned to the reference).

The kernel never walks a run: per 8-element chunk it only knows nb, the position of the last NATURAL run boundary
before the chunk, and derives from it (1) the forced boundaries of the 255 cap, nb + 255 k, (2) the start of the run
a boundary closes, nb + 255 * floor((h - 1 - nb) / 255), and (3), per region, how many pairs all lower regions emit
(their own count of natural + interior forced boundaries, plus the forced boundaries of their leading stretch, which
follow from the exchanged (first, last) natural boundary of every region).  This test checks exactly those three
closed forms on random delta streams full of long runs."""
import ctypes as C

import numpy as np
import pytest

from oracle.oracle import Port

CHUNK = 8


def rle_oracle(delta: np.ndarray) -> np.ndarray:
    L = Port.lib()
    L.oracle_rle_encode.restype = C.c_size_t
    L.oracle_rle_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    d = np.ascontiguousarray(delta.astype(np.uint8)).view(np.int8)
    out = np.zeros(2 * d.size + 2, np.uint8)
    n = L.oracle_rle_encode(d.ctypes.data, d.size, out.ctypes.data)
    return out[:n].reshape(-1, 2)


def forced_head_in(c, lead_len, nb):
    if nb < 0 or lead_len <= 0:
        return -1
    r = (c - nb) % 255
    pf = c if r == 0 else c + 255 - r
    return pf if pf < c + lead_len else -1


def natural_heads(delta):
    nat = np.ones(delta.size, bool)
    nat[1:] = delta[1:] != delta[:-1]
    return nat


def emit_by_chunks(delta):
    """long_emit: pairs from per-chunk knowledge only."""
    n = delta.size
    nat = natural_heads(delta)
    pairs, nb = [], -1
    for c in range(0, n, CHUNK):
        m = nat[c:c + CHUNK]
        lead_len = int(np.argmax(m)) if m.any() else CHUNK
        pf = forced_head_in(c, lead_len, nb)
        prev = nb + 255 * ((c - 1 - nb) // 255) if nb >= 0 else 0
        for j in range(CHUNK):
            p = c + j
            if m[j] or p == pf:
                if p > 0:
                    pairs.append((int(delta[p - 1]), p - prev))
                prev = p
        if m.any():
            nb = c + int(np.flatnonzero(m)[-1])
    prev_end = nb + 255 * ((n - 1 - nb) // 255)
    pairs.append((int(delta[n - 1]), n - prev_end))
    return np.array(pairs, np.int64).reshape(-1, 2)


def region_words(delta, region):
    """What every region publishes: (natural + interior forced boundaries, first, last natural boundary or -1)."""
    nat = natural_heads(delta)
    words = []
    for a0 in range(0, delta.size, region):
        count, nb, first = 0, -1, region
        for c in range(a0, a0 + region, CHUNK):
            m = nat[c:c + CHUNK]
            lead_len = int(np.argmax(m)) if m.any() else CHUNK
            if nb >= 0 and forced_head_in(c, lead_len, nb) >= 0:        # interior: a natural boundary of THIS region precedes it
                count += 1
            count += int(m.sum())
            if m.any():
                if first == region:
                    first = c - a0 + lead_len
                nb = c + int(np.flatnonzero(m)[-1])
        words.append((count, first, nb - a0 if nb >= 0 else -1))
    return words


def pairs_before_regions(words, region):
    """long_reduce: pairs emitted by the lower regions, for every region, from the published words alone."""
    out, total, nb = [], 0, -1
    for j, (count, first, last) in enumerate(words):
        out.append(total)
        lead = 0
        if j > 0 and nb >= 0:
            a0, bnd = j * region, j * region + min(first, region)
            lead = (bnd - 1 - nb) // 255 - (a0 - 1 - nb) // 255
        total += count + lead
        if last >= 0:
            nb = j * region + last
    return out, total


def random_deltas(rng, n):
    d = rng.integers(0, 256, n).astype(np.int64)
    pos = 0
    while pos < n:
        seg = int(rng.choice([1, 2, 7, 8, 9, 16, 17, 100, 254, 255, 256, 300, 509, 510, 511, 700, 2000]))
        if rng.random() < 0.6:
            d[pos:pos + seg] = d[pos]
        pos += seg + int(rng.integers(0, 30))
    return d


@pytest.mark.parametrize("seed", range(6))
def test_chunk_closed_forms_reproduce_the_sequential_encoder(seed):
    rng = np.random.default_rng(seed)
    for _ in range(40):
        n = CHUNK * int(rng.integers(1, 600))
        d = random_deltas(rng, n)
        want = rle_oracle(d).astype(np.int64)
        got = emit_by_chunks(d)
        assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("region", [64, 256, 2048])
def test_region_words_give_every_region_its_offset(region):
    rng = np.random.default_rng(region)
    for _ in range(25):
        n = region * int(rng.integers(1, 9))
        d = random_deltas(rng, n)
        pairs = emit_by_chunks(d)                       # checked against the oracle above
        # pairs emitted at boundaries inside lower regions (the pair of a boundary at position p sits at index #boundaries < p, minus the first)
        nat = natural_heads(d)
        emitted = np.zeros(n, bool)
        nb = -1
        for c in range(0, n, CHUNK):
            m = nat[c:c + CHUNK]
            lead_len = int(np.argmax(m)) if m.any() else CHUNK
            pf = forced_head_in(c, lead_len, nb)
            emitted[c:c + CHUNK] = m
            if pf >= 0:
                emitted[pf] = True
            if m.any():
                nb = c + int(np.flatnonzero(m)[-1])
        before, total = pairs_before_regions(region_words(d, region), region)
        for j, b in enumerate(before):
            assert b == int(emitted[:j * region].sum()), (region, j)
        assert total == int(emitted.sum()) == pairs.shape[0]      # boundaries = pairs (position 0 emits none, the end emits one)
