"""CPU: bit-level models of the byte tricks in kv_codec_fast.cu, checked against the plain per-position definition.
Python integers masked to 32 bits stand in for the registers; every function below is a transcription of a few
lines of the kernel (named in its docstring), every check enumerates all cases of its small domain."""
import itertools
import random

M32 = 0xFFFFFFFF


def byte_perm(a, b, sel):
    """PRMT, default mode: result byte i = byte (sel nibble i) of the 8-byte pool {a: 0-3, b: 4-7}."""
    pool = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
    return sum(pool[(sel >> (4 * i)) & 7] << (8 * i) for i in range(4))


def funnelshift_r(lo, hi, s):
    return (((hi << 32) | lo) >> s) & M32


def funnelshift_l(lo, hi, s):
    return ((((hi << 32) | lo) << s) >> 32) & M32


def pack(bytes4):
    return sum((b & 0xFF) << (8 * i) for i, b in enumerate(bytes4))


def unpack(w):
    return [(w >> (8 * i)) & 0xFF for i in range(4)]


def test_head_mask8_multiply_trick():
    """head_mask8: bit 7 of byte j of (nz0, nz1) -> bit j of an 8-bit mask, via one multiply per word."""
    for m in range(256):
        nz0 = pack([0x80 if (m >> j) & 1 else 0 for j in range(4)])
        nz1 = pack([0x80 if (m >> (4 + j)) & 1 else 0 for j in range(4)])
        lo = ((((nz0 >> 7) & 0x01010101) * 0x01020408) & M32) >> 24
        hi = ((((nz1 >> 7) & 0x01010101) * 0x01020408) & M32) >> 24
        assert ((lo & 0xF) | ((hi & 0xF) << 4)) == m


def test_run_boundary_flags_swar():
    """phase 2a: nz = (((x & 0x7f7f7f7f) + 0x7f7f7f7f) | x) & 0x80808080 flags the non-zero bytes of x = d ^ dsh."""
    rng = random.Random(1)
    for _ in range(20000):
        x = pack([rng.choice([0, 0, rng.randrange(256)]) for _ in range(4)])
        nz = ((((x & 0x7F7F7F7F) + 0x7F7F7F7F) & M32) | x) & 0x80808080
        assert unpack(nz) == [0x80 if b else 0 for b in unpack(x)]


def emit_common_path(dsh, nz, lf):
    """phase 2b, common case (compress_fast_kernel): 8 positions, at most one of them continues a run (a "hole").
    dsh[j] = delta[pos_j - 1], nz[j] = position j starts a run, lf = distance from position 0 back to the previous
    head.  Returns the 16-bit units w0..w3 (two per word) and how many of them are stored."""
    dsh0, dsh1 = pack(dsh[:4]), pack(dsh[4:])
    nz0, nz1 = pack([0x80 if b else 0 for b in nz[:4]]), pack([0x80 if b else 0 for b in nz[4:]])
    hw0, hw1 = ~nz0 & 0x80808080, ~nz1 & 0x80808080
    nhole = bin(hw0).count("1") + bin(hw1).count("1")
    u0, u1 = hw0 >> 7, hw1 >> 7
    c0, c1 = ((0x01010100 | lf) + u0) & M32, (0x01010101 + u1) & M32
    keep0 = (u0 - 1) & M32
    keep1 = 0 if u0 else (u1 - 1) & M32
    sh0, sh1 = funnelshift_r(dsh0, dsh1, 8), dsh1 >> 8
    v0 = (dsh0 & keep0) | (sh0 & ~keep0 & M32)
    v1 = (dsh1 & keep1) | (sh1 & ~keep1 & M32)
    w = [byte_perm(v0, c0, 0x5140), byte_perm(v0, c0, 0x7362), byte_perm(v1, c1, 0x5140), byte_perm(v1, c1, 0x7362)]
    units = [(x >> s) & 0xFFFF for x in w for s in (0, 16)]
    return units[: 8 - nhole]


def emit_by_definition(dsh, nz, lf):
    """a head at position j closes the previous run: pair (delta[j-1], j - previous head)"""
    out, run = [], lf
    for j in range(8):
        if nz[j]:
            out.append(dsh[j] | (run << 8))
            run = 1
        else:
            run += 1
    return out


def test_branch_free_hole_deletion_all_positions():
    rng = random.Random(2)
    for hole in [None] + list(range(8)):
        for _ in range(300):
            dsh = [rng.randrange(256) for _ in range(8)]
            nz = [j != hole for j in range(8)]
            lf = rng.randrange(1, 247)          # 1 or 2 in the kernel; the identity holds for any count that stays below 256
            assert emit_common_path(dsh, nz, lf) == emit_by_definition(dsh, nz, lf), (hole, lf)


def expand_common_path(vals, cnts, qb):
    """phase C, common case (decompress_fast_kernel): 8 pairs with counts of 1, at most one count of 2.  Returns the
    codes (mod 256) of the 8 or 9 elements produced, starting from code qb."""
    va, vb = pack(vals[:4]), pack(vals[4:])
    ca, cb = pack(cnts[:4]), pack(cnts[4:])
    ta, tb = (ca - 0x01010101) & M32, (cb - 0x01010101) & M32
    assert ((ta | tb) & 0xFEFEFEFE) == 0 and bin(ta).count("1") + bin(tb).count("1") <= 1
    ntwo = bin(ta).count("1") + bin(tb).count("1")
    keep0 = ((ta << 8) - 1) & M32
    keep1 = 0 if ta else ((tb << 8) - 1) & M32
    up0, up1 = (va << 8) & M32, funnelshift_l(va, vb, 8)
    v0 = (va & keep0) | (up0 & ~keep0 & M32)
    v1 = (vb & keep1) | (up1 & ~keep1 & M32)
    v2 = vb >> 24
    def dp4a(a, b, acc):
        return (acc + sum(x * y for x, y in zip(unpack(a), unpack(b)))) & M32
    q4 = dp4a(v0, 0x01010101, qb)
    q8 = dp4a(v1, 0x01010101, q4)
    codes = [dp4a(v0, 0x00000001, qb), dp4a(v0, 0x00000101, qb), dp4a(v0, 0x00010101, qb), q4,
             dp4a(v1, 0x00000001, q4), dp4a(v1, 0x00000101, q4), dp4a(v1, 0x00010101, q4), q8, (q8 + v2) & M32]
    return [c & 0xFF for c in codes[: 8 + ntwo]]


def test_branch_free_duplication_all_positions():
    rng = random.Random(3)
    for two in [None] + list(range(8)):
        for _ in range(300):
            vals = [rng.randrange(256) for _ in range(8)]
            cnts = [2 if j == two else 1 for j in range(8)]
            qb = rng.randrange(256)
            want, q = [], qb
            for v, c in zip(vals, cnts):                      # run_length_decode + delta_decode, cache_engine.cpp:241-273
                for _ in range(c):
                    q = (q + v) & 0xFF
                    want.append(q)
            assert expand_common_path(vals, cnts, qb) == want, two


def test_count_classification_mask():
    """phase C: "every count is 1 or 2" == ((ca - 0x01010101) | (cb - 0x01010101)) & 0xfefefefe == 0, for all byte values
    (a count of 0 borrows into the next byte or out of the word; the borrow never hides a bad count)."""
    for c in itertools.product([0, 1, 2, 3, 255], repeat=4):
        ca = pack(list(c))
        ta = (ca - 0x01010101) & M32
        small = (ta & 0xFEFEFEFE) == 0
        assert small == all(x in (1, 2) for x in c), c


# ---- head-stationary emission of the long-run path (kv_codec_fast.cu, long_emit / HeadSelTable) ----------------------
def head_sel_table():
    """HeadSelTable: nibble i of entry m = position of the i-th set bit of m (0 beyond popc(m))."""
    tab = []
    for m in range(256):
        sel, n = 0, 0
        for j in range(8):
            if (m >> j) & 1:
                sel |= j << (4 * n)
                n += 1
        tab.append(sel)
    return tab


def long_emit_units(m, dsh, c_minus_prev):
    """The 16-bit units long_emit stages for a chunk with head mask m, shifted deltas dsh[0..7] and c - prev
    (distance from the chunk's first position back to the previous emitted head; <= 0 after a forced head)."""
    sel = head_sel_table()[m]
    stx, sty = pack(dsh[:4]), pack(dsh[4:])
    v0, v1 = byte_perm(stx, sty, sel & 0xFFFF), byte_perm(stx, sty, sel >> 16)
    p0, p1 = byte_perm(0x03020100, 0x07060504, sel & 0xFFFF), byte_perm(0x03020100, 0x07060504, sel >> 16)
    c0 = (p0 - ((p0 << 8) & M32) + (c_minus_prev & M32)) & M32
    c1 = (p1 - funnelshift_l(p0, p1, 8)) & M32
    w = [byte_perm(v0, c0, 0x5140), byte_perm(v0, c0, 0x7362), byte_perm(v1, c1, 0x5140), byte_perm(v1, c1, 0x7362)]
    units = []
    for x in w:
        units += [x & 0xFFFF, x >> 16]
    return units[:bin(m).count("1")]


def test_long_emit_selector_table_and_counts():
    """Every head mask, with and without a forced head in the leading stretch: values are the shifted deltas at the
    head positions, counts the distances between heads (the first one reaching back to the previous emitted head)."""
    rng = random.Random(5)
    for m in range(1, 256):
        pos = [j for j in range(8) if (m >> j) & 1]
        for trial in range(40):
            dsh = [rng.randrange(256) for _ in range(8)]
            if pos[0] > 0 and trial % 3 == 0:
                cmp_ = -rng.randrange(0, pos[0])             # a forced head at chunk position -cmp_ (< first natural head)
            else:
                cmp_ = rng.randrange(1, 256 - pos[0])         # (c + first head) - prev <= 255: no forced head in between
            want = [dsh[p] | (((p + cmp_) if i == 0 else (p - pos[i - 1])) << 8) for i, p in enumerate(pos)]
            assert long_emit_units(m, dsh, cmp_) == want, (m, cmp_)


def test_div255_and_forced_head_closed_forms():
    """div255 (umulhi by 0x80808081, >> 7) for the ranges the kernel feeds it, and the count of forced heads in a
    chunk's leading stretch: multiples of 255 in [d, d + lead) = div255(d + lead - 1) - div255(d - 1)."""
    def div255(x):
        return ((x * 0x80808081) >> 32) >> 7
    for x in list(range(0, 70000)) + [2 ** 31 - 1, 2 ** 32 - 1, 255 * 1234567, 255 * 1234567 - 1]:
        assert div255(x) == x // 255, x
    for d in range(1, 1200):
        for lead in range(0, 9):
            want = sum(1 for t in range(d, d + lead) if t % 255 == 0)
            assert div255(d + lead - 1) - div255(d - 1) == want, (d, lead)
