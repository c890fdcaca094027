"""The line-packed pyramid layout of csrc/packed.cu, restated in numpy and checked on the CPU.

One 128-byte line per (volume row, block q of 8 level-0 columns) must be enough for every lookup with
floor(x) in [8q, 8q+8): slots 0..16 hold L0[8q-4 .. 8q+12], slots 17..31 the five border entries of levels 1..3,
and the remaining window entries are re-derived with the pyramid's own 0.5 (a + b) (`line_levels`), then shifted
to the window start and blended (`blend_windows`).  This test builds the lines from the oracle's float32 pyramid,
performs the lookup from ONE line exactly as the kernel does, and compares with the oracle's closed-form lookup
over the levels (reference corr.py:93-115) - equal up to the float32 rounding of the blend, borders and far-away
coordinates included.  It pins the layout algebra without a GPU; the CUDA kernels are compared with the same oracle in
tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

from oracle import corr_oracle as O

Q_MIN = -5  # first block index (csrc/packed.cu: kQMin)


def packed_blocks(w):
    return w // 8 + 9


def slot_map(s):
    """slot -> (level, offset): the entry in slot s of block q is L_level[(q << (3 - level)) + offset]."""
    if s < 17:
        return 0, s - 4
    t = s - 17
    level = 1 + t // 5
    r = t % 5
    hi = {1: 6, 2: 4, 3: 3}[level]
    return level, (r - 4 if r < 2 else hi + (r - 2))


def pack_row(levels):
    """levels: four 1-D float32 arrays of one volume row -> [nblk, 32] lines, zeros outside the levels."""
    w = levels[0].shape[0]
    lines = np.zeros((packed_blocks(w), 32), dtype=np.float32)
    for bi in range(lines.shape[0]):
        q = bi + Q_MIN
        for s in range(32):
            lv, off = slot_map(s)
            idx = q * (8 >> lv) + off
            if 0 <= idx < levels[lv].shape[0]:
                lines[bi, s] = levels[lv][idx]
    return lines


def lookup_from_line(ln, x):
    """`line_levels` + `blend_windows` of csrc/packed.cu for one pixel, in float32."""
    f32 = np.float32
    half = f32(0.5)
    l0 = ln[:17].copy()
    l1 = np.zeros(13, f32); l2 = np.zeros(11, f32); l3 = np.zeros(10, f32)
    l1[[0, 1, 10, 11, 12]] = ln[17:22]
    for t in range(8):
        l1[2 + t] = (l0[2 * t] + l0[2 * t + 1]) * half
    l2[[0, 1, 8, 9, 10]] = ln[22:27]
    for t in range(6):
        l2[2 + t] = (l1[2 * t] + l1[2 * t + 1]) * half
    l3[[0, 1, 7, 8, 9]] = ln[27:32]
    for t in range(5):
        l3[2 + t] = (l2[2 * t] + l2[2 * t + 1]) * half
    x = f32(x)
    x0 = int(np.floor(x))
    out = np.zeros(36, f32)

    def blend(a, b, f):  # sa_common.cuh: fma(f, b, (1 - f) * a) - the product rounded on its own, the sum once
        prod = f32(f32(1.0) - f) * a
        return f32(np.float64(f) * np.float64(b) + np.float64(prod))

    for lv, (win, scale) in enumerate(((l0, 1.0), (l1, 0.5), (l2, 0.25), (l3, 0.125))):
        start = (x0 >> lv) & ((8 >> lv) - 1)  # window start inside the stored range: [0,8), [0,4), [0,2), 0
        xs = f32(x * f32(scale))
        f = f32(xs - np.floor(xs))
        for k in range(9):
            out[9 * lv + k] = blend(win[start + k], win[start + k + 1], f)
    return out


@pytest.mark.parametrize("w", [8, 40, 128, 312])
def test_one_line_serves_every_lookup_in_its_block(w):
    rng = np.random.RandomState(100 + w)
    row = rng.randn(w).astype(np.float32)
    levels = O.closed_pyramid(row, 4)               # float32 pooling = the reference's, bit for bit
    assert [lv.shape[0] for lv in levels] == [w, w // 2, w // 4, w // 8]
    lines = pack_row(levels)
    xs = np.concatenate([rng.uniform(-45, w + 40, 400), np.arange(-41, w + 34, dtype=np.float64),
                         rng.uniform(0, w, 200)]).astype(np.float32)
    n = xs.shape[0]
    ref = O.closed_lookup([np.broadcast_to(lv, (1, 1, n, lv.shape[0])) for lv in levels],
                          xs.astype(np.float64).reshape(1, 1, n), 4)[0, :, 0, :]      # [36, n]
    for i, x in enumerate(xs):
        q = (int(np.floor(x)) >> 3) - Q_MIN
        got = lookup_from_line(lines[q], x) if 0 <= q < lines.shape[0] else np.zeros(36, np.float32)
        assert np.abs(got - ref[:, i]).max() <= 2e-6 * max(1.0, np.abs(ref[:, i]).max()), (w, float(x))


def test_blocks_cover_every_coordinate_with_a_live_tap():
    """Blocks q = -5 .. W/8+3: outside them every tap of every level is outside the image (all zeros)."""
    w = 64
    levels = [np.ones(w >> i, np.float32) for i in range(4)]
    for x in (-39.9, -33.0, w + 31.9):      # just inside: the widest window (level 3) still reaches a pixel
        ref = O.closed_lookup([lv.reshape(1, 1, 1, -1) for lv in levels], np.array([[[x]]]), 4)
        assert np.abs(ref).max() > 0
        assert 0 <= (int(np.floor(x)) >> 3) - Q_MIN < packed_blocks(w)
    for x in (-40.0, -100.0, w + 32.0, w + 500.0):
        ref = O.closed_lookup([lv.reshape(1, 1, 1, -1) for lv in levels], np.array([[[x]]]), 4)
        assert np.abs(ref).max() == 0


@pytest.mark.parametrize("w", [40, 312])
def test_factored_line_is_the_combination_of_the_right_normal_lines(w):
    """mono_mode "factored" (csrc/packed.cu, FV path): the packed line of the rank-3 mono volume row
    V[w3] = k * sum_c nL[c] nR[c, w3] is, slot by slot, the combination of the packed lines of the three right-normal
    rows with k * nL[c] as coefficients (one product + two FMAs per slot, as in the kernel)."""
    f32 = np.float32
    rng = np.random.RandomState(7 + w)
    nr = rng.randn(3, w); nr = (nr / np.linalg.norm(nr, axis=0)).astype(f32)
    nl = rng.randn(3); nl = (nl / np.linalg.norm(nl)).astype(f32)
    k = f32(f32(1.73) * f32(1.0 / np.float64(f32(np.sqrt(3.0)))))
    lines_c = [pack_row(O.closed_pyramid(nr[c], 4)) for c in range(3)]       # the 14 MB array of the kernel, one row
    a = [f32(nl[c] * k) for c in range(3)]
    comb = np.float64(a[0] * lines_c[0])                                     # FMUL
    comb = f32(np.float64(a[1]) * np.float64(lines_c[1]) + comb)             # FFMA
    comb = f32(np.float64(a[2]) * np.float64(lines_c[2]) + np.float64(comb)) # FFMA
    vol_row = (1.73 * (nl.astype(np.float64) @ nr.astype(np.float64)) / np.float64(f32(np.sqrt(3.0))))
    levels = O.closed_pyramid(vol_row, 4)                                    # float64 closed form of the volume row
    xs = np.concatenate([rng.uniform(-45, w + 40, 300), np.arange(-41, w + 34, dtype=np.float64)]).astype(f32)
    n = xs.shape[0]
    ref = O.closed_lookup([np.broadcast_to(lv, (1, 1, n, lv.shape[0])) for lv in levels],
                          xs.astype(np.float64).reshape(1, 1, n), 4)[0, :, 0, :]
    for i, x in enumerate(xs):
        q = (int(np.floor(x)) >> 3) - Q_MIN
        got = lookup_from_line(comb[q], x) if 0 <= q < comb.shape[0] else np.zeros(36, f32)
        assert np.abs(got - ref[:, i]).max() <= 2e-6, (w, float(x))
