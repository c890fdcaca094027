"""Out-of-bounds check of the hot-path entry points without a sanitizer (compute-sanitizer is closed on the GPU pool).

Every buffer a kernel touches is carved out of ONE arena whose every float starts as a NaN with a marker payload.
Around each buffer sits a guard band that no call may change (a stray WRITE shows up as a changed guard word), and
because the bands are NaNs a stray READ that reaches an output poisons it (the outputs are compared, bit for bit, with
the results of the ordinary API on ordinary tensors).  Shapes are ragged on purpose: pixel counts that are not
multiples of the lookup's 32 / 64-pixel CTAs, widths that leave a partial 128-row stripe, buffers that start at
16-byte but not 128-byte boundaries."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MARK = 0x7FC0DEAD               # int32 view of a quiet NaN with payload 0xDEAD
GUARD = 260                     # floats between two buffers: > one 1 KB row, and NOT a multiple of 32 (keeps buffers off 128-byte boundaries)


@pytest.fixture(scope="module")
def sa():
    import stereoanywhere_b200 as sa_

    return sa_


class Arena:
    def __init__(self, floats):
        self.buf = torch.full((floats,), MARK, dtype=torch.int32, device=DEV)
        self.pos = GUARD
        self.used = []

    def carve(self, *shape, dtype=torch.float32, fill=None):
        n = int(np.prod(shape))
        per = 4 // torch.empty((), dtype=dtype).element_size()
        words = (n + per - 1) // per
        start = (self.pos + 3) // 4 * 4        # 16-byte aligned, nothing more
        t = self.buf[start:start + words].view(dtype)[:n].view(*shape)
        if fill is not None:
            t.copy_(fill)
        self.used.append((start, start + words))
        self.pos = start + words + GUARD
        assert self.pos < self.buf.numel(), "arena too small"
        return t

    def check(self, what):
        torch.cuda.synchronize()
        keep = torch.ones(self.buf.numel(), dtype=torch.bool, device=DEV)
        for s, e in self.used:
            keep[s:e] = False
        bad = (self.buf != MARK) & keep
        assert not bool(bad.any()), f"{what}: wrote outside its buffers at arena words {bad.nonzero()[:8].flatten().tolist()}"


def _inputs(b, c, h, w2, w3, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    fl = torch.randn(b, c, h, w2, device=DEV, generator=g)
    fr = torch.randn(b, c, h, w3, device=DEV, generator=g)
    nl = torch.nn.functional.normalize(torch.randn(b, 3, h, w2, device=DEV, generator=g), dim=1)
    nr = torch.nn.functional.normalize(torch.randn(b, 3, h, w3, device=DEV, generator=g), dim=1)
    x = torch.arange(w2, device=DEV, dtype=torch.float32).view(1, 1, 1, w2).expand(b, 1, h, w2)
    disp = torch.rand(b, 1, h, w2, device=DEV, generator=g) * (w3 / 3)
    # coordinates that also leave the row on both sides (lines of the zero padding, pixels without a line)
    cx = x - disp + (torch.rand(b, 1, h, w2, device=DEV, generator=g) - 0.5) * 1.5 * w3 * (torch.rand(b, 1, h, w2, device=DEV, generator=g) < 0.1)
    coords = torch.cat([cx, torch.zeros_like(cx)], 1).contiguous()
    conf = torch.rand(b, 1, h, w2, device=DEV, generator=g)
    return fl, fr, nl, nr, coords, disp.contiguous(), conf


@pytest.mark.parametrize("b,c,h,w2,w3", [(1, 32, 3, 40, 72), (2, 64, 2, 168, 168), (1, 32, 5, 312, 312), (1, 32, 1, 8, 8),
                                         (3, 32, 1, 44, 40)])
def test_hot_path_entry_points_stay_inside_their_buffers(sa, b, c, h, w2, w3):
    from stereoanywhere_b200 import _lib, ops

    lib = _lib.load()
    B = sa.CorrBlockB200
    fl, fr, nl, nr, coords, disp, conf = _inputs(b, c, h, w2, w3, seed=w2 + w3)
    st = torch.cuda.current_stream().cuda_stream
    rowf = ops.packed_row_floats(w3)
    rows = b * h * w2
    # the ordinary API on ordinary tensors: what every call below must reproduce bit for bit
    ref_s = B.from_features(fl, fr, truncate=(disp, conf, 0.9))
    ref_m = B.from_normals(nl, nr)
    want_s, want_m = B.lookup_pair(ref_s, ref_m, coords)
    old_mode = B.mono_mode
    try:
        B.mono_mode = "packed"
        ref_mp = B.from_normals(nl, nr)
    finally:
        B.mono_mode = old_mode
    want_s2, want_mp = B.lookup_pair(ref_s, ref_mp, coords)
    ref_h = B.from_features(fl, fr, truncate=(disp, conf, 0.9), storage="fp16")
    want_h, want_hm = B.lookup_pair(ref_h, ref_m, coords)
    g = torch.Generator(device=DEV).manual_seed(1)
    wgt = torch.randn(64, 36, device=DEV, generator=g) * 0.2
    bias = torch.randn(64, device=DEV, generator=g)
    want_cs, want_cm = sa.lookup_pair_convc1(ref_s, ref_m, coords, wgt.view(64, 36, 1, 1), bias)
    torch.cuda.synchronize()

    ar = Arena(2 * (fl.numel() + fr.numel()) + 3 * rows * rowf + 8 * b * 64 * h * w2 + 64 * GUARD + (1 << 16))
    a_fl, a_fr = ar.carve(*fl.shape, fill=fl), ar.carve(*fr.shape, fill=fr)
    a_nl, a_nr = ar.carve(*nl.shape, fill=nl), ar.carve(*nr.shape, fill=nr)
    a_co, a_di, a_cf = ar.carve(*coords.shape, fill=coords), ar.carve(*disp.shape, fill=disp), ar.carve(*conf.shape, fill=conf)
    a_w, a_b = ar.carve(64, 36, fill=wgt), ar.carve(64, fill=bias)
    packed = ar.carve(rows, rowf)
    packed_m = ar.carve(rows, rowf)
    packed_nr = ar.carve(b * 3 * h, rowf)
    packed_h = ar.carve(rows, rowf, dtype=torch.float16)   # 64-byte lines
    o_s, o_m = ar.carve(b, 36, h, w2), ar.carve(b, 36, h, w2)
    c_s, c_m = ar.carve(b, 64, h, w2), ar.carve(b, 64, h, w2)
    dv, dv3 = ops._divisor(c), ops._divisor(3)

    def ok(rc, what):
        assert rc == 0, f"{what}: {lib.sa_last_error().decode()}"
        ar.check(what)

    if ops.corr_packable(c, w2, w3):
        ok(lib.sa_corr_pack_tf32(a_fl.data_ptr(), a_fr.data_ptr(), b, c, h, w2, w3, dv, 1.0, a_di.data_ptr(), a_cf.data_ptr(), 0.9,
                                 packed.data_ptr(), st), "sa_corr_pack_tf32")
        assert torch.equal(packed, ref_s._packed)
        ok(lib.sa_corr_pack_tf32_half(a_fl.data_ptr(), a_fr.data_ptr(), b, c, h, w2, w3, dv, 1.0, a_di.data_ptr(), a_cf.data_ptr(),
                                      0.9, 1, packed_h.data_ptr(), st), "sa_corr_pack_tf32_half")
        assert torch.equal(packed_h.view(torch.int16), ref_h._packed_h.view(torch.int16).view(rows, -1))
    else:   # shapes the fused kernel does not take: the two-step construction wrote ref_s._packed
        packed.copy_(ref_s._ensure_packed())
        packed_h = None
    ok(lib.sa_pack_pyramid(a_nr.data_ptr(), b * 3 * h, w3, None, None, 0.0, 0, packed_nr.data_ptr(), st), "sa_pack_pyramid")
    assert torch.equal(packed_nr, ref_m._packed_nr)
    ok(lib.sa_pack_pyramid_normals(a_nl.data_ptr(), a_nr.data_ptr(), dv3, 1.73, b, h, w2, w3, packed_m.data_ptr(), st),
       "sa_pack_pyramid_normals")
    assert torch.equal(packed_m, ref_mp._packed)

    ok(lib.sa_lookup_packed_factored(packed.data_ptr(), packed_nr.data_ptr(), a_nl.data_ptr(), dv3, 1.73, w3, a_co.data_ptr(),
                                     a_co.stride(0), o_s.data_ptr(), o_m.data_ptr(), b, h, w2, st), "sa_lookup_packed_factored")
    assert torch.equal(o_s, want_s) and torch.equal(o_m, want_m)
    o_s.view(torch.int32).fill_(MARK); o_m.view(torch.int32).fill_(MARK)
    ok(lib.sa_lookup_packed(packed.data_ptr(), packed_m.data_ptr(), w3, a_co.data_ptr(), a_co.stride(0), o_s.data_ptr(),
                            o_m.data_ptr(), b, h, w2, st), "sa_lookup_packed (two volumes)")
    assert torch.equal(o_s, want_s2) and torch.equal(o_m, want_mp)
    o_s.view(torch.int32).fill_(MARK)
    ok(lib.sa_lookup_packed(packed.data_ptr(), None, w3, a_co.data_ptr(), a_co.stride(0), o_s.data_ptr(), None, b, h, w2, st),
       "sa_lookup_packed (one volume)")
    assert torch.equal(o_s, want_s)
    if packed_h is not None:
        o_s.view(torch.int32).fill_(MARK); o_m.view(torch.int32).fill_(MARK)
        ok(lib.sa_lookup_packed_half(packed_h.data_ptr(), 1, 2, packed_nr.data_ptr(), a_nl.data_ptr(), dv3, 1.73, w3, a_co.data_ptr(),
                                     a_co.stride(0), o_s.data_ptr(), o_m.data_ptr(), b, h, w2, st), "sa_lookup_packed_half")
        assert torch.equal(o_s, want_h) and torch.equal(o_m, want_hm)
    ok(lib.sa_lookup_factored_conv(packed.data_ptr(), packed_nr.data_ptr(), a_nl.data_ptr(), dv3, 1.73, w3, a_co.data_ptr(),
                                   a_co.stride(0), a_w.data_ptr(), a_b.data_ptr(), c_s.data_ptr(), c_m.data_ptr(), b, h, w2, st),
       "sa_lookup_factored_conv")
    assert torch.equal(c_s, want_cs) and torch.equal(c_m, want_cm)
    c_s.view(torch.int32).fill_(MARK); c_m.view(torch.int32).fill_(MARK)
    ok(lib.sa_lookup_packed_conv(packed.data_ptr(), packed_m.data_ptr(), w3, a_co.data_ptr(), a_co.stride(0), a_w.data_ptr(),
                                 a_b.data_ptr(), c_s.data_ptr(), c_m.data_ptr(), b, h, w2, st), "sa_lookup_packed_conv")
    want_ps, want_pm = sa.lookup_pair_convc1(ref_s, ref_mp, coords, wgt.view(64, 36, 1, 1), bias)
    assert torch.equal(c_s, want_ps) and torch.equal(c_m, want_pm)
    assert not torch.isnan(o_s).any() and not torch.isnan(c_s).any() and not torch.isnan(c_m).any()


class arena_alloc:
    """While active, every CUDA `torch.empty` / `empty_like` / `zeros` of a 2- or 4-byte dtype is carved from the arena:
    the wrappers of the package allocate their outputs (and the backward its accumulators) with exactly these calls."""

    def __init__(self, ar):
        self.ar = ar

    def __enter__(self):
        self.saved = (torch.empty, torch.empty_like, torch.zeros)
        e0, el0, z0 = self.saved
        ar = self.ar
        small = (torch.float32, torch.float16, torch.bfloat16, torch.int32)

        def norm(size):
            return tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)

        def empty(*size, dtype=None, device=None, **kw):
            dt = dtype or torch.float32
            if kw or device is None or torch.device(device).type != "cuda" or dt not in small:
                return e0(*size, dtype=dtype, device=device, **kw)
            return ar.carve(*norm(size), dtype=dt)

        def empty_like(t, **kw):
            if kw or not t.is_cuda or t.dtype not in small:
                return el0(t, **kw)
            return ar.carve(*t.shape, dtype=t.dtype)

        def zeros(*size, dtype=None, device=None, **kw):
            dt = dtype or torch.float32
            if kw or device is None or torch.device(device).type != "cuda" or dt not in small:
                return z0(*size, dtype=dtype, device=device, **kw)
            return ar.carve(*norm(size), dtype=dt).zero_()

        torch.empty, torch.empty_like, torch.zeros = empty, empty_like, zeros
        return self

    def __exit__(self, *exc):
        torch.empty, torch.empty_like, torch.zeros = self.saved


def _same(got, want, what, exact=True):
    got = got if isinstance(got, (tuple, list)) else [got]
    want = want if isinstance(want, (tuple, list)) else [want]
    assert len(got) == len(want), what
    for i, (g, w) in enumerate(zip(got, want)):
        if g is None and w is None:
            continue
        assert g.shape == w.shape and g.dtype == w.dtype, f"{what}[{i}]"
        if exact:
            assert torch.equal(g, w), f"{what}[{i}]: differs from the plain-tensor run (max |d| = {float((g.float() - w.float()).abs().max())})"
        else:
            assert torch.allclose(g, w, rtol=1e-5, atol=1e-5), f"{what}[{i}]"


def test_every_wrapper_allocation_is_respected(sa):
    """The whole public surface, outputs allocated from the guarded arena (general kernels, ragged shapes, backward)."""
    from stereoanywhere_b200 import producers

    B = sa.CorrBlockB200
    b, c, h, w2, w3 = 2, 40, 3, 44, 36          # C % 32 != 0: SIMT correlation; W3 % 8 != 0: general pyramid / lookup
    fl, fr, nl, nr, coords, disp, conf = _inputs(b, c, h, w2, w3, seed=3)
    fl32, fr32, *_ = _inputs(b, 64, h, w2, w2, seed=4)
    g = torch.Generator(device=DEV).manual_seed(2)
    mde_l, mde_r = torch.rand(b, 1, h, w2, device=DEV, generator=g), torch.rand(b, 1, h, w3, device=DEV, generator=g)
    mde_r[0, 0, 0, :4] = 1.0
    mde_full = torch.rand(b, 1, 4 * h + 3, 4 * w2 + 2, device=DEV, generator=g)
    bin_mask = (torch.rand(b, 1, h, w2, 1, device=DEV, generator=g) < 0.3).float()
    noise = torch.rand(b, 1, h, w2, 1, device=DEV, generator=g)
    agg = torch.randn(b, 1, h, w2, w3, device=DEV, generator=g) * 3

    def every_op():
        out = {}
        out["corr_simt"] = B.corr(fl, fr)
        out["corr_tf32"] = B.corr(fl32, fr32)
        out["mono_corr"] = B.mono_corr(nl, nr)
        vol = out["corr_simt"]
        blk = B(vol, num_levels=3, radius=3, pad=(2, 1), truncate=(disp, conf, 0.8))       # general kernels
        out["lookup_general"] = blk(coords)
        out["levels"] = [lv.contiguous() for lv in blk.corr_pyramid]
        blk2 = B(vol)                                                                      # W3 = 36: levels layout, radius 4
        out["lookup_r4"] = blk2(coords)
        out["pair_levels"] = B.lookup_pair(blk2, B(out["mono_corr"]), coords)
        out["trunc_mask"] = sa.truncation_mask(disp, conf, 0.9)
        out["trunc_prod"] = sa.truncation_mask(disp, conf, 0.9, vol.view(b, 1, h, w2, w3))
        out["masked"] = sa.masked_volume(agg, mde_l, mde_r, 8)
        out["masked_mono"] = sa.masked_mono_volume(nl, nr, mde_l, mde_r, 8)
        v5 = agg
        out["roll"] = sa.corrupt_volume(v5, bin_mask, "roll", shift=5)
        out["noise"] = sa.corrupt_volume(v5, bin_mask, "noise", noise=noise)
        out["gauss"] = sa.corrupt_volume(v5, bin_mask, "gauss", gauss_k=0.7)
        out["disp"] = sa.estimate_disparities(agg)
        out["conf"] = sa.estimate_confidences(agg)
        out["mono_inputs"] = producers.mono_inputs(mde_full)
        out["lsq"] = producers.weighted_lsq_b200(torch.cat([mde_l, mde_l], 1), torch.cat([disp, disp * 0.9], 1),
                                                 torch.cat([conf, conf], 1))
        return out

    def training():
        f2 = fl32.clone().requires_grad_(True)
        f3 = fr32.clone().requires_grad_(True)
        blk = B(B.corr(f2, f3), truncate=(disp, conf, 0.9))
        x = torch.arange(w2, device=DEV, dtype=torch.float32).view(1, 1, 1, w2).expand(b, 1, h, w2)
        cc = torch.cat([x - disp, torch.zeros_like(x)], 1)
        loss = (blk(cc) * torch.linspace(-1, 1, 36, device=DEV).view(1, 36, 1, 1)).sum() + blk(cc + 0.3).square().sum()
        loss.backward()
        return f2.grad, f3.grad

    want = every_op()
    want_g = training()
    torch.cuda.synchronize()
    ar = Arena(48 << 20)
    with arena_alloc(ar):
        got = every_op()
        ar.check("forward entry points")
        got_g = training()
        ar.check("training step (lookup / pyramid / corr backward)")
    assert len(ar.used) > 40, "the wrappers did not allocate from the arena"
    for k in want:
        _same(got[k], want[k], k)
    _same(got_g, want_g, "gradients", exact=False)   # atomics in the lookup backward: summation order


def test_the_arena_notices_a_stray_write():
    ar = Arena(4096)
    t = ar.carve(10)
    ar.check("nothing written yet")
    s, e = ar.used[0]
    ar.buf[e] = 0          # one word past the buffer
    with pytest.raises(AssertionError, match="wrote outside"):
        ar.check("deliberate overrun")
    assert t.numel() == 10


def test_full_size_results_repeat_bit_for_bit(sa):
    """No sanitizer on the pool for race checks either: at KITTI size (7 488 lookup CTAs, 2 256 stripes, every SM busy) the
    fused constructor, the dual lookups and the fused lookup + convc1 must return the same bits on every run."""
    import bench

    B = sa.CorrBlockB200
    b, c, h, w = bench.WORKLOADS["c2_kitti_375x1242_b8"]
    _, d = bench.make_inputs(b, c, h, w, torch.device(DEV), seed=0)
    first = B.from_features(d["fl"], d["fr"], truncate=(d["tdisp"], d["tconf"], 0.9))
    mono = B.from_normals(d["nl"], d["nr"])
    for _ in range(3):
        again = B.from_features(d["fl"], d["fr"], truncate=(d["tdisp"], d["tconf"], 0.9))
        assert torch.equal(again._packed, first._packed)
        del again
    half = B.from_features(d["fl"], d["fr"], truncate=(d["tdisp"], d["tconf"], 0.9), storage="fp16")
    g = torch.Generator(device=DEV).manual_seed(1)
    wgt, bias = torch.randn(64, 36, 1, 1, device=DEV, generator=g) * 0.2, torch.randn(64, device=DEV, generator=g)
    coords = d["coords0"]
    s0, m0 = B.lookup_pair(first, mono, coords)
    h0, hm0 = B.lookup_pair(half, mono, coords)
    c0, cm0 = sa.lookup_pair_convc1(first, mono, coords, wgt, bias)
    for _ in range(10):
        s, m = B.lookup_pair(first, mono, coords)
        assert torch.equal(s, s0) and torch.equal(m, m0)
        hs, hm = B.lookup_pair(half, mono, coords)
        assert torch.equal(hs, h0) and torch.equal(hm, hm0)
        cs, cm = sa.lookup_pair_convc1(first, mono, coords, wgt, bias)
        assert torch.equal(cs, c0) and torch.equal(cm, cm0)


def test_lookup_store_policy_does_not_change_results(sa):
    """The dual lookup's TMA output stores carry an L2 evict_first hint only while a launch's output is below ~0.6 of the
    L2 (csrc/packed.cu, launch_packed_tt).  Ten KITTI-size pairs are above that bound, slices of five pairs below it:
    same bits either way."""
    B = sa.CorrBlockB200
    b, c, h, w = 10, 32, 96, 312
    fl, fr, nl, nr, coords, disp, conf = _inputs(b, c, h, w, w, seed=11)
    assert 2 * 36 * 4 * h * w * b * 10 > torch.cuda.get_device_properties(0).L2_cache_size * 6   # the no-hint branch
    s, m = B.lookup_pair(B.from_features(fl, fr, truncate=(disp, conf, 0.9)), B.from_normals(nl, nr), coords)
    for lo, hi in ((0, 5), (5, 10)):
        sl = slice(lo, hi)
        assert 2 * 36 * 4 * h * w * (hi - lo) * 10 <= torch.cuda.get_device_properties(0).L2_cache_size * 6
        s2, m2 = B.lookup_pair(B.from_features(fl[sl], fr[sl], truncate=(disp[sl], conf[sl], 0.9)), B.from_normals(nl[sl], nr[sl]),
                               coords[sl].contiguous())
        assert torch.equal(s[sl], s2) and torch.equal(m[sl], m2)
