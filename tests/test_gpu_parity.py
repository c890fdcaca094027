"""Parity of the CUDA path (through torch.ops.sa_b200 -> C ABI) against the oracle and the golden
fixtures.  Everything here needs a B200; nothing reads /root/reference.

Tolerances (north_star): correlation normwise max|d| / max|ref| <= 1e-3 for TF32 and <= 2e-6 for the
fp32 SIMT kernel; lookup / pyramid / masks compare fp32 arithmetic and must agree to <= 2e-5
absolute on O(1) data (grid_sample's normalise/un-normalise round trip alone is ~1e-5, SURVEY
Appendix A); pyramid levels are bit-exact (halving is exact).
"""
import math

import numpy as np
import pytest
import torch

from oracle import corr_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
WIDTH_TAGS = ["w24", "w39", "w40", "w50x34"]
COORD_TAGS = ["left", "right", "far", "int"]


def G(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(scope="module")
def sa():
    import stereoanywhere_b200 as sa_

    return sa_


@pytest.fixture(autouse=True)
def packed_mono_mode(sa):
    """The bit-for-bit comparisons of `from_normals` in this module are statements about mono_mode "packed"; the
    default "factored" mode has its own tests (`test_mono_lookup_factored`, the random-shape sweep)."""
    B = sa.CorrBlockB200
    old, B.mono_mode = B.mono_mode, "packed"
    yield
    B.mono_mode = old


def normwise(got, ref):
    got = got.detach().double().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got, dtype=np.float64)
    ref = ref.detach().double().cpu().numpy() if isinstance(ref, torch.Tensor) else np.asarray(ref, dtype=np.float64)
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30)


def maxabs(got, ref):
    got = got.detach().double().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got, dtype=np.float64)
    ref = ref.detach().double().cpu().numpy() if isinstance(ref, torch.Tensor) else np.asarray(ref, dtype=np.float64)
    return np.abs(got - ref).max()


def make_coords(b, h, w, gen, kind="left"):
    x = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w).expand(b, 1, h, w)
    y = torch.arange(h, dtype=torch.float32).view(1, 1, h, 1).expand(b, 1, h, w)
    if kind == "left":
        dx = -torch.rand(b, 1, h, w, generator=gen) * (w / 4)
    else:
        dx = torch.rand(b, 1, h, w, generator=gen) * 8
    return torch.cat([x + dx, y], dim=1).contiguous()


# ------------------------------------------------------------------------------------------ golden

def test_golden_corr_fp32(sa, golden_path):
    g = golden_path
    sa.CorrBlockB200.precision = "fp32"
    try:
        v = sa.CorrBlockB200.corr(G(g["a1_fl"]), G(g["a1_fr"]))
    finally:
        sa.CorrBlockB200.precision = "tf32"
    assert v.shape == g["a1_vol"].shape and v.dtype == torch.float32
    assert normwise(v, g["a1_vol"]) < 2e-6
    m = sa.CorrBlockB200.mono_corr(G(g["a2_nl"]), G(g["a2_nr"]))
    assert normwise(m, g["a2_vol"]) < 2e-6
    # protocol path for the mono volume: 1.73 * corr(normals) as the model writes it
    m2 = 1.73 * sa.CorrBlockB200.corr(G(g["a2_nl"]), G(g["a2_nr"]))
    assert normwise(m2, g["a2_vol"]) < 2e-6


@pytest.mark.parametrize("tag", WIDTH_TAGS)
def test_golden_pyramid_bit_exact(sa, golden_path, tag):
    g = golden_path
    blk = sa.CorrBlockB200(G(g[f"a3_{tag}_vol"]), num_levels=4, radius=4)
    pyr = blk.corr_pyramid
    assert len(pyr) == 4
    for i, p in enumerate(pyr):
        want = g[f"a3_{tag}_p{i}"]
        assert tuple(p.shape) == want.shape
        assert np.array_equal(p.cpu().numpy(), want), f"level {i}"


@pytest.mark.parametrize("tag", WIDTH_TAGS)
@pytest.mark.parametrize("ctag", COORD_TAGS)
def test_golden_lookup(sa, golden_path, tag, ctag):
    g = golden_path
    blk = sa.CorrBlockB200(G(g[f"a3_{tag}_vol"]), num_levels=4, radius=4)
    out = blk(G(g[f"a4_{tag}_{ctag}_coords"]))
    want = g[f"a4_{tag}_{ctag}_out"]
    assert tuple(out.shape) == want.shape and out.is_contiguous()
    tol = 2e-4 if ctag == "far" else 3e-5  # 'far': |x| ~ 3W makes grid_sample's own coordinate noise larger
    assert maxabs(out, want) < tol
    # and against the float64 closed form, which has no grid_sample noise
    levels = O.closed_pyramid(g[f"a3_{tag}_vol"][:, :, :, 0], 4)
    closed = O.closed_lookup(levels, g[f"a4_{tag}_{ctag}_coords"][:, 0], radius=4)
    assert maxabs(out, closed) < (2e-4 if ctag == "far" else 5e-6)


def test_golden_lookup_radius_levels_pad(sa, golden_path):
    g = golden_path
    vol, coords = G(g["a4_alt_vol"]), G(g["a4_alt_coords"])
    assert maxabs(sa.CorrBlockB200(vol, num_levels=2, radius=3)(coords), g["a4_alt_r3l2"]) < 3e-5
    assert maxabs(sa.CorrBlockB200(vol, num_levels=3, radius=2)(coords), g["a4_alt_r2l3"]) < 3e-5
    out = sa.CorrBlockB200(vol, num_levels=4, radius=4, pad=[2, 3])(coords)
    assert tuple(out.shape) == (1, 36, 2, 27)
    assert maxabs(out, g["a4_alt_pad23"]) < 3e-5


def test_golden_truncation(sa, golden_path):
    g = golden_path
    d, c = G(g["a5_disp"]), G(g["a5_conf"])
    assert maxabs(sa.truncation_mask(d, c, 0.9), g["a5_mask"]) < 1e-6
    vol = G(g["a1_vol"])
    prod = sa.truncation_mask(d, c, 0.9, vol=vol.squeeze(3).unsqueeze(1))
    assert maxabs(prod.squeeze(1).unsqueeze(3), g["a5_product"]) < 2e-6
    # fused: block built with truncate= has T*V as level 0 and pools it
    blk = sa.CorrBlockB200(vol, num_levels=4, radius=4, truncate=(d, c, 0.9))
    assert maxabs(blk.fullcorr, g["a5_product"]) < 2e-6
    ref = O.closed_pyramid(g["a5_product"][:, :, :, 0], 4)
    for lv, want in zip(blk.corr_pyramid, ref):
        assert maxabs(lv.view(want.shape), want) < 2e-6


def test_golden_masked_volume(sa, golden_path):
    g = golden_path
    mv = G(g["a2_vol"]).squeeze(3).unsqueeze(1)
    out = sa.masked_volume(mv, G(g["a6_mde_l"]), G(g["a6_mde_r"]), 8)
    assert tuple(out.shape) == g["a6_masked"].shape
    assert np.array_equal(out.cpu().numpy(), g["a6_masked"])  # masks are 0/1: exact
    fused = sa.masked_mono_volume(G(g["a2_nl"]), G(g["a2_nr"]), G(g["a6_mde_l"]), G(g["a6_mde_r"]), 8)
    assert maxabs(fused, g["a6_masked"]) < 2e-6
    assert np.array_equal((fused != 0).cpu().numpy(), g["a6_masked"] != 0)


def test_golden_corruption(sa, golden_path):
    g = golden_path
    v5 = G(g["a1_vol"]).squeeze(3).unsqueeze(1)
    m = G(g["a7_binmask"].astype(np.float32))
    assert np.array_equal(sa.corrupt_volume(v5, m, "roll", shift=5).cpu().numpy(), g["a7_roll5"])
    noise = G(g["a7_noise"].astype(np.float32))
    assert maxabs(sa.corrupt_volume(v5, m, "noise", noise=noise), g["a7_noised"]) < 1e-6
    k = float(v5.max())
    assert maxabs(sa.corrupt_volume(v5, m, "gauss", gauss_k=k), g["a7_gaussed"]) < 1e-5


def test_golden_model_slice(sa, golden_model):
    """Tensors captured at the hot-path call sites of a real reference forward."""
    g = golden_model
    sa.CorrBlockB200.precision = "fp32"
    try:
        sv32 = sa.CorrBlockB200.corr(G(g["stereo_fl"]), G(g["stereo_fr"]))
    finally:
        sa.CorrBlockB200.precision = "tf32"
    assert normwise(sv32, g["stereo_vol"]) < 2e-6
    sv = sa.CorrBlockB200.corr(G(g["stereo_fl"]), G(g["stereo_fr"]))  # tensor-core path
    assert normwise(sv, g["stereo_vol"]) < 1e-3
    mv = 1.73 * sa.CorrBlockB200.corr(G(g["mono_nl"]), G(g["mono_nr"]))
    assert normwise(mv, g["mono_vol"]) < 2e-6
    blk_s = sa.CorrBlockB200(G(g["stereo_ctor"]), radius=4, num_levels=4)
    blk_m = sa.CorrBlockB200(G(g["mono_ctor"]), radius=4, num_levels=4)
    for it in range(int(g["iters"])):
        c = G(g[f"it{it}_coords"])
        s_want, m_want = g[f"it{it}_stereo"], g[f"it{it}_mono"]
        assert maxabs(blk_s(c), s_want) < 3e-5 * max(1.0, np.abs(s_want).max())
        assert maxabs(blk_m(c), m_want) < 3e-5 * max(1.0, np.abs(m_want).max())
        pa, pb = sa.CorrBlockB200.lookup_pair(blk_s, blk_m, c)
        assert torch.equal(pa, blk_s(c)) and torch.equal(pb, blk_m(c))


# ------------------------------------------------------------------------------------------ oracle, seeded

@pytest.mark.parametrize("shape", [(1, 64, 24, 128), (2, 256, 5, 312), (1, 128, 3, 168), (1, 32, 2, 40)])
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_corr_vs_oracle(sa, shape, precision):
    b, c, h, w = shape
    gen = torch.Generator().manual_seed(b * 1000 + c + w)
    fl = torch.randn(b, c, h, w, generator=gen)
    fr = torch.randn(b, c, h, w, generator=gen)
    ref = O.aten_corr_volume(fl, fr)
    sa.CorrBlockB200.precision = precision
    try:
        got = sa.CorrBlockB200.corr(fl.to(DEV), fr.to(DEV))
    finally:
        sa.CorrBlockB200.precision = "tf32"
    assert got.shape == ref.shape
    err = normwise(got, ref)
    assert err < (1e-3 if precision == "tf32" else 2e-6), err


def test_corr_rectangular_and_odd_fp32(sa):
    gen = torch.Generator().manual_seed(5)
    fl = torch.randn(2, 19, 3, 37, generator=gen)
    fr = torch.randn(2, 19, 3, 53, generator=gen)
    ref = O.aten_corr_volume(fl, fr)
    got = sa.CorrBlockB200.corr(fl.to(DEV), fr.to(DEV))  # C % 8 != 0 -> SIMT kernel
    assert got.shape == (2, 3, 37, 1, 53)
    assert normwise(got, ref) < 2e-6


@pytest.mark.parametrize("b,h,w", [(1, 96, 128), (2, 16, 312), (1, 8, 168), (1, 4, 240), (1, 3, 100), (1, 2, 77)])
def test_block_vs_oracle_seeded(sa, b, h, w):
    """Constructor + call against the ATen restatement, model-like and awkward widths."""
    gen = torch.Generator().manual_seed(w)
    vol = torch.randn(b, h, w, 1, w, generator=gen)
    ref_blk = O.OracleCorrBlock(vol, num_levels=4, radius=4)
    blk = sa.CorrBlockB200(vol.to(DEV), num_levels=4, radius=4)
    for i, p in enumerate(blk.corr_pyramid):
        assert np.array_equal(p.cpu().numpy(), ref_blk.corr_pyramid[i].numpy())
    levels64 = O.closed_pyramid(vol[:, :, :, 0].numpy(), 4)
    for kind in ("left", "right"):
        coords = make_coords(b, h, w, gen, kind)
        want = ref_blk(coords)
        got = blk(coords.to(DEV))
        assert got.shape == want.shape
        # float64 closed form on the same fp32 coords: only our own fp32 blend error remains
        assert maxabs(got, O.closed_lookup(levels64, coords[:, 0].numpy(), radius=4)) < 5e-6
        # ATen reference: grid_sample's normalise / un-normalise round trip perturbs the sample
        # position by ~4 eps W/2 px; on N(0,1) rows (tap-to-tap differences up to ~8) that is
        # ~2e-6 W absolute (6e-4 at W=312) of *reference* noise
        assert maxabs(got, want) < 2e-6 * max(w, 16)


@pytest.mark.parametrize("b,h,w", [(1, 5, 24), (2, 7, 312), (1, 3, 128), (1, 2, 768)])
def test_packed_layout_is_bit_identical_to_levels(sa, b, h, w):
    """The line-packed fast path (csrc/packed.cu) against the plain pyramid + lookup kernels."""
    gen = torch.Generator().manual_seed(100 + w)
    vol = torch.randn(b, h, w, 1, w, generator=gen).to(DEV)
    disp = (torch.rand(b, 1, h, w, generator=gen) * (w / 4)).to(DEV)
    conf = torch.rand(b, 1, h, w, generator=gen).to(DEV)
    B = sa.CorrBlockB200
    for trunc in (None, (disp, conf, 0.9)):
        old = B.layout
        try:
            B.layout = "levels"
            ref = B(vol, num_levels=4, radius=4, truncate=trunc)
            B.layout = "packed"
            blk = B(vol, num_levels=4, radius=4, truncate=trunc)
        finally:
            B.layout = old
        assert blk._packed is not None and ref._packed is None
        assert torch.equal(blk.fullcorr, ref.fullcorr)
        for lp, lr in zip(blk.corr_pyramid, ref.corr_pyramid):
            assert torch.equal(lp, lr)
        x = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w).expand(b, 1, h, w)
        for dx in (-torch.rand(b, 1, h, w, generator=gen) * (w / 4), torch.rand(b, 1, h, w, generator=gen) * 60,
                   (torch.rand(b, 1, h, w, generator=gen) - 0.5) * 4 * w, -torch.randint(0, 9, (b, 1, h, w), generator=gen).float()):
            coords = torch.cat([x + dx, torch.zeros(b, 1, h, w)], 1).to(DEV)
            assert torch.equal(blk(coords), ref(coords))
            pa, pb = B.lookup_pair(blk, blk, coords)
            assert torch.equal(pa, ref(coords)) and torch.equal(pb, pa)


def test_from_normals_equals_mono_corr_block(sa, golden_path):
    """A2 + pyramid fused (`from_normals`) against the two-step path, bit for bit, and the golden volume."""
    gen = torch.Generator().manual_seed(77)
    B = sa.CorrBlockB200
    for (b, h, w) in [(2, 5, 312), (1, 3, 24)]:
        nl = torch.nn.functional.normalize(torch.randn(b, 3, h, w, generator=gen), dim=1).to(DEV)
        nr = torch.nn.functional.normalize(torch.randn(b, 3, h, w, generator=gen), dim=1).to(DEV)
        fused = B.from_normals(nl, nr)
        two = B(B.mono_corr(nl, nr))
        x = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w).expand(b, 1, h, w)
        coords = torch.cat([x - torch.rand(b, 1, h, w, generator=gen) * (w / 4), torch.zeros(b, 1, h, w)], 1).to(DEV)
        assert torch.equal(fused(coords), two(coords))
        assert torch.equal(fused._ensure_packed(), two._packed)
        assert torch.equal(fused.fullcorr, two.fullcorr)
    g = golden_path
    blk = B.from_normals(G(g["a2_nl"]), G(g["a2_nr"]))
    assert normwise(blk.fullcorr, g["a2_vol"]) < 2e-6


def test_lookup_pair_equals_two_calls(sa):
    gen = torch.Generator().manual_seed(11)
    b, h, w = 2, 12, 312
    va = torch.randn(b, h, w, 1, w, generator=gen).to(DEV)
    vb = torch.randn(b, h, w, 1, w, generator=gen).to(DEV)
    ba, bb = sa.CorrBlockB200(va), sa.CorrBlockB200(vb)
    coords = make_coords(b, h, w, gen).to(DEV)
    oa, ob = sa.CorrBlockB200.lookup_pair(ba, bb, coords)
    assert torch.equal(oa, ba(coords)) and torch.equal(ob, bb(coords))


def test_edge_cases(sa):
    gen = torch.Generator().manual_seed(3)
    # single pixel row / tiny widths / W not multiple of 4 -> generic kernels
    for (b, h, w1, w3) in [(1, 1, 1, 16), (1, 1, 9, 9), (3, 2, 17, 8)]:
        vol = torch.randn(b, h, w1, 1, w3, generator=gen)
        ref = O.OracleCorrBlock(vol, num_levels=3, radius=2)
        blk = sa.CorrBlockB200(vol.to(DEV), num_levels=3, radius=2)
        x = (torch.rand(b, 1, h, w1, generator=gen) * (w3 + 8) - 4)
        coords = torch.cat([x, torch.zeros_like(x)], 1)
        assert maxabs(blk(coords.to(DEV)), ref(coords)) < 3e-5
    # coordinates far outside / non-finite magnitude: every tap is zero
    vol = torch.randn(1, 2, 8, 1, 8, generator=gen).to(DEV)
    blk = sa.CorrBlockB200(vol, num_levels=2, radius=4)
    far = torch.full((1, 2, 2, 8), 1e9, device=DEV)
    assert float(blk(far).abs().max()) == 0.0
    assert float(blk(-far).abs().max()) == 0.0
    # fp16 coords come back as fp16 (reference casts to coords.dtype, corr.py:115)
    c16 = torch.zeros(1, 2, 2, 8, device=DEV, dtype=torch.float16)
    assert blk(c16).dtype == torch.float16
    # non-contiguous volume is accepted
    v2 = torch.randn(1, 2, 8, 8, 1, generator=gen).to(DEV).permute(0, 1, 2, 4, 3)
    assert sa.CorrBlockB200(v2).fullcorr.shape == (1, 2, 8, 1, 8)
    # errors surface as exceptions
    with pytest.raises(ValueError):
        sa.CorrBlockB200(torch.zeros(2, 8, 8, device=DEV))
    with pytest.raises((ValueError, RuntimeError)):
        blk(torch.zeros(1, 2, 3, 8, device=DEV))  # wrong H


# ------------------------------------------------------------------------------------------ full size, properties

def sampled_closed_lookup(vol_rows, xs, num_levels=4, radius=4):
    """float64 closed form of the lookup for sampled volume rows `[n, W]` at coordinates `xs [n]`:
    `[n, num_levels * (2 radius + 1)]`."""
    lv = vol_rows.double()
    xs = xs.double()
    outs = []
    for i in range(num_levels):
        w_i = lv.shape[1]
        padded = torch.nn.functional.pad(lv, (radius + 2, radius + 2))      # zeros outside the row
        x = xs / (2 ** i)
        x0 = torch.floor(x)
        f = (x - x0).unsqueeze(1)
        k = torch.arange(-radius, radius + 1, device=lv.device, dtype=torch.float64).unsqueeze(0)
        j = (x0.unsqueeze(1) + k).clamp(-radius - 2, w_i + radius).long() + radius + 2
        a = torch.gather(padded, 1, j.clamp(0, padded.shape[1] - 1))
        b_ = torch.gather(padded, 1, (j + 1).clamp(0, padded.shape[1] - 1))
        inside_a = ((x0.unsqueeze(1) + k) >= 0) & ((x0.unsqueeze(1) + k) < w_i)
        inside_b = ((x0.unsqueeze(1) + k + 1) >= 0) & ((x0.unsqueeze(1) + k + 1) < w_i)
        outs.append((1 - f) * a * inside_a + f * b_ * inside_b)
        lv = 0.5 * (lv[:, 0:2 * (w_i // 2):2] + lv[:, 1:2 * (w_i // 2):2])
    return torch.cat(outs, 1)


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4_tile", "c5_w768", "c5_c128_w768"])
def test_full_size_properties(sa, cfg):
    """BASELINE configs at full size through size-independent properties (the oracle would take
    minutes here): (1) integer coords + tap k=0 at level 0 gathers the volume itself; (2) the
    lookup is linear in the volume; (3) a constant volume looks up to a constant inside the
    image; (4) pyramid levels are exact means of level 0; (5) TF32 corr matches an fp64 dot
    product on sampled entries."""
    b, c, h, w = {"c1": (1, 256, 96, 128), "c2": (8, 256, 96, 312), "c3": (8, 256, 136, 240),
                  "c4_tile": (1, 256, 280, 168), "c5_w768": (1, 256, 96, 768), "c5_c128_w768": (1, 128, 96, 768)}[cfg]
    gen = torch.Generator(device=DEV).manual_seed(0)
    fl = torch.randn(b, c, h, w, device=DEV, generator=gen)
    fr = torch.randn(b, c, h, w, device=DEV, generator=gen)
    vol = sa.CorrBlockB200.corr(fl, fr)
    assert vol.shape == (b, h, w, 1, w)
    # (5) sampled entries against float64
    idx = torch.randint(0, b * h * w * w, (4096,), device=DEV, generator=gen)
    bb = idx // (h * w * w)
    hh = (idx // (w * w)) % h
    w2 = (idx // w) % w
    w3 = idx % w
    ref = (fl[bb, :, hh, w2].double() * fr[bb, :, hh, w3].double()).sum(1) / float(np.float32(np.sqrt(c)))
    got = vol.view(-1)[idx].double()
    assert float((got - ref).abs().max() / vol.abs().max()) < 1e-3
    blk = sa.CorrBlockB200(vol, num_levels=4, radius=4)
    # (4) pyramid
    v4 = vol.view(b * h * w, w)
    l1 = 0.5 * (v4[:, 0::2] + v4[:, 1::2])
    assert torch.equal(blk.corr_pyramid[1].view(b * h * w, -1), l1)
    l2 = 0.5 * (l1[:, 0::2] + l1[:, 1::2])
    assert torch.equal(blk.corr_pyramid[2].view(b * h * w, -1), l2)
    l3 = 0.5 * (l2[:, 0 : 2 * (l2.shape[1] // 2) : 2] + l2[:, 1 : 2 * (l2.shape[1] // 2) : 2])
    assert torch.equal(blk.corr_pyramid[3].view(b * h * w, -1), l3)
    # (1) gather
    tgt = torch.randint(0, w, (b, 1, h, w), device=DEV, generator=gen).float()
    coords = torch.cat([tgt, torch.zeros_like(tgt)], 1)
    out = blk(coords)
    want = torch.gather(vol.view(b, h, w, w), 3, tgt.view(b, h, w, 1).long()).view(b, h, w)
    assert torch.equal(out[:, 4], want)
    # (2) linearity
    coords = torch.cat([tgt - torch.rand(b, 1, h, w, device=DEV, generator=gen) * 40, torch.zeros_like(tgt)], 1)
    o1 = blk(coords)
    # (7) sampled pixels against the float64 closed form of the lookup (borders included: x - U(0, 40) leaves the row)
    pix = torch.randint(0, b * h * w, (4096,), device=DEV, generator=gen)
    want_px = sampled_closed_lookup(vol.view(b * h * w, w)[pix], coords[:, 0].reshape(-1)[pix])
    got_px = o1.permute(0, 2, 3, 1).reshape(b * h * w, 36)[pix].double()
    assert float((got_px - want_px).abs().max()) < 5e-6 * max(1.0, float(vol.abs().max()))
    blk2 = sa.CorrBlockB200(vol * 2.0 + 1.0, num_levels=4, radius=4)
    ones = sa.CorrBlockB200(torch.ones_like(vol), num_levels=4, radius=4)
    o2 = blk2(coords)
    assert float((o2 - (2.0 * o1 + ones(coords))).abs().max()) < 1e-4
    # (3) constant volume: taps fully inside the row read exactly 1
    xin = torch.rand(b, 1, h, w, device=DEV, generator=gen) * (w - 80) + 40
    oc = ones(torch.cat([xin, torch.zeros_like(xin)], 1))
    assert float((oc[:, :9] - 1.0).abs().max()) < 1e-6
    # (6) the fused constructors at full size: same packed arrays as the two-step forms, bit for bit
    del blk2, ones, o2, oc
    tdisp = torch.rand(b, 1, h, w, device=DEV, generator=gen) * (w / 4)
    tconf = torch.rand(b, 1, h, w, device=DEV, generator=gen)
    two = sa.CorrBlockB200(vol, truncate=(tdisp, tconf, 0.9))
    one = sa.CorrBlockB200.from_features(fl, fr, truncate=(tdisp, tconf, 0.9))
    assert torch.equal(one._packed, two._packed)
    del one, two
    nl = torch.nn.functional.normalize(torch.randn(b, 3, h, w, device=DEV, generator=gen), dim=1)
    nr = torch.nn.functional.normalize(torch.randn(b, 3, h, w, device=DEV, generator=gen), dim=1)
    assert torch.equal(sa.CorrBlockB200.from_normals(nl, nr)._ensure_packed(), sa.CorrBlockB200(sa.CorrBlockB200.mono_corr(nl, nr))._packed)
    # (8) the benchmarked pair (fused stereo block + factored mono block, one launch) on sampled pixels vs float64
    fs = sa.CorrBlockB200.from_features(fl, fr, truncate=(tdisp, tconf, 0.9))
    fm = sa.CorrBlockB200.from_normals(nl, nr)
    s_, m_ = sa.CorrBlockB200.lookup_pair(fs, fm, coords)
    xs = coords[:, 0].reshape(-1)[pix]
    mono_rows = sa.CorrBlockB200.mono_corr(nl, nr).view(b * h * w, w)[pix]
    got_m = m_.permute(0, 2, 3, 1).reshape(b * h * w, 36)[pix].double()
    assert float((got_m - sampled_closed_lookup(mono_rows, xs)).abs().max()) < 5e-6
    t_rows = sa.truncation_mask(tdisp, tconf, 0.9, vol=vol.squeeze(3).unsqueeze(1)).view(b * h * w, w)[pix]
    got_s = s_.permute(0, 2, 3, 1).reshape(b * h * w, 36)[pix].double()
    assert float((got_s - sampled_closed_lookup(t_rows, xs)).abs().max()) < 5e-6 * max(1.0, float(vol.abs().max()))


def test_lookup_fused_with_convc1(sa):
    """relu(convc1(lookup)) for both volumes in one kernel vs lookup + fp32 torch conv."""
    torch.backends.cudnn.allow_tf32 = False
    gen = torch.Generator().manual_seed(21)
    B = sa.CorrBlockB200
    for (b, h, w) in [(1, 4, 128), (2, 9, 312), (1, 3, 40)]:
        va = torch.randn(b, h, w, 1, w, generator=gen).to(DEV)
        vb = torch.randn(b, h, w, 1, w, generator=gen).to(DEV)
        weight = (torch.randn(64, 36, 1, 1, generator=gen) / 6).to(DEV)
        bias = (torch.randn(64, generator=gen) * 0.1).to(DEV)
        x = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w).expand(b, 1, h, w)
        coords = torch.cat([x - torch.rand(b, 1, h, w, generator=gen) * (w / 4), torch.zeros(b, 1, h, w)], 1).to(DEV)
        ba, bb = B(va), B(vb)
        fa, fb = sa.lookup_pair_convc1(ba, bb, coords, weight, bias)
        la, lb = B.lookup_pair(ba, bb, coords)
        ra = torch.relu(torch.nn.functional.conv2d(la.double(), weight.double(), bias.double()))
        rb = torch.relu(torch.nn.functional.conv2d(lb.double(), weight.double(), bias.double()))
        assert fa.shape == (b, 64, h, w) and fa.dtype == torch.float32
        assert normwise(fa, ra) < 1e-3 and normwise(fb, rb) < 1e-3, (normwise(fa, ra), normwise(fb, rb))
        assert float(fa.min()) >= 0.0  # ReLU


@pytest.mark.parametrize("shape", [(1, 64, 3, 128, 128), (2, 256, 5, 312, 312), (1, 32, 2, 40, 40), (1, 128, 2, 168, 168),
                                   (1, 64, 1, 240, 240), (1, 64, 1, 520, 776), (1, 32, 2, 8, 8), (1, 64, 2, 132, 264)])
@pytest.mark.parametrize("trunc", [False, True])
def test_corr_pack_fused_matches_two_step(sa, shape, trunc):
    """A1 (+A5) + A3 in the GEMM epilogue (sa_corr_pack_tf32) vs sa_corr_tf32 -> sa_pack_pyramid: same packed
    array bit for bit, hence identical lookups; and vs the fp64 closed form within the TF32 tolerance."""
    b, c, h, w2, w3 = shape
    gen = torch.Generator().manual_seed(31 + w3)
    fl = torch.randn(b, c, h, w2, generator=gen).to(DEV)
    fr = torch.randn(b, c, h, w3, generator=gen).to(DEV)
    t = None
    if trunc:
        t = ((torch.rand(b, 1, h, w2, generator=gen) * (w3 / 4)).to(DEV), torch.rand(b, 1, h, w2, generator=gen).to(DEV), 0.9)
    B = sa.CorrBlockB200
    B.precision = "tf32"
    two = B(B.corr(fl, fr), truncate=t)
    one = B.from_features(fl, fr, truncate=t)
    assert one._packed is not None and two._packed is not None
    assert one._packed.shape == two._packed.shape
    assert torch.equal(one._packed, two._packed), float((one._packed - two._packed).abs().max())
    x = torch.arange(w2, dtype=torch.float32).view(1, 1, 1, w2).expand(b, 1, h, w2)
    coords = torch.cat([x - torch.rand(b, 1, h, w2, generator=gen) * (w3 / 4), torch.zeros(b, 1, h, w2)], 1).to(DEV)
    assert torch.equal(one(coords), two(coords))
    # on-demand volume of the fused block == the two-step block's
    assert torch.equal(one.fullcorr, two.fullcorr)
    # against the fp64 einsum (TF32 tolerance, normwise)
    ref = torch.einsum("bchw,bchv->bhwv", fl.double(), fr.double()) / math.sqrt(c)
    if trunc:
        w2i = torch.arange(w2, device=DEV, dtype=torch.float64).view(1, 1, w2, 1)
        w3i = torch.arange(w3, device=DEV, dtype=torch.float64).view(1, 1, 1, w3)
        d, cf = t[0].double().squeeze(1).unsqueeze(-1), t[1].double().squeeze(1).unsqueeze(-1)
        ref = ref * ((1 - cf) + cf * (torch.sigmoid((w2i - d) - w3i) * (1 - 0.9) + 0.9))
    assert normwise(one.fullcorr.squeeze(3), ref) < 1e-3


# ------------------------------------------------------------------------------------------ 8f-2 reductions
@pytest.mark.parametrize("tag", ["sq24", "r40x50", "flat33"])
def test_golden_volume_reductions(sa, golden_reductions, tag):
    """Soft-argmax disparities / entropy confidences vs the reference's own outputs (fixtures)."""
    g = golden_reductions
    v = G(g[f"{tag}_vol"])
    dl, dr = sa.estimate_disparities(v)
    cl, cr = sa.estimate_confidences(v)
    assert dl.shape == g[f"{tag}_dl"].shape and dr.shape == g[f"{tag}_dr"].shape
    assert maxabs(dl, g[f"{tag}_dl"]) < 2e-4 and maxabs(dr, g[f"{tag}_dr"]) < 2e-4   # pixels
    assert maxabs(cl, g[f"{tag}_cl"]) < 2e-5 and maxabs(cr, g[f"{tag}_cr"]) < 2e-5
    if tag == "sq24":
        assert maxabs(sa.estimate_left_disparity(v, vol_pad=[2, 3]), g["sq24_dl_pad"]) < 2e-4
        assert maxabs(sa.estimate_right_disparity(v, vol_pad=[2, 3]), g["sq24_dr_pad"]) < 2e-4
        assert torch.equal(sa.estimate_left_confidence(v), cl) and torch.equal(sa.estimate_right_confidence(v), cr)


@pytest.mark.parametrize("shape,gain", [((2, 1, 5, 312, 312), 4.0), ((1, 1, 3, 240, 240), 1.0), ((1, 1, 2, 130, 37), 8.0),
                                        ((1, 1, 2, 37, 130), 8.0), ((1, 1, 1, 768, 768), 2.0), ((1, 1, 2, 5, 3), 1.0)])
def test_volume_reductions_vs_oracle(sa, shape, gain):
    """Seeded volumes at model-like and awkward shapes vs the oracle (ATen fp32 op sequence and fp64 closed form)."""
    gen = torch.Generator().manual_seed(77 + shape[3])
    v = torch.randn(*shape, generator=gen) * gain
    dl, dr = sa.estimate_disparities(v.to(DEV))
    cl, cr = sa.estimate_confidences(v.to(DEV))
    w = max(shape[3], shape[4])
    assert maxabs(dl, O.aten_estimate_left_disparity(v)) < 1e-6 * w * 2 and maxabs(dr, O.aten_estimate_right_disparity(v)) < 1e-6 * w * 2
    assert maxabs(cl, O.aten_estimate_left_confidence(v)) < 3e-5 and maxabs(cr, O.aten_estimate_right_confidence(v)) < 3e-5
    rdl, rdr, rcl, rcr = O.closed_volume_reductions(v.numpy())
    assert maxabs(dl, rdl) < 1e-6 * w * 2 and maxabs(dr, rdr) < 1e-6 * w * 2
    assert maxabs(cl, rcl) < 3e-5 and maxabs(cr, rcr) < 3e-5


def test_volume_reductions_properties_full_size(sa):
    """c2-size volume: a one-hot-like volume recovers its own disparity; a constant volume gives the uniform
    expectation and zero confidence (size-independent properties)."""
    b, h, w = 2, 96, 312
    gen = torch.Generator(device=DEV).manual_seed(5)
    d = torch.randint(0, 40, (b, h, w), device=DEV, generator=gen)
    x = torch.arange(w, device=DEV).view(1, 1, w)
    tgt = (x - d).clamp(min=0)                                   # matching column of every left pixel
    vol = torch.full((b, 1, h, w, w), -40.0, device=DEV)
    vol[:, 0].scatter_(3, tgt.unsqueeze(3), 40.0)
    dl, _ = sa.estimate_disparities(vol)
    assert float((dl[:, 0] - (x - tgt).float()).abs().max()) < 1e-3
    cl, _ = sa.estimate_confidences(vol)
    assert float(cl.min()) > 0.999
    flat = torch.zeros(1, 1, 4, w, w, device=DEV)
    dl, dr = sa.estimate_disparities(flat)
    assert float((dl[0, 0, 0] - (torch.arange(w, device=DEV) - (w - 1) / 2)).abs().max()) < 1e-3
    assert float((dr[0, 0, 0] - ((w - 1) / 2 - torch.arange(w, device=DEV))).abs().max()) < 1e-3
    cl, cr = sa.estimate_confidences(flat)
    assert float(cl.abs().max()) < 2e-3 and float(cr.abs().max()) < 2e-3  # log2(1/W + 1e-6) is not exactly -log2 W


@pytest.mark.parametrize("shape", [(2, 5, 312, 312), (1, 3, 128, 128), (1, 2, 40, 40), (1, 2, 132, 264), (1, 1, 520, 776), (1, 2, 8, 8)])
def test_pack_normals_matches_two_step(sa, shape):
    """The mono packer (pack_kernel<normals>, behind from_normals; right normals register-resident over a
    contiguous chunk of rows) vs mono_corr + pack: identical bit for bit (C = 3 FMA order shared with sa_corr_fp32)."""
    b, h, w2, w3 = shape
    gen = torch.Generator().manual_seed(41 + w3)
    nl = torch.nn.functional.normalize(torch.randn(b, 3, h, w2, generator=gen), dim=1).to(DEV)
    nr = torch.nn.functional.normalize(torch.randn(b, 3, h, w3, generator=gen), dim=1).to(DEV)
    B = sa.CorrBlockB200
    fused = B.from_normals(nl, nr)
    two = B(B.mono_corr(nl, nr))
    assert torch.equal(fused._ensure_packed(), two._packed)
    x = torch.arange(w2, dtype=torch.float32).view(1, 1, 1, w2).expand(b, 1, h, w2)
    coords = torch.cat([x - torch.rand(b, 1, h, w2, generator=gen) * (w3 / 4), torch.zeros(b, 1, h, w2)], 1).to(DEV)
    assert torch.equal(fused(coords), two(coords))


# ------------------------------------------------------------------------------------------ 8f-4 backward
@pytest.mark.parametrize("tag", ["w24", "w39", "w40"])
@pytest.mark.parametrize("layout", ["packed", "levels"])
def test_golden_gradients(sa, golden_grads, tag, layout):
    """Gradient w.r.t. the volume of three lookups (and through the truncation product) vs the gradients the
    reference's own autograd produced (fixtures)."""
    g = golden_grads
    B = sa.CorrBlockB200
    old = B.layout
    B.layout = layout
    try:
        coords = [G(g[f"{tag}_coords{k}"]) for k in range(3)]
        wts = [G(g[f"{tag}_w{k}"]) for k in range(3)]
        v = G(g[f"{tag}_vol"]).requires_grad_(True)
        blk = B(v, num_levels=4, radius=4)
        assert (blk._packed is not None) == (layout == "packed" and tag != "w39")
        loss = sum((blk(c) * w).sum() for c, w in zip(coords, wts))
        (dv,) = torch.autograd.grad(loss, v)
        assert dv.shape == v.shape
        assert maxabs(dv, g[f"{tag}_dvol"]) < 2e-5
        # fused truncation: gradient w.r.t. V of the block built from T * V (T detached)
        v2 = G(g[f"{tag}_vol"]).requires_grad_(True)
        blk2 = B(v2, num_levels=4, radius=4, truncate=(G(g[f"{tag}_tdisp"]), G(g[f"{tag}_tconf"]), 0.9))
        loss2 = sum((blk2(c) * w).sum() for c, w in zip(coords, wts))
        (dv2,) = torch.autograd.grad(loss2, v2)
        assert maxabs(dv2, g[f"{tag}_dvol_trunc"]) < 2e-5
        # strict protocol: the caller forms the product itself, autograd handles the multiply
        v3 = G(g[f"{tag}_vol"]).requires_grad_(True)
        t = sa.truncation_mask(G(g[f"{tag}_tdisp"]), G(g[f"{tag}_tconf"]), 0.9)
        blk3 = B((t * v3.squeeze(3).unsqueeze(1)).squeeze(1).unsqueeze(3), num_levels=4, radius=4)
        (dv3,) = torch.autograd.grad(sum((blk3(c) * w).sum() for c, w in zip(coords, wts)), v3)
        assert maxabs(dv3, g[f"{tag}_dvol_trunc"]) < 2e-5
        # forward values under grad mode are those of the no-grad path
        with torch.no_grad():
            ref = B(v.detach())(coords[0])
        assert torch.equal(blk(coords[0]).detach(), ref)
    finally:
        B.layout = old


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_golden_corr_gradients(sa, golden_grads, prec):
    g = golden_grads
    B = sa.CorrBlockB200
    old = B.precision
    B.precision = prec
    try:
        fl, fr = G(g["c_fl"]).requires_grad_(True), G(g["c_fr"]).requires_grad_(True)
        vol = B.corr(fl, fr)
        dfl, dfr = torch.autograd.grad((vol * G(g["c_w"])).sum(), (fl, fr))
        # "fp32": two fp32 library GEMMs; "tf32": sa_corr_backward_tf32 on the tensor cores (north_star: 1e-3 for TF32)
        tol = 2e-6 if prec == "fp32" else 1e-3
        assert normwise(dfl, g["c_dfl"]) < tol and normwise(dfr, g["c_dfr"]) < tol
    finally:
        B.precision = old


@pytest.mark.parametrize("b,c,h,w2,w3", [(1, 256, 3, 312, 312), (2, 128, 2, 168, 168), (1, 64, 2, 40, 72), (1, 32, 1, 8, 8),
                                         (1, 256, 1, 768, 768), (1, 160, 2, 240, 236), (1, 96, 2, 132, 264)])
def test_corr_backward_tensor_core_vs_fp64(sa, b, c, h, w2, w3):
    """`sa_corr_backward_tf32` (both products of the adjoint of corr(), K-major / MN-major operand paths, K not a
    multiple of the 32-column stage, channel counts that do not fill the 128-lane tile) against float64."""
    from stereoanywhere_b200 import ops

    gen = torch.Generator().manual_seed(c + w2 + w3)
    fl, fr = torch.randn(b, c, h, w2, generator=gen), torch.randn(b, c, h, w3, generator=gen)
    gv = torch.randn(b, h, w2, 1, w3, generator=gen)
    dl, dr = ops.corr_backward(gv.to(DEV), fl.to(DEV), fr.to(DEV), 1.0, True, True)
    g64 = gv.squeeze(3).double() / float(np.float32(np.sqrt(c)))
    want_l = torch.einsum("bhwv,bchv->bchw", g64, fr.double())
    want_r = torch.einsum("bhwv,bchw->bchv", g64, fl.double())
    assert normwise(dl, want_l) < 1e-3 and normwise(dr, want_r) < 1e-3, (normwise(dl, want_l), normwise(dr, want_r))
    only_l, none_r = ops.corr_backward(gv.to(DEV), fl.to(DEV), fr.to(DEV), 1.0, True, False)
    assert none_r is None and torch.equal(only_l, dl)
    # through autograd: corr() in tf32 precision now differentiates on the tensor cores
    f2, f3 = fl.to(DEV).requires_grad_(True), fr.to(DEV).requires_grad_(True)
    if c % 32 == 0:
        vol = sa.CorrBlockB200.corr(f2, f3)
        a2, a3 = torch.autograd.grad((vol * gv.to(DEV)).sum(), (f2, f3))
        assert torch.equal(a2, dl) and torch.equal(a3, dr)


def test_training_step_gradients_vs_oracle(sa):
    """A miniature training step - corr(), truncation, both blocks, four GRU-like iterations with detached
    coords - differentiated through the CUDA path and through the oracle's ATen op sequence on the CPU."""
    gen = torch.Generator().manual_seed(99)
    b, c, h, w = 2, 64, 4, 72
    fl, fr = torch.randn(b, c, h, w, generator=gen), torch.randn(b, c, h, w, generator=gen)
    mono = torch.randn(b, h, w, 1, w, generator=gen)   # stands in for the hourglass output (stereoanywhere.py:210)
    tdisp, tconf = torch.rand(b, 1, h, w, generator=gen) * (w / 4), torch.rand(b, 1, h, w, generator=gen)
    x = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w).repeat(b, 1, h, 1)
    coords = [torch.cat([x - torch.rand(b, 1, h, w, generator=gen) * (w / 4), torch.zeros(b, 1, h, w)], 1) for _ in range(4)]
    ws = [torch.randn(b, 36, h, w, generator=gen) for _ in range(4)]
    wm = [torch.randn(b, 36, h, w, generator=gen) for _ in range(4)]

    def run(block_cls, corr_fn, tmask_fn, dev):
        f2, f3 = fl.to(dev).requires_grad_(True), fr.to(dev).requires_grad_(True)
        mv = mono.to(dev).requires_grad_(True)
        vol = corr_fn(f2, f3)
        t = tmask_fn(tdisp.to(dev), tconf.to(dev))
        sfn = block_cls((t * vol.squeeze(3).unsqueeze(1)).squeeze(1).unsqueeze(3), num_levels=4, radius=4)
        mfn = block_cls(mv, num_levels=4, radius=4)
        loss = 0
        for k in range(4):
            cd = coords[k].to(dev).detach()
            loss = loss + (sfn(cd) * ws[k].to(dev)).sum() + (mfn(cd) * wm[k].to(dev)).sum()
        return torch.autograd.grad(loss, (f2, f3, mv))

    sa.CorrBlockB200.precision = "fp32"
    try:
        got = run(sa.CorrBlockB200, sa.CorrBlockB200.corr, lambda d, cf: sa.truncation_mask(d, cf, 0.9), DEV)
    finally:
        sa.CorrBlockB200.precision = "tf32"
    want = run(O.OracleCorrBlock, O.aten_corr_volume, lambda d, cf: O.aten_truncation_mask(d, cf, 0.9), "cpu")
    for a_, b_ in zip(got, want):
        assert normwise(a_, b_) < 1e-5


@pytest.mark.parametrize("shape", [(2, 3, 72, 72), (1, 2, 40, 520), (1, 1, 132, 264), (1, 2, 20, 52), (1, 2, 17, 30)])
def test_volume_passes_vs_oracle(sa, shape):
    """A2 / A5 / A6 / A7 standalone passes at widths that take the register-resident row kernels (W3 <= 384), the
    streaming row kernels (W3 > 384) and the generic kernels (W3 % 4 != 0), against the oracle's ATen sequence."""
    b, h, w2, w3 = shape
    gen = torch.Generator().manual_seed(11 + w3)
    nl = torch.nn.functional.normalize(torch.randn(b, 3, h, w2, generator=gen), dim=1)
    nr = torch.nn.functional.normalize(torch.randn(b, 3, h, w3, generator=gen), dim=1)
    vol = torch.randn(b, 1, h, w2, w3, generator=gen)
    # A2
    mono = sa.CorrBlockB200.mono_corr(nl.to(DEV), nr.to(DEV))
    ref = 1.73 * (torch.einsum("bchw,bchv->bhwv", nl.double(), nr.double()) / float(np.float32(np.sqrt(3.0))))
    assert mono.shape == (b, h, w2, 1, w3) and maxabs(mono.squeeze(3), ref) < 2e-6
    # A5 (square volumes only: the reference's mask is [B,1,H,W,W])
    if w2 == w3:
        disp = torch.rand(b, 1, h, w2, generator=gen) * (w3 / 4)
        conf = torch.rand(b, 1, h, w2, generator=gen)
        t_ref = O.aten_truncation_mask(disp, conf, 0.9)
        assert maxabs(sa.truncation_mask(disp.to(DEV), conf.to(DEV), 0.9), t_ref) < 1e-6
        assert maxabs(sa.truncation_mask(disp.to(DEV), conf.to(DEV), 0.9, vol=vol.to(DEV)), t_ref * vol) < 1e-5
    # A6
    mde_l, mde_r = torch.rand(b, 1, h, w2, generator=gen), torch.rand(b, 1, h, w3, generator=gen)
    mde_l[0, 0, 0, 0] = 1.0
    want = O.aten_masked_volume(vol, O.aten_depth_bin_masks(mde_l, 8), O.aten_depth_bin_masks(mde_r, 8))
    got = sa.masked_volume(vol.to(DEV), mde_l.to(DEV), mde_r.to(DEV), 8)
    assert np.array_equal(got.cpu().numpy(), want.numpy())
    fused = sa.masked_mono_volume(nl.to(DEV), nr.to(DEV), mde_l.to(DEV), mde_r.to(DEV), 8)
    want_m = O.aten_masked_volume(mono.cpu().squeeze(3).unsqueeze(1), O.aten_depth_bin_masks(mde_l, 8), O.aten_depth_bin_masks(mde_r, 8))
    assert np.array_equal(fused.cpu().numpy(), want_m.numpy())
    # A7
    binm = (mde_l < 0.3).to(torch.float32)
    assert np.array_equal(sa.corrupt_volume(vol.to(DEV), binm.to(DEV), "roll", shift=7).cpu().numpy(),
                          O.aten_corrupt_roll(vol, binm, 7).numpy())
    noise = torch.rand(b, 1, h, w2, generator=gen)
    assert maxabs(sa.corrupt_volume(vol.to(DEV), binm.to(DEV), "noise", noise=noise.to(DEV)),
                  O.aten_corrupt_scale(vol, binm, noise.unsqueeze(4))) < 1e-6


def test_fused_constructors_random_shapes(sa):
    """Thirty random shapes (ragged W2 / W3, single rows, widths around the tile / chunk boundaries): the fused
    constructors must reproduce the two-step forms bit for bit, and a lookup from them the fp64 closed form."""
    rng = np.random.RandomState(2024)
    B = sa.CorrBlockB200
    B.precision = "tf32"
    for it in range(30):
        b, h = int(rng.randint(1, 3)), int(rng.randint(1, 5))
        w2 = int(rng.choice([4, 8, 28, 32, 36, 124, 128, 132, 160, 252, 256, 260, 312, 388]))
        w3 = int(rng.choice([8, 16, 24, 32, 40, 120, 128, 136, 160, 248, 256, 264, 312, 384, 392, 520]))
        c = int(rng.choice([32, 64, 96, 128]))
        gen = torch.Generator().manual_seed(1000 + it)
        fl = torch.randn(b, c, h, w2, generator=gen).to(DEV)
        fr = torch.randn(b, c, h, w3, generator=gen).to(DEV)
        t = None
        if it % 2:
            t = ((torch.rand(b, 1, h, w2, generator=gen) * (w3 / 4)).to(DEV), torch.rand(b, 1, h, w2, generator=gen).to(DEV), 0.9)
        one, two = B.from_features(fl, fr, truncate=t), B(B.corr(fl, fr), truncate=t)
        assert torch.equal(one._packed, two._packed), (it, b, c, h, w2, w3)
        nl = torch.nn.functional.normalize(torch.randn(b, 3, h, w2, generator=gen), dim=1).to(DEV)
        nr = torch.nn.functional.normalize(torch.randn(b, 3, h, w3, generator=gen), dim=1).to(DEV)
        mono = B.from_normals(nl, nr)
        assert torch.equal(mono._ensure_packed(), B(B.mono_corr(nl, nr))._packed), (it, b, h, w2, w3)
        x = torch.arange(w2, dtype=torch.float32).view(1, 1, 1, w2).expand(b, 1, h, w2)
        coords = torch.cat([x * (w3 / max(w2, 1)) - torch.rand(b, 1, h, w2, generator=gen) * (w3 / 4) + (it % 3) * 3,
                            torch.zeros(b, 1, h, w2)], 1).to(DEV)
        s_, m_ = B.lookup_pair(one, mono, coords)
        levels = O.closed_pyramid(mono.fullcorr.squeeze(3).cpu().numpy(), 4)[:4]
        ref = O.closed_lookup([lv.reshape(b, h, w2, -1) for lv in levels], coords[:, 0].cpu().numpy(), 4)
        assert maxabs(m_, ref) < 1e-5, (it, b, h, w2, w3)
        assert torch.equal(s_, two(coords))
        B.mono_mode = "factored"   # (restored by the module's autouse fixture)
        s_f, m_f = B.lookup_pair(one, B.from_normals(nl, nr), coords)
        B.mono_mode = "packed"
        assert maxabs(m_f, ref) < 1e-5 and maxabs(m_f, m_) <= 1e-6 and torch.equal(s_f, s_), (it, b, h, w2, w3)


@pytest.mark.parametrize("shape", [(2, 5, 312, 312), (1, 3, 128, 128), (1, 2, 40, 40), (1, 2, 132, 264), (1, 1, 388, 520), (1, 2, 8, 8)])
def test_mono_lookup_on_the_fly_is_bit_identical(sa, shape):
    """`from_normals` blocks serve their lookups from the normal maps inside the lookup kernel (no packed mono
    array): same bits as the packed path, alone and paired with a packed stereo block, borders included."""
    b, h, w2, w3 = shape
    gen = torch.Generator().manual_seed(51 + w3)
    nl = torch.nn.functional.normalize(torch.randn(b, 3, h, w2, generator=gen), dim=1).to(DEV)
    nr = torch.nn.functional.normalize(torch.randn(b, 3, h, w3, generator=gen), dim=1).to(DEV)
    B = sa.CorrBlockB200
    old = B.mono_mode
    try:
        B.mono_mode = "otf"
        otf = B.from_normals(nl, nr)
        assert otf._packed is None and otf._otf
        B.mono_mode = "packed"
        pk = B.from_normals(nl, nr)
        assert pk._packed is not None
    finally:
        B.mono_mode = old
    stereo = B(torch.randn(b, h, w2, 1, w3, generator=gen).to(DEV))
    x = torch.arange(w2, dtype=torch.float32).view(1, 1, 1, w2).expand(b, 1, h, w2) * (w3 / w2)
    for kind, dx in (("left", -torch.rand(b, 1, h, w2, generator=gen) * (w3 / 3)), ("right", torch.rand(b, 1, h, w2, generator=gen) * 60),
                     ("far", (torch.rand(b, 1, h, w2, generator=gen) - 0.5) * 4 * w3), ("int", -torch.randint(0, 9, (b, 1, h, w2), generator=gen).float())):
        coords = torch.cat([x + dx, torch.zeros(b, 1, h, w2)], 1).to(DEV)
        want = pk(coords)
        assert torch.equal(otf(coords), want), kind
        s1, m1 = B.lookup_pair(stereo, otf, coords)
        s2, m2 = B.lookup_pair(stereo, pk, coords)
        assert torch.equal(m1, want) and torch.equal(m2, want) and torch.equal(s1, s2), kind
    # consumers that need the packed array get it on demand; the reference attributes still work
    assert torch.equal(otf.fullcorr, pk.fullcorr)
    w = (torch.randn(64, 36, 1, 1, generator=gen) / 6).to(DEV)
    bias = torch.zeros(64, device=DEV)
    fa, fb = sa.lookup_pair_convc1(stereo, otf, coords, w, bias)
    ga, gb = sa.lookup_pair_convc1(stereo, pk, coords, w, bias)
    assert torch.equal(fb, gb) and torch.equal(fa, ga)


@pytest.mark.parametrize("shape", [(2, 5, 312, 312), (1, 3, 128, 128), (1, 2, 40, 40), (1, 2, 132, 264), (1, 1, 388, 520), (1, 2, 8, 8)])
def test_mono_lookup_factored(sa, shape):
    """`mono_mode = "factored"`: the lookup kernel combines three packed lines of the RIGHT NORMAL MAP with the
    pixel's left normal (the mono volume has rank 3; pooling and packing are linear).  Agrees with the packed
    mode to fp32 rounding (the scale sits on the coefficients; stored border entries are pooled before the
    contraction) and with the float64 closed form as closely as the packed mode does."""
    b, h, w2, w3 = shape
    gen = torch.Generator().manual_seed(77 + w3)
    nl = torch.nn.functional.normalize(torch.randn(b, 3, h, w2, generator=gen), dim=1).to(DEV)
    nr = torch.nn.functional.normalize(torch.randn(b, 3, h, w3, generator=gen), dim=1).to(DEV)
    B = sa.CorrBlockB200
    old = B.mono_mode
    try:
        B.mono_mode = "factored"
        fc = B.from_normals(nl, nr)
        assert fc._packed is None and fc._packed_nr is not None and fc._packed_nr.shape[0] == b * 3 * h
        B.mono_mode = "packed"
        pk = B.from_normals(nl, nr)
    finally:
        B.mono_mode = old
    stereo = B(torch.randn(b, h, w2, 1, w3, generator=gen).to(DEV))
    levels = O.closed_pyramid(pk.fullcorr.squeeze(3).cpu().numpy(), 4)[:4]
    levels = [lv.reshape(b, h, w2, -1) for lv in levels]
    x = torch.arange(w2, dtype=torch.float32).view(1, 1, 1, w2).expand(b, 1, h, w2) * (w3 / w2)
    for kind, dx in (("left", -torch.rand(b, 1, h, w2, generator=gen) * (w3 / 3)), ("right", torch.rand(b, 1, h, w2, generator=gen) * 60),
                     ("far", (torch.rand(b, 1, h, w2, generator=gen) - 0.5) * 4 * w3), ("int", -torch.randint(0, 9, (b, 1, h, w2), generator=gen).float())):
        coords = torch.cat([x + dx, torch.zeros(b, 1, h, w2)], 1).to(DEV)
        want = pk(coords)
        got = fc(coords)
        assert float((got - want).abs().max()) <= 1e-6, kind         # |V| <= 1: a few ulp
        ref = O.closed_lookup(levels, coords[:, 0].cpu().numpy(), 4)
        assert maxabs(got, ref) < 5e-6, kind
        s1, m1 = B.lookup_pair(stereo, fc, coords)
        s2, m2 = B.lookup_pair(stereo, pk, coords)
        assert torch.equal(m1, got) and torch.equal(s1, s2), kind
    assert torch.equal(fc.fullcorr, pk.fullcorr)  # the reference attributes still work
    w = (torch.randn(64, 36, 1, 1, generator=gen) / 6).to(DEV)
    # lookup + convc1 takes the factored block as it is (sa_lookup_factored_conv): the stereo half is
    # the packed kernel's bit for bit, the mono half sees taps that differ by fp32 rounding before the tf32 product
    bias = (torch.randn(64, generator=gen) / 4).to(DEV)
    fa, fb = sa.lookup_pair_convc1(stereo, fc, coords, w, bias)
    ga, gb = sa.lookup_pair_convc1(stereo, pk, coords, w, bias)
    assert fc._packed is None and fc._packed_nr is not None
    assert torch.equal(fa, ga) and normwise(fb, gb) < 1e-3
    conv = torch.relu(torch.nn.functional.conv2d(pk(coords).double(), w.double(), bias.double()))
    assert normwise(fb, conv) < 1e-3
    # consumers that need the packed volume get it on demand
    assert torch.equal(fc._ensure_packed(), pk._packed) and fc._packed_nr is None


def test_more_than_2_31_packed_floats(sa):
    """SceneFlow size at batch 64 on one GPU (BASELINE config 3 before sharding): each packed array holds 2.6e9
    floats - past int32.  Every sample of the big batch must equal the same sample run alone (64-bit indexing in
    the packers, the TMA maps and the lookup)."""
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs ~35 GB of device memory")
    b, c, h, w = 64, 64, 136, 240
    gen = torch.Generator(device=DEV).manual_seed(9)
    fl = torch.randn(b, c, h, w, device=DEV, generator=gen)
    fr = torch.randn(b, c, h, w, device=DEV, generator=gen)
    nl = torch.nn.functional.normalize(torch.randn(b, 3, h, w, device=DEV, generator=gen), dim=1)
    nr = torch.nn.functional.normalize(torch.randn(b, 3, h, w, device=DEV, generator=gen), dim=1)
    td = torch.rand(b, 1, h, w, device=DEV, generator=gen) * (w / 4)
    tc = torch.rand(b, 1, h, w, device=DEV, generator=gen)
    x = torch.arange(w, device=DEV, dtype=torch.float32).view(1, 1, 1, w).expand(b, 1, h, w)
    coords = torch.cat([x - torch.rand(b, 1, h, w, device=DEV, generator=gen) * (w / 4), torch.zeros(b, 1, h, w, device=DEV)], 1)
    B = sa.CorrBlockB200
    fs = B.from_features(fl, fr, truncate=(td, tc, 0.9))
    fm = B.from_normals(nl, nr)
    assert fs._packed.numel() > 2 ** 31
    s_all, m_all = B.lookup_pair(fs, fm, coords)
    del fs, fm
    for i in (0, 31, 63):
        sl = slice(i, i + 1)
        f1 = B.from_features(fl[sl].contiguous(), fr[sl].contiguous(), truncate=(td[sl].contiguous(), tc[sl].contiguous(), 0.9))
        m1 = B.from_normals(nl[sl].contiguous(), nr[sl].contiguous())
        s1, mm1 = B.lookup_pair(f1, m1, coords[sl].contiguous())
        assert torch.equal(s_all[sl], s1) and torch.equal(m_all[sl], mm1), i


# ------------------------------------------------------------------------------------------ 16-bit storage of the packed pyramid

@pytest.mark.parametrize("storage", ["fp16", "bf16"])
@pytest.mark.parametrize("b,c,h,w2,w3", [(1, 64, 3, 312, 312), (2, 32, 2, 40, 72), (1, 256, 2, 168, 168), (1, 32, 1, 8, 8),
                                         (1, 96, 2, 132, 264)])
def test_half_storage_is_the_fp32_pyramid_rounded(sa, storage, b, c, h, w2, w3):
    """`from_features(..., storage=)` (sa_corr_pack_tf32_half): the 16-bit packed array is the fp32 one rounded to
    nearest, element for element; a lookup from it equals, bit for bit, the fp32 lookup of the widened array; dual
    lookups with a factored / an fp32-packed partner equal the single lookups."""
    from stereoanywhere_b200 import ops

    kind, dt = ops.HALF_KINDS[storage]
    B = sa.CorrBlockB200
    gen = torch.Generator(device=DEV).manual_seed(w2 + w3 + c)
    fl = torch.randn(b, c, h, w2, device=DEV, generator=gen)
    fr = torch.randn(b, c, h, w3, device=DEV, generator=gen)
    tdisp = torch.rand(b, 1, h, w2, device=DEV, generator=gen) * (w3 / 4)
    tconf = torch.rand(b, 1, h, w2, device=DEV, generator=gen)
    x = torch.arange(w2, device=DEV, dtype=torch.float32).view(1, 1, 1, w2).expand(b, 1, h, w2)
    coords = torch.cat([x - torch.rand(b, 1, h, w2, device=DEV, generator=gen) * (w3 / 3) + 2, torch.zeros(b, 1, h, w2, device=DEV)], 1)
    for trunc in (None, (tdisp, tconf, 0.9)):
        full = B.from_features(fl, fr, truncate=trunc)
        half = B.from_features(fl, fr, truncate=trunc, storage=storage)
        assert half._packed is None and half._packed_h.dtype == dt
        assert torch.equal(half._packed_h, full._packed.to(dt))
        widened = torch.ops.sa_b200.lookup_packed(half._packed_h.float(), w3, coords)
        got = half(coords)
        assert torch.equal(got, widened)
        tol = 1e-3 if storage == "fp16" else 1e-2
        assert normwise(got, full(coords)) < tol
        if w2 == w3:   # partners need the same geometry
            nl = torch.nn.functional.normalize(torch.randn(b, 3, h, w2, device=DEV, generator=gen), dim=1)
            nr = torch.nn.functional.normalize(torch.randn(b, 3, h, w3, device=DEV, generator=gen), dim=1)
            for mode in ("factored", "packed"):
                old, B.mono_mode = B.mono_mode, mode
                try:
                    mono = B.from_normals(nl, nr)
                finally:
                    B.mono_mode = old
                s_, m_ = B.lookup_pair(half, mono, coords)
                assert torch.equal(s_, got) and torch.equal(m_, mono(coords)), mode
            dense = B(torch.randn(b, h, w2, 1, w3, device=DEV, generator=gen))
            s_, m_ = B.lookup_pair(half, dense, coords)
            assert torch.equal(s_, got) and torch.equal(m_, dense(coords))
        # the reference attributes still work (formed in fp32 from the feature maps)
        assert torch.equal(half.fullcorr, full.fullcorr)


def test_half_storage_vs_float64_at_kitti_width(sa):
    """TF32 product + fp16 storage against float64 on sampled pixels at W = 312, C = 256: inside the 1e-3 of the TF32
    class; bf16 inside 1e-2."""
    b, c, h, w = 2, 256, 8, 312
    gen = torch.Generator(device=DEV).manual_seed(5)
    fl = torch.randn(b, c, h, w, device=DEV, generator=gen)
    fr = torch.randn(b, c, h, w, device=DEV, generator=gen)
    x = torch.arange(w, device=DEV, dtype=torch.float32).view(1, 1, 1, w).expand(b, 1, h, w)
    coords = torch.cat([x - torch.rand(b, 1, h, w, device=DEV, generator=gen) * 78, torch.zeros(b, 1, h, w, device=DEV)], 1)
    vol64 = torch.einsum("bchw,bchv->bhwv", fl.double(), fr.double()) / 16.0
    want = sampled_closed_lookup(vol64.view(b * h * w, w), coords[:, 0].reshape(-1))
    for storage, tol in (("fp32", 1e-3), ("fp16", 1e-3), ("bf16", 1e-2)):
        got = sa.CorrBlockB200.from_features(fl, fr, storage=storage)(coords).permute(0, 2, 3, 1).reshape(b * h * w, 36).double()
        err = float((got - want).abs().max() / vol64.abs().max())
        print(f"storage {storage}: normwise lookup error {err:.2e}")
        assert err < tol, (storage, err)
