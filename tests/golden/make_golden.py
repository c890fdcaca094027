#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Everything is seeded.  On the same torch build (2.11.0+cu128, CPU, 8 threads) re-running reproduces
path_small / tiles / reductions / grads bit for bit; model_slice (a whole-model forward: multi-threaded
convolutions) and producers (QR `lstsq`) come back equal to fp32 rounding only (<= 2e-6 relative from run to
run, checked) - each committed file is ONE self-consistent run of the reference, which is what the parity
tests need (the lookups of model_slice were produced from the very volumes stored next to them).  Outputs:

* path_small.npz   - corr / pyramid / lookup / truncation / bins / masked volume / gauss /
                     corruption on small seeded tensors (reference functions called directly).
* model_slice.npz  - tensors captured at the hot-path call sites of a real
                     `StereoAnywhere({}).forward(..., iters=6, test_mode=True)` on a 64x128 pair:
                     inputs of both `corr()` calls, the tensors handed to both block
                     constructors, and (coords, stereo lookup, mono lookup) per GRU iteration.
* producers_chain.npz - SURVEY 8f-3 as the model chains it: full-resolution mono depth -> 1/4 resize -> normals -> depth
                     bins, and weighted_lsq at a model-like size (fixtures of the CUDA producer kernels).
* tiles.npz        - tile enumeration for several (H, W, preset) pairs, blend weights, pad
                     geometry and a stitched output of the real `TileWrapper` around a toy model.
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def path_small():
    CorrBlock1D, U = ref_shim.import_reference_corr()
    g = torch.Generator().manual_seed(1234)
    out = {}

    # --- A1: stereo-like correlation, C=16 -------------------------------------------------
    fl = torch.randn(2, 16, 3, 24, generator=g)
    fr = torch.randn(2, 16, 3, 24, generator=g)
    vol = CorrBlock1D.corr(fl, fr)
    out.update(a1_fl=_np(fl), a1_fr=_np(fr), a1_vol=_np(vol))

    # --- A2: mono correlation from unit normals, then x1.73 ---------------------------------
    nl = torch.nn.functional.normalize(torch.randn(2, 3, 3, 24, generator=g), dim=1)
    nr = torch.nn.functional.normalize(torch.randn(2, 3, 3, 24, generator=g), dim=1)
    mvol = 1.73 * CorrBlock1D.corr(nl, nr)
    out.update(a2_nl=_np(nl), a2_nr=_np(nr), a2_vol=_np(mvol))

    # --- A3 + A4: pyramid and lookup for several widths (even, odd tails, model-like) -------
    for tag, (b, h, w1, w3) in {
        "w24": (2, 3, 24, 24),
        "w39": (1, 2, 39, 39),   # odd at every level: 39,19,9,4
        "w40": (1, 2, 40, 40),   # 40,20,10,5
        "w50x34": (1, 2, 50, 34),  # W2 != W3 (rectangular volume)
    }.items():
        v = torch.randn(b, h, w1, 1, w3, generator=g)
        blk = CorrBlock1D(v, num_levels=4, radius=4)
        for i, p in enumerate(blk.corr_pyramid):
            out[f"a3_{tag}_p{i}"] = _np(p)
        x = torch.arange(w1, dtype=torch.float32).view(1, 1, 1, w1).repeat(b, 1, h, 1)
        y = torch.arange(h, dtype=torch.float32).view(1, 1, h, 1).repeat(b, 1, 1, w1)
        # left-leaning disparities (run off the left border), right-leaning, and far outside
        for ctag, dx in {
            "left": -torch.rand(b, 1, h, w1, generator=g) * (w1 / 2),
            "right": torch.rand(b, 1, h, w1, generator=g) * 12.0,
            "far": (torch.rand(b, 1, h, w1, generator=g) - 0.5) * 6 * w1,
            "int": -torch.randint(0, 8, (b, 1, h, w1), generator=g).float(),
        }.items():
            coords = torch.cat([x + dx, y], dim=1)
            out[f"a4_{tag}_{ctag}_coords"] = _np(coords)
            out[f"a4_{tag}_{ctag}_out"] = _np(blk(coords))
        out[f"a3_{tag}_vol"] = _np(v)

    # lookup with other radius / level count and with pad
    v = torch.randn(1, 2, 32, 1, 32, generator=g)
    x = torch.arange(32, dtype=torch.float32).view(1, 1, 1, 32).repeat(1, 1, 2, 1)
    coords = torch.cat([x - torch.rand(1, 1, 2, 32, generator=g) * 10, torch.zeros(1, 1, 2, 32)], dim=1)
    out["a4_alt_vol"] = _np(v)
    out["a4_alt_coords"] = _np(coords)
    out["a4_alt_r3l2"] = _np(CorrBlock1D(v, num_levels=2, radius=3)(coords))
    out["a4_alt_r2l3"] = _np(CorrBlock1D(v, num_levels=3, radius=2)(coords))
    out["a4_alt_pad23"] = _np(CorrBlock1D(v, num_levels=4, radius=4, pad=[2, 3])(coords))

    # --- A5: truncation mask ------------------------------------------------------------------
    disp = torch.rand(2, 1, 3, 24, generator=g) * 8
    conf = torch.rand(2, 1, 3, 24, generator=g)
    conf[0, 0, 0, :4] = torch.tensor([0.0, 1.0, 0.5, 0.25])
    tmask = U.truncate_corr_volume_v2(disp, conf, conf_th=None, attenuation_gain=0.9)
    out.update(a5_disp=_np(disp), a5_conf=_np(conf), a5_mask=_np(tmask))
    out["a5_mask_th"] = _np(U.truncate_corr_volume_v2(disp, conf, conf_th=0.5, attenuation_gain=0.1))
    # the product the model hands to the stereo block (stereoanywhere.py:253-255)
    v5 = vol.squeeze(3).unsqueeze(1)
    out["a5_product"] = _np((tmask * v5).squeeze(1).unsqueeze(3))

    # --- A6: depth-bin masks + masked volume ----------------------------------------------------
    mde_l = torch.rand(2, 1, 3, 24, generator=g)
    mde_r = torch.rand(2, 1, 3, 24, generator=g)
    mde_l[0, 0, 0, :6] = torch.tensor([0.0, 1.0, 0.125, 0.25, 0.875, 0.9999999])
    mde_r[0, 0, 0, :3] = torch.tensor([1.0, 0.0, 0.5])
    ml = U.generate_masks(mde_l, N=8)
    mr = U.generate_masks(mde_r, N=8)
    mv = mvol.squeeze(3).unsqueeze(1)
    masked = mv * ml.unsqueeze(4) * mr.unsqueeze(3)
    out.update(a6_mde_l=_np(mde_l), a6_mde_r=_np(mde_r), a6_ml=_np(ml), a6_mr=_np(mr), a6_masked=_np(masked))
    out["a6_ml16"] = _np(U.generate_masks(mde_l, N=16))

    # --- A7: corruption pieces ------------------------------------------------------------------
    gz = U.gauss_corr_volume_naive(torch.zeros_like(disp), float(torch.max(v5)))
    out["a7_gauss0"] = _np(gz)
    out["a7_gauss_d"] = _np(U.gauss_corr_volume_naive(disp, 10, 1))
    lm = U.generate_masks(mde_l, N=4)[:, [2]].unsqueeze(4)  # [B,1,H,W,1] fp16 mask, bin 2
    out["a7_binmask"] = _np(lm.squeeze(4))
    rolled = torch.roll(v5, shifts=5, dims=3)
    out["a7_roll5"] = _np(v5 * (1 - lm) + rolled * lm)
    noise = torch.rand(lm.shape, generator=g).to(lm.dtype)  # rand_like(_left_mask) is fp16
    out["a7_noise"] = _np(noise)
    out["a7_noised"] = _np(v5 * (1 - lm) + v5 * noise * lm)
    out["a7_gaussed"] = _np(v5 * (1 - lm) + v5 * gz * lm)
    np.savez_compressed(os.path.join(HERE, "path_small.npz"), **out)
    print("path_small.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


def model_slice():
    pkg = ref_shim.import_reference()
    import importlib

    sa_mod = importlib.import_module("models.stereoanywhere.stereoanywhere")
    RefBlock = sa_mod.CorrBlock1D

    cap = {"corr_in": [], "ctor_in": [], "calls": []}

    class Spy(RefBlock):
        def __init__(self, fullcorr, num_levels=4, radius=4, pad=[0, 0]):
            super().__init__(fullcorr, num_levels=num_levels, radius=radius, pad=pad)
            self._idx = len(cap["ctor_in"])
            cap["ctor_in"].append(fullcorr.detach().clone())

        def __call__(self, coords):
            o = super().__call__(coords)
            cap["calls"].append((self._idx, coords.detach().clone(), o.detach().clone()))
            return o

        @staticmethod
        def corr(a, b):
            cap["corr_in"].append((a.detach().clone(), b.detach().clone()))
            return RefBlock.corr(a, b)

    torch.manual_seed(0)
    random.seed(0)
    model = pkg.StereoAnywhere({}).eval()
    sa_mod.CorrBlock1D = Spy
    try:
        g = torch.Generator().manual_seed(1)
        h, w = 64, 128
        im2 = torch.rand(1, 3, h, w, generator=g)
        im3 = torch.roll(im2, -8, dims=3)
        ramp = torch.linspace(0.3, 0.8, w).view(1, 1, 1, w).repeat(1, 1, h, 1)
        mde2 = ramp
        mde3 = torch.roll(ramp, -8, dims=3)
        iters = 6
        with torch.no_grad():
            disp, _ = model(im2, im3, mde2, mde3, iters=iters, test_mode=True)
    finally:
        sa_mod.CorrBlock1D = RefBlock

    out = {
        "stereo_fl": _np(cap["corr_in"][0][0]),
        "stereo_fr": _np(cap["corr_in"][0][1]),
        "mono_nl": _np(cap["corr_in"][1][0]),
        "mono_nr": _np(cap["corr_in"][1][1]),
        "stereo_ctor": _np(cap["ctor_in"][0]),
        "mono_ctor": _np(cap["ctor_in"][1]),
        "final_disp": _np(disp),
    }
    # reference volumes as produced at stereoanywhere.py:135-136
    out["stereo_vol"] = _np(RefBlock.corr(*cap["corr_in"][0]))
    out["mono_vol"] = _np(1.73 * RefBlock.corr(*cap["corr_in"][1]))
    for n, (idx, coords, o) in enumerate(cap["calls"]):
        it, which = divmod(n, 2)
        assert idx == which
        if which == 0:
            out[f"it{it}_coords"] = _np(coords)
        out[f"it{it}_{'stereo' if which == 0 else 'mono'}"] = _np(o)
    out["iters"] = np.int64(iters)
    # fp16 storage for the two big feature maps would break parity checks; keep fp32 but small
    np.savez_compressed(os.path.join(HERE, "model_slice.npz"), **out)
    print("model_slice.npz", sum(np.asarray(v).nbytes for v in out.values()) / 1e6, "MB raw")


def tiles():
    T = ref_shim.import_reference_tiles()
    out = {}
    cases = {
        "mid_1984x2880": (1984, 2880, 1120, 672, 112),   # `middlebury` preset on config 4
        "def_1984x2880": (1984, 2880, 448, 448, 96),     # `default` preset
        "sf_544x960": (544, 960, 448, 448, 112),         # `sceneflow` preset
        "odd_300x500": (300, 500, 128, 160, 32),
        "fit_96x96": (96, 96, 128, 128, 16),
    }
    for tag, (h, w, th, tw, ov) in cases.items():
        tw_ = T.TileWrapper(torch.nn.Identity(), tile_width=tw, tile_height=th, overlap=ov, device=torch.device("cpu"))
        specs = tw_._enumerate_tiles(h, w)
        out[f"tiles_{tag}"] = np.array([[s.y_start, s.y_end, s.x_start, s.x_end] for s in specs], dtype=np.int64)
        out[f"args_{tag}"] = np.array([h, w, th, tw, ov], dtype=np.int64)
    for (h, w) in [(7, 5), (40, 24), (1, 9)]:
        out[f"blend_{h}x{w}"] = _np(T._make_blend_weight(h, w, torch.device("cpu")))

    # Stitch through the real TileWrapper around a deterministic toy "model" that returns a
    # negative-disparity map depending on absolute content (so seams are exercised).
    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))

        def forward(self, l, r, ml, mr, **kw):
            d = (l.mean(1, keepdim=True) - r.mean(1, keepdim=True)) * 10 + ml * 3 + 0.01 * l.shape[-1]
            return -(d + 0.1 * torch.tanh(mr)), None

    g = torch.Generator().manual_seed(7)
    h, w = 150, 220
    l = torch.rand(1, 3, h, w, generator=g)
    r = torch.rand(1, 3, h, w, generator=g)
    ml = torch.rand(1, 1, h, w, generator=g)
    mr = torch.rand(1, 1, h, w, generator=g)
    wrap = T.TileWrapper(Toy(), tile_width=96, tile_height=80, overlap=24, device=torch.device("cpu"))
    with torch.no_grad():
        st = wrap(l, r, ml, mr)
    out.update(st_l=_np(l), st_r=_np(r), st_ml=_np(ml), st_mr=_np(mr), st_out=_np(st),
               st_args=np.array([80, 96, 24], dtype=np.int64))
    np.savez_compressed(os.path.join(HERE, "tiles.npz"), **out)
    print("tiles.npz", sum(np.asarray(v).nbytes for v in out.values()) / 1e6, "MB raw")


def reductions():
    """the reference's four soft-argmax / entropy reductions (utils/utils.py:112-170)."""
    _, U = ref_shim.import_reference_corr()
    g = torch.Generator().manual_seed(4321)
    out = {}
    for tag, shape, gain in [("sq24", (2, 1, 3, 24, 24), 3.0), ("r40x50", (1, 1, 2, 40, 50), 6.0), ("flat33", (1, 1, 2, 33, 33), 0.05)]:
        v = torch.randn(*shape, generator=g) * gain
        out[f"{tag}_vol"] = _np(v)
        out[f"{tag}_dl"] = _np(U.estimate_left_disparity(v))
        out[f"{tag}_dr"] = _np(U.estimate_right_disparity(v))
        out[f"{tag}_cl"] = _np(U.estimate_left_confidence(v))
        out[f"{tag}_cr"] = _np(U.estimate_right_confidence(v))
    v = torch.from_numpy(out["sq24_vol"])
    out["sq24_dl_pad"] = _np(U.estimate_left_disparity(v, vol_pad=[2, 3]))
    out["sq24_dr_pad"] = _np(U.estimate_right_disparity(v, vol_pad=[2, 3]))
    np.savez_compressed(os.path.join(HERE, "reductions.npz"), **out)
    print("reductions.npz", sum(np.asarray(v).nbytes for v in out.values()) / 1e6, "MB raw")


def grads():
    """gradients of the reference block (autograd through einsum / avg_pool2d / grid_sample and the
    detached truncation product, as in train.py:277,383) on seeded inputs."""
    CorrBlock1D, U = ref_shim.import_reference_corr()
    g = torch.Generator().manual_seed(2468)
    out = {}
    for tag, (b, h, w) in [("w24", (1, 3, 24)), ("w39", (1, 2, 39)), ("w40", (2, 2, 40))]:
        v = torch.randn(b, h, w, 1, w, generator=g).requires_grad_(True)
        x = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w).repeat(b, 1, h, 1)
        coords = [torch.cat([x - torch.rand(b, 1, h, w, generator=g) * (w / 3) + k, torch.zeros(b, 1, h, w)], 1) for k in range(3)]
        wts = [torch.randn(b, 36, h, w, generator=g) for _ in range(3)]
        blk = CorrBlock1D(v, num_levels=4, radius=4)
        loss = sum((blk(c) * wt).sum() for c, wt in zip(coords, wts))
        (dv,) = torch.autograd.grad(loss, v)
        out.update({f"{tag}_vol": _np(v.detach()), f"{tag}_dvol": _np(dv)})
        for k in range(3):
            out[f"{tag}_coords{k}"] = _np(coords[k])
            out[f"{tag}_w{k}"] = _np(wts[k])
        # through the truncation product (mask detached, stereoanywhere.py:203, 253-255)
        disp = torch.rand(b, 1, h, w, generator=g) * (w / 4)
        conf = torch.rand(b, 1, h, w, generator=g)
        tmask = U.truncate_corr_volume_v2(disp, conf, conf_th=None, attenuation_gain=0.9).detach()
        v2 = v.detach().clone().requires_grad_(True)
        blk2 = CorrBlock1D((tmask * v2.squeeze(3).unsqueeze(1)).squeeze(1).unsqueeze(3), num_levels=4, radius=4)
        loss2 = sum((blk2(c) * wt).sum() for c, wt in zip(coords, wts))
        (dv2,) = torch.autograd.grad(loss2, v2)
        out.update({f"{tag}_tdisp": _np(disp), f"{tag}_tconf": _np(conf), f"{tag}_dvol_trunc": _np(dv2)})
    # corr(): gradients to the feature maps
    fl = torch.randn(2, 32, 3, 24, generator=g).requires_grad_(True)
    fr = torch.randn(2, 32, 3, 24, generator=g).requires_grad_(True)
    wv = torch.randn(2, 3, 24, 1, 24, generator=g)
    vol = CorrBlock1D.corr(fl, fr)
    dfl, dfr = torch.autograd.grad((vol * wv).sum(), (fl, fr))
    out.update(c_fl=_np(fl.detach()), c_fr=_np(fr.detach()), c_w=_np(wv), c_dfl=_np(dfl), c_dfr=_np(dfr))
    np.savez_compressed(os.path.join(HERE, "grads.npz"), **out)
    print("grads.npz", sum(np.asarray(v).nbytes for v in out.values()) / 1e6, "MB raw")


def producers():
    """generate_masks, estimate_normals (kornia stub), weighted_lsq of the reference."""
    _, U = ref_shim.import_reference_corr()
    g = torch.Generator().manual_seed(1357)
    out = {}
    mde = torch.rand(2, 1, 12, 20, generator=g)
    mde[0, 0, 0, 0], mde[0, 0, 0, 1], mde[1, 0, 3, 3] = 1.0, 0.0, 0.5
    out["gm_mde"] = _np(mde)
    for n in (4, 8, 16):
        out[f"gm_masks{n}"] = _np(U.generate_masks(mde, N=n))
    depth = torch.rand(2, 1, 12, 20, generator=g)
    out["en_depth"] = _np(depth)
    out["en_normals"] = _np(U.estimate_normals(depth, normal_gain=20 / 10))
    b, h, w = 3, 24, 40
    mono = torch.rand(b, 2, h, w, generator=g)
    disp = (mono * torch.tensor([3.0, 5.0, 0.5]).view(b, 1, 1, 1) + torch.tensor([1.0, -0.5, 2.0]).view(b, 1, 1, 1)
            + 0.05 * torch.randn(b, 2, h, w, generator=g))
    conf = torch.rand(b, 2, h, w, generator=g)
    sc, sh = U.weighted_lsq(mono, disp, conf)
    out.update(wl_mono=_np(mono), wl_disp=_np(disp), wl_conf=_np(conf), wl_scale=_np(sc), wl_shift=_np(sh))
    np.savez_compressed(os.path.join(HERE, "producers.npz"), **out)
    print("producers.npz", sum(np.asarray(v).nbytes for v in out.values()) / 1e6, "MB raw")


def producers_chain():
    """SURVEY 8f-3, the chain as the model runs it (stereoanywhere.py:109-114, 138-139): full-resolution mono depth ->
    1/4 bilinear resize (align_corners) -> estimate_normals (kornia stub) -> generate_masks; and weighted_lsq at a
    model-like size with a quantile window that cuts through ties (relu zeros)."""
    import torch.nn.functional as F

    _, U = ref_shim.import_reference_corr()
    g = torch.Generator().manual_seed(2468)
    out = {}
    for tag, (b, h, w) in {"a": (2, 64, 96), "b": (1, 36, 52)}.items():
        yy = torch.linspace(0, 1, h).view(1, 1, h, 1)
        xx = torch.linspace(0, 1, w).view(1, 1, 1, w)
        mde = (0.3 + 0.5 * xx + 0.1 * torch.sin(7 * yy + 3 * xx) + 0.05 * torch.rand(b, 1, h, w, generator=g)).clamp(0, 1)
        mde[0, 0, 0, 0] = 1.0
        low = F.interpolate(mde, scale_factor=1 / 4, mode="bilinear", align_corners=True)
        w_low = w // 4
        normals = U.estimate_normals(low, normal_gain=(w_low / 10))
        out[f"{tag}_mde"] = _np(mde)
        out[f"{tag}_low"] = _np(low)
        out[f"{tag}_normals"] = _np(normals)
        out[f"{tag}_masks8"] = _np(U.generate_masks(low, N=8))
        out[f"{tag}_gain"] = np.array([w_low / 10], dtype=np.float64)
    b, h, w = 4, 48, 64
    mono = torch.rand(b, 2, h, w, generator=g)
    disp = (mono * torch.tensor([30.0, 12.0, 5.0, 1.5]).view(b, 1, 1, 1) - torch.tensor([10.0, 2.0, 0.0, -1.0]).view(b, 1, 1, 1)
            + 0.3 * torch.randn(b, 2, h, w, generator=g))      # a third of the first sample is negative -> relu zeros
    conf = torch.rand(b, 2, h, w, generator=g) * 0.01
    sc, sh = U.weighted_lsq(mono, disp, conf)
    out.update(wl_mono=_np(mono), wl_disp=_np(disp), wl_conf=_np(conf), wl_scale=_np(sc), wl_shift=_np(sh))
    np.savez_compressed(os.path.join(HERE, "producers_chain.npz"), **out)
    print("producers_chain.npz", sum(np.asarray(v).nbytes for v in out.values()) / 1e6, "MB raw")


if __name__ == "__main__":
    torch.set_num_threads(8)
    table = {"path_small": path_small, "model_slice": model_slice, "tiles": tiles, "reductions": reductions, "grads": grads,
             "producers": producers, "producers_chain": producers_chain}
    for name in sys.argv[1:] or list(table):
        table[name]()
