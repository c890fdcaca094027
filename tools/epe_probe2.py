#!/usr/bin/env python
"""Localise the iteration-1 disparity difference: record intermediates of the reference forward under both blocks."""
import importlib
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

DEV = "cuda:0"


def main():
    import stereoanywhere_b200 as sa
    from stereoanywhere_b200 import integration

    pkg = ref_shim.import_reference()
    sa_mod = importlib.import_module("models.stereoanywhere.stereoanywhere")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    h, w = 384, 512
    g = torch.Generator().manual_seed(1)
    im2 = torch.rand(1, 3, h, w, generator=g)
    im3 = torch.roll(im2, -8, dims=3)
    ramp = torch.linspace(0.3, 0.8, w).view(1, 1, 1, w).expand(1, 1, h, w).contiguous()
    inputs = [t.to(DEV) for t in (im2, im3, ramp, torch.roll(ramp, -8, dims=3))]
    torch.manual_seed(0)
    model = pkg.StereoAnywhere({}).to(DEV).eval()
    rec = {}
    names = ["estimate_left_disparity", "estimate_right_disparity", "estimate_left_confidence", "weighted_lsq",
             "softlrc", "handcrafted_mirror_detector"]
    orig = {n: getattr(sa_mod, n) for n in names}

    def wrap(n):
        def f(*a, **k):
            out = orig[n](*a, **k)
            rec.setdefault(n, []).append(([x.detach().clone() if torch.is_tensor(x) else x for x in a],
                                          [o.detach().clone() for o in (out if isinstance(out, tuple) else (out,))]))
            return out
        return f

    for n in names:
        setattr(sa_mod, n, wrap(n))

    class RecBlockMixin:
        pass

    def run(block_cls):
        rec.clear()
        sa_mod.CorrBlock1D = block_cls
        corr_orig = block_cls.corr
        vols = []

        class Rec(block_cls):
            @staticmethod
            def corr(a, b):
                v = corr_orig(a, b)
                vols.append(v.detach().clone())
                return v

            def __call__(self, coords):
                o = super().__call__(coords)
                rec.setdefault("lookup", []).append((coords.detach().clone(), o.detach().clone()))
                return o

        sa_mod.CorrBlock1D = Rec
        random.seed(0)
        with torch.no_grad():
            d, _ = model(*inputs, iters=1, test_mode=True)
        torch.cuda.synchronize()
        return d, vols, {k: v for k, v in rec.items()}

    integration.uninstall(sa_mod)
    ref_block = sa_mod.CorrBlock1D
    d_ref, v_ref, r_ref = run(ref_block)
    sa.CorrBlockB200.precision = "fp32"
    d_b, v_b, r_b = run(sa.CorrBlockB200)

    def diff(a, b):
        return f"max|d| {float((a - b).abs().max()):.3e}  max|ref| {float(a.abs().max()):.3e}  mean|ref| {float(a.abs().mean()):.3e}"

    print("disp:", diff(d_ref, d_b))
    print("stereo vol:", diff(v_ref[0], v_b[0]))
    print("mono vol (before 1.73):", diff(v_ref[1], v_b[1]), " std over W3 of ref:", float(v_ref[1].std()))
    for n in names:
        for i, ((ai, ao), (bi, bo)) in enumerate(zip(r_ref.get(n, []), r_b.get(n, []))):
            for j, (x, y) in enumerate(zip(ai, bi)):
                if torch.is_tensor(x):
                    print(f"{n}[{i}] in{j}:", diff(x, y))
            for j, (x, y) in enumerate(zip(ao, bo)):
                print(f"{n}[{i}] out{j}:", diff(x, y), " values", x.flatten()[:3].tolist() if x.numel() < 8 else "")
    for i, ((c_r, o_r), (c_b, o_b)) in enumerate(zip(r_ref["lookup"], r_b["lookup"])):
        print(f"lookup[{i}] coords:", diff(c_r, c_b), " out:", diff(o_r, o_b))
        # same coords through both blocks?  feed the reference's coords to our block: done in the parity tests
    # spectrum of the hourglass input volume: how flat is the softmax?
    vol = r_ref["estimate_left_disparity"][0][0][0]
    print("agg volume: std along W3 (mean over pixels)", float(vol.std(dim=-1).mean()), " range", float(vol.max() - vol.min()))


if __name__ == "__main__":
    main()
