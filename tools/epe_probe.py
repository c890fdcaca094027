#!/usr/bin/env python
"""Where does the end-to-end disparity difference of the real model come from?  (GPU box; reference from oracle/_ref.)

Runs `StereoAnywhere.forward` (random init, seed 0, 384x512, 32 iterations) with the reference block and with the B200
block in several settings and prints the EPE of each against the reference's own fp32 run:
  * the reference itself with TF32 matmul enabled for its einsum (its natural sensitivity to TF32 rounding),
  * B200 protocol wiring with precision fp32 / tf32,
  * B200 fused wiring.
Also the disparity after fewer iterations, to see how the difference grows along the GRU loop.
"""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

DEV = "cuda:0"


def main():
    import importlib

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
    B = sa.CorrBlockB200
    for margs in ({}, {"use_aggregate_mono_vol": False}):
        torch.manual_seed(0)
        model = pkg.StereoAnywhere(dict(margs)).to(DEV).eval()

        def fwd(iters):
            random.seed(0)
            with torch.no_grad():
                d, _ = model(*inputs, iters=iters, test_mode=True)
            torch.cuda.synchronize()
            return d

        for iters in (1, 4, 12, 32):
            integration.uninstall(sa_mod)
            ref = fwd(iters)
            torch.backends.cuda.matmul.allow_tf32 = True
            ref_tf32 = fwd(iters)
            torch.backends.cuda.matmul.allow_tf32 = False
            rows = [("reference, einsum in TF32", ref_tf32)]
            for prec in ("fp32", "tf32", "tf32x3"):
                if prec == "tf32x3" and not hasattr(B, "_has_tf32x3"):
                    continue
                B.precision = prec
                integration.install(sa_mod, fused=False)
                rows.append((f"B200 protocol {prec}", fwd(iters)))
                integration.install(sa_mod, fused=True)
                rows.append((f"B200 fused {prec}", fwd(iters)))
                integration.uninstall(sa_mod)
            B.precision = "tf32"
            print(f"--- model args {margs}, iters {iters}: mean |disp| {float(ref.abs().mean()):.2f} px")
            for name, d in rows:
                print(f"   {name:32s} EPE {float((d - ref).abs().mean()):.3e}  max {float((d - ref).abs().max()):.3e}")


if __name__ == "__main__":
    main()
