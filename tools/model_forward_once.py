#!/usr/bin/env python
"""Whole-model view (GPU box; reference from oracle/_ref): wall time of the UNMODIFIED `StereoAnywhere.forward`
(random init, 32 iterations, test mode) with its own CorrBlock1D and with the B200 block installed by
`integration.install` (protocol / fused wiring).  Everything outside the correlation block (encoders, hourglass,
update block: cuDNN / ATen) is the reference's own code in all three runs."""
import importlib
import os
import random
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

DEV = "cuda:0"


def main():
    import stereoanywhere_b200 as sa  # noqa: F401
    from stereoanywhere_b200 import integration

    pkg = ref_shim.import_reference()
    sa_mod = importlib.import_module("models.stereoanywhere.stereoanywhere")
    sizes = [(1, 384, 512), (1, 384, 1248), (2, 384, 1248)]
    for margs in ({}, {"use_aggregate_mono_vol": False}):
        torch.manual_seed(0)
        model = pkg.StereoAnywhere(dict(margs)).to(DEV).eval()
        for b, h, w in sizes:
            g = torch.Generator().manual_seed(1)
            im2 = torch.rand(b, 3, h, w, generator=g)
            im3 = torch.roll(im2, -8, dims=3)
            yy = torch.linspace(0, 1, h).view(1, 1, h, 1)
            xx = torch.linspace(0, 1, w).view(1, 1, 1, w)
            mde = (torch.linspace(0.3, 0.8, w).view(1, 1, 1, w) + 0.08 * torch.sin(6.3 * yy + 2.0 * xx) * torch.cos(9.1 * xx - 3.0 * yy)
                   + 0.05 * yy).clamp(0, 1).expand(b, 1, h, w).contiguous()
            inputs = [t.to(DEV) for t in (im2, im3, mde, torch.roll(mde, -8, dims=3))]

            def fwd():
                random.seed(0)
                with torch.no_grad():
                    d, _ = model(*inputs, iters=32, test_mode=True)
                return d

            rows = []
            for name, setup in (("reference CorrBlock1D", lambda: integration.uninstall(sa_mod)),
                                ("B200 protocol", lambda: integration.install(sa_mod, fused=False)),
                                ("B200 fused", lambda: integration.install(sa_mod, fused=True))):
                setup()
                for _ in range(2):
                    d = fwd()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                n = 5
                for _ in range(n):
                    d = fwd()
                torch.cuda.synchronize()
                rows.append((name, (time.perf_counter() - t0) / n * 1e3, d))
            integration.uninstall(sa_mod)
            ref = rows[0][2]
            print(f"--- model args {margs}, batch {b}, {h}x{w}, 32 iterations")
            for name, ms, d in rows:
                print(f"   {name:24s} {ms:8.2f} ms per forward ({b / ms * 1e3:7.2f} pairs/s)   EPE vs reference {float((d - ref).abs().mean()):.2e} px")


if __name__ == "__main__":
    main()
