#!/usr/bin/env python
"""launches, host syncs and time of the producer chain at SceneFlow size, batch 64 - the reference's own
functions (oracle/_ref, run by ATen on the GPU) against the two producer kernels."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402


def count(fn, reps=3):
    from torch.profiler import ProfilerActivity, profile

    fn(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    ev = prof.events()
    kernels = sum(1 for e in ev if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower() and "memset" not in e.name.lower())
    syncs = sum(1 for e in ev if e.name in ("cudaStreamSynchronize", "cudaDeviceSynchronize", "aten::item", "aten::_local_scalar_dense", "aten::nonzero"))
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return kernels, syncs, (time.perf_counter() - t0) / reps * 1e3


def main():
    import torch.nn.functional as F

    from stereoanywhere_b200 import producers as P

    _, U = ref_shim.import_reference_corr()
    dev = "cuda:0"
    b, h, w = 64, 544, 960
    g = torch.Generator().manual_seed(0)
    mde = torch.rand(b, 1, h, w, generator=g).to(dev)
    gain = (w // 4) / 10

    def ref_chain():
        low = F.interpolate(mde, scale_factor=1 / 4, mode="bilinear", align_corners=True)
        return low, U.estimate_normals(low, normal_gain=gain), U.generate_masks(low, N=8)

    def ours_chain():
        return P.mono_inputs(mde, 2, gain, 8)

    low = ref_chain()[0]
    mono2 = torch.cat([low, low.flip(3)], 1)
    disp2 = 30 * mono2 - 8 + torch.randn(mono2.shape, device=dev)
    conf2 = torch.rand(mono2.shape, device=dev)
    for name, fn in (("reference resize + estimate_normals + generate_masks", ref_chain), ("sa_mono_inputs", ours_chain),
                     ("reference weighted_lsq (per-sample loop)", lambda: U.weighted_lsq(mono2, disp2, conf2)),
                     ("sa_weighted_lsq", lambda: P.weighted_lsq_b200(mono2, disp2, conf2))):
        k, s, ms = count(fn)
        print(f"{name:58s} batch {b} @ {h}x{w}: {k:5d} kernel launches, {s:4d} host syncs, {ms:8.3f} ms wall")


if __name__ == "__main__":
    main()
