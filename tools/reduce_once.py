#!/usr/bin/env python
"""Time the 8f-2 reduction kernels (one read of the volume per pair) against the reference's four ATen
softmax passes on the same GPU (oracle functions moved to the device), c2-size aggregated volume."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import stereoanywhere_b200 as sa
name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
b, c, h, w = bench.WORKLOADS[name]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
vol = torch.randn(b, 1, h, w, w, device=dev, generator=g) * 4
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=10):
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], out
def aten_all():
    idx3 = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, 1, w)
    idx2 = idx3.view(1, 1, w, 1)
    v = vol.squeeze(1)
    p3, p2 = torch.softmax(v, 3), torch.softmax(v, 2)
    dl = torch.arange(w, device=dev).view(1, 1, w) - (p3 * idx3).sum(3)
    dr = (p2 * idx2).sum(2) - torch.arange(w, device=dev).view(1, 1, w)
    cl = 1 + (p3 * torch.log2(p3 + 1e-6)).sum(3) / torch.log2(torch.tensor(float(w)))
    cr = 1 + (p2 * torch.log2(p2 + 1e-6)).sum(2) / torch.log2(torch.tensor(float(w)))
    return dl, dr, cl, cr
t_d, (dl, dr) = timeit(lambda: sa.estimate_disparities(vol))
t_c, (cl, cr) = timeit(lambda: sa.estimate_confidences(vol))
t_a, ref = timeit(aten_all, reps=5)
byt = vol.numel() * 4
print(f"{name}: volume {byt/1e6:.0f} MB | softargmax pair {t_d:.1f} us = {byt/t_d/1e3:.0f} GB/s | entropy pair {t_c:.1f} us = {byt/t_c/1e3:.0f} GB/s"
      f" | ATen, 4 reductions {t_a:.1f} us")
print("max abs diff vs ATen-on-GPU: dl %.2e dr %.2e cl %.2e cr %.2e" % (
    float((dl[:, 0] - ref[0]).abs().max()), float((dr[:, 0] - ref[1]).abs().max()),
    float((cl[:, 0] - ref[2]).abs().max()), float((cr[:, 0] - ref[3]).abs().max())))
