#!/usr/bin/env python
"""Run the tcgen05 correlation a few times at a BASELINE workload (timing / ncu target)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import stereoanywhere_b200 as sa
wl = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
b, c, h, w = bench.WORKLOADS[wl]
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
fl = torch.randn(b, c, h, w, device=dev, generator=g); fr = torch.randn(b, c, h, w, device=dev, generator=g)
for _ in range(3):
    v = torch.ops.sa_b200.corr_volume(fl, fr, "tf32", 1.0)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5):
    v = torch.ops.sa_b200.corr_volume(fl, fr, "tf32", 1.0)
t1.record(); torch.cuda.synchronize()
idx = torch.randint(0, b * h * w * w, (16384,), device=dev, generator=g)
bb, hh, w2, w3 = idx // (h * w * w), (idx // (w * w)) % h, (idx // w) % w, idx % w
ref = (fl[bb, :, hh, w2].double() * fr[bb, :, hh, w3].double()).sum(1) / float(torch.sqrt(torch.tensor(c)))
err = float((v.view(-1)[idx].double() - ref).abs().max() / v.abs().max())
print(f"{wl}: corr tf32 {t0.elapsed_time(t1)/5*1e3:.1f} us/call, sampled normwise err {err:.3e}")
