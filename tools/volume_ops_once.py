#!/usr/bin/env python
"""Time the standalone A5 / A6 / A7 volume kernels and the plain corr kernels at a BASELINE workload
(CUDA events, L2 flushed between repetitions); bytes = what each kernel must read + write."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import stereoanywhere_b200 as sa
B = sa.CorrBlockB200
name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
b, c, h, w = bench.WORKLOADS[name]
dev = torch.device("cuda:0")
_, d = bench.make_inputs(b, c, h, w, dev, seed=0)
peak, _ = bench.load_peaks()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(1)
vol = torch.randn(b, 1, h, w, w, device=dev, generator=g)
mde_l, mde_r = torch.rand(b, 1, h, w, device=dev, generator=g), torch.rand(b, 1, h, w, device=dev, generator=g)
binm = (mde_l < 0.25).to(torch.float16).unsqueeze(4)
V = vol.numel() * 4

def t(fn, reps=5):
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
        del out
    ts.sort(); return ts[len(ts) // 2]

rows = [
    ("A1 corr tf32 (volume written)", lambda: B.corr(d["fl"], d["fr"]), 2 * b * c * h * w * 4 + V),
    ("A2 mono corr fp32 x1.73", lambda: B.mono_corr(d["nl"], d["nr"]), V),
    ("A3 pack, any volume", lambda: B(vol.squeeze(1).unsqueeze(3)), V + b * h * w * (w // 8 + 9) * 128),
    ("A3+A5 pack with truncation", lambda: B(vol.squeeze(1).unsqueeze(3), truncate=(d["tdisp"], d["tconf"], 0.9)), V + b * h * w * (w // 8 + 9) * 128),
    ("A5 truncation mask only", lambda: sa.truncation_mask(d["tdisp"], d["tconf"], 0.9), V),
    ("A5 mask * volume", lambda: sa.truncation_mask(d["tdisp"], d["tconf"], 0.9, vol), 2 * V),
    ("A6 masked volume, 8 bins", lambda: sa.masked_volume(vol, mde_l, mde_r, 8), V + 8 * V),
    ("A2+A6 masked mono volume from normals", lambda: sa.masked_mono_volume(d["nl"], d["nr"], mde_l, mde_r, 8), 8 * V),
    ("A7 corrupt: roll", lambda: sa.corrupt_volume(vol, binm, "roll", shift=17), 3 * V),
]
print(f"{name}: volume {V / 1e6:.0f} MB, measured HBM peak {peak:.0f} GB/s")
for label, fn, byt in rows:
    us = t(fn)
    print(f"  {label:40s} {us:8.1f} us  {byt / us / 1e3:7.0f} GB/s  {byt / us / 1e3 / peak:5.2f} of peak")
