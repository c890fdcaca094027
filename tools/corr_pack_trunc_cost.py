#!/usr/bin/env python
"""What the truncation mask costs inside sa_corr_pack_tf32: no truncation / the benchmark's random per-pixel
disparities (every warp diverges over the sigmoid band) / a smooth disparity field (coherent warps)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import stereoanywhere_b200 as sa
B = sa.CorrBlockB200
b, c, h, w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD]
storage = sys.argv[2] if len(sys.argv) > 2 else "fp32"
dev = torch.device("cuda:0")
_, d = bench.make_inputs(b, c, h, w, dev, seed=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
x = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, 1, w).expand(b, 1, h, w)
smooth = (20 + 10 * torch.sin(x / 40)).contiguous()
def timeit(fn, reps=10):
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); blk = fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
        del blk
    ts.sort()
    return ts[len(ts) // 2]
for name, t in (("no truncation", None), ("random disparities (bench)", (d["tdisp"], d["tconf"], 0.9)),
                ("smooth disparities", (smooth, d["tconf"], 0.9))):
    print(f"{storage} {name:32s} {timeit(lambda: B.from_features(d['fl'], d['fr'], truncate=t, storage=storage)):7.1f} us")
