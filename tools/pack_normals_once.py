#!/usr/bin/env python
"""Run the mono packer (from_normals) a few times at a BASELINE workload (timing / ncu target)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import stereoanywhere_b200 as sa
wl = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
b, c, h, w = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
_, d = bench.make_inputs(b, c, h, w, dev, seed=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(6):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); blk = sa.CorrBlockB200.from_normals(d["nl"], d["nr"]); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3); del blk
wr = b * h * w * (w // 8 + 9) * 128
print(f"{wl}: mono pack {sorted(ts)[len(ts)//2]:.1f} us = {wr / sorted(ts)[len(ts)//2] / 1e3:.0f} GB/s written ({wr/1e6:.0f} MB)")
