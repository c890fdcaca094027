#!/usr/bin/env python
"""Time sa_corr_pack_tf32 (corr + truncation + pyramid in the GEMM epilogue) against the two-step path
(sa_corr_tf32 -> sa_pack_pyramid) on one workload; CUDA events, L2 flushed between repetitions."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import stereoanywhere_b200 as sa
B = sa.CorrBlockB200
name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
b, c, h, w = bench.WORKLOADS[name]
dev = torch.device("cuda:0")
_, d = bench.make_inputs(b, c, h, w, dev, seed=0)
t = (d["tdisp"], d["tconf"], 0.9)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def fused():
    return B.from_features(d["fl"], d["fr"], truncate=t)
def two_step():
    return B(B.corr(d["fl"], d["fr"]), truncate=t)
def timeit(fn):
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); blk = fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
        del blk
    ts.sort()
    return ts[len(ts) // 2], ts[0]
f_med, f_min = timeit(fused)
t_med, t_min = timeit(two_step)
p = b * h * w
wr = p * (w // 8 + 9) * 128
rd = 2 * b * c * h * w * 4
print(f"{name}: fused corr+trunc+pack {f_med:.1f} us (min {f_min:.1f}) = {(rd + wr) / f_med / 1e3:.0f} GB/s of {(rd + wr) / 1e6:.0f} MB"
      f" | two-step {t_med:.1f} us (min {t_min:.1f})")
a, bb = fused(), two_step()
print("bit-identical packed arrays:", bool(torch.equal(a._packed, bb._packed)))
