#!/usr/bin/env python
"""Run the hot path a few times at a BASELINE workload (for ncu captures / quick timing).

    python tools/run_path_once.py [--workload c2_kitti_375x1242_b8] [--steps 2] [--variant fused] [--precision tf32]
Prints per-stage CUDA-event timings (ms) so the same command is useful without a profiler.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--iters", type=int, default=4)
    ap.add_argument("--variant", default="fused")
    ap.add_argument("--precision", default="tf32")
    a = ap.parse_args()
    import stereoanywhere_b200 as sa

    sa.CorrBlockB200.precision = a.precision
    B = sa.CorrBlockB200
    b, c, h, w = bench.WORKLOADS[a.workload]
    dev = torch.device("cuda:0")
    _, d = bench.make_inputs(b, c, h, w, dev)
    torch.cuda.synchronize()

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    for s in range(a.steps):
        t = [ev()]
        vs = B.corr(d["fl"], d["fr"]); t.append(ev())
        vm = B.mono_corr(d["nl"], d["nr"]); t.append(ev())
        fs = B(vs, truncate=(d["tdisp"], d["tconf"], 0.9)); t.append(ev())
        fm = B(vm); t.append(ev())
        coords = d["coords0"]
        for _ in range(a.iters):
            if a.variant == "fused":
                B.lookup_pair(fs, fm, coords)
            else:
                fs(coords); fm(coords)
        t.append(ev())
        torch.cuda.synchronize()
        names = ["corr", "mono_corr", "pyr+trunc", "pyr", f"{a.iters}x lookup"]
        print(f"step {s}: " + "  ".join(f"{n} {t[i].elapsed_time(t[i+1]):.3f} ms" for i, n in enumerate(names)))


if __name__ == "__main__":
    main()
