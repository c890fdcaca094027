#!/usr/bin/env python
"""How much of the dual lookup's time depends on the packed lines surviving in the L2 from one GRU iteration to the next:
32 graph-replayed launches with (a) the benchmark's slowly drifting coordinates (a pixel mostly re-reads the line it read
an iteration ago), (b) the SAME coordinates every iteration, (c) fresh random coordinates every iteration (no reuse)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import stereoanywhere_b200 as sa

B = sa.CorrBlockB200
b, c, h, w = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
dev = torch.device("cuda:0")
_, d = bench.make_inputs(b, c, h, w, dev, seed=0)
fs = B.from_features(d["fl"], d["fr"], truncate=(d["tdisp"], d["tconf"], 0.9))
fm = B.from_normals(d["nl"], d["nr"])
x = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, 1, w).expand(b, 1, h, w)
g = torch.Generator(device=dev).manual_seed(7)
sets = {
    "drifting (bench)": [d["coords0"] + k * d["delta"] for k in range(32)],
    "same every iteration": [d["coords0"].clone() for _ in range(32)],
    "fresh random every iteration": [torch.cat([x - torch.rand(b, 1, h, w, device=dev, generator=g) * (w / 4), torch.zeros_like(x)], 1).contiguous()
                                     for _ in range(32)],
}
for name, coords in sets.items():
    B.lookup_pair(fs, fm, coords[0]); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for k in range(32):
            o = B.lookup_pair(fs, fm, coords[k])
    ts = []
    for i in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3 / 32)
    ts.sort()
    print(f"{name:32s} {ts[len(ts) // 2]:6.2f} us per dual lookup")
    del gr, o
