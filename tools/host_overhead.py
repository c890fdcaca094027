#!/usr/bin/env python
"""Host-side cost of one lookup call (eager, no CUDA graph) at the smallest workload, where the kernel itself
takes ~3 us: wall time per call through the public API, the torch custom op, and the bare ctypes call."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import stereoanywhere_b200 as sa
from stereoanywhere_b200 import ops, _lib
B = sa.CorrBlockB200
b, c, h, w = bench.WORKLOADS["c1_384x512_b1"]
dev = torch.device("cuda:0")
_, d = bench.make_inputs(b, c, h, w, dev, seed=0)
fs = B.from_features(d["fl"], d["fr"], truncate=(d["tdisp"], d["tconf"], 0.9)); fm = B.from_normals(d["nl"], d["nr"])
coords = d["coords0"]
def wall(fn, n=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
lib = _lib.load()
oa = torch.empty(b, 36, h, w, device=dev); ob = torch.empty_like(oa)
st = torch.cuda.current_stream().cuda_stream
args = (fs._packed.data_ptr(), fm._packed.data_ptr(), w, coords.data_ptr(), coords.stride(0), oa.data_ptr(), ob.data_ptr(), b, h, w, st)
print(f"public API  B.lookup_pair        : {wall(lambda: B.lookup_pair(fs, fm, coords)):6.1f} us per call")
print(f"public API  fs(coords)           : {wall(lambda: fs(coords)):6.1f} us per call")
print(f"torch op    lookup_packed2       : {wall(lambda: torch.ops.sa_b200.lookup_packed2(fs._packed, fm._packed, w, coords)):6.1f} us per call")
print(f"python impl ops._lookup_packed   : {wall(lambda: ops._lookup_packed(fs._packed, fm._packed, w, coords)):6.1f} us per call")
print(f"ctypes      sa_lookup_packed     : {wall(lambda: lib.sa_lookup_packed(*args)):6.1f} us per call")
print(f"torch       2 x torch.empty      : {wall(lambda: (torch.empty(b, 36, h, w, device=dev), torch.empty(b, 36, h, w, device=dev))):6.1f} us per call")
