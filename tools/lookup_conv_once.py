#!/usr/bin/env python
"""Time the fused lookup+convc1 kernel against lookup_pair + two cuDNN 1x1 convolutions (c2 size)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import stereoanywhere_b200 as sa
B = sa.CorrBlockB200
b, c, h, w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD]
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
va = torch.randn(b, h, w, 1, w, device=dev, generator=g); vb = torch.randn(b, h, w, 1, w, device=dev, generator=g)
ba, bb = B(va), B(vb)
del va, vb
conv = torch.nn.Conv2d(36, 64, 1).to(dev)
x = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, 1, w).expand(b, 1, h, w)
coords = torch.cat([x - torch.rand(b, 1, h, w, device=dev, generator=g) * (w / 4), torch.zeros(b, 1, h, w, device=dev)], 1)
def fused():
    return sa.lookup_pair_convc1(ba, bb, coords, conv.weight, conv.bias)
def unfused():
    s, m = B.lookup_pair(ba, bb, coords)
    return torch.relu(conv(s)), torch.relu(conv(m))
def timeit(fn, n=20):
    with torch.no_grad():
        for _ in range(3): fn()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(n): out = fn()
        gr.replay(); torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); gr.replay(); t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n * 1e3, out
tf, of = timeit(fused)
tu, ou = timeit(unfused)
err = float((of[0] - ou[0]).abs().max() / ou[0].abs().max())
print(f"fused lookup+convc1: {tf:.1f} us   lookup_pair + 2x(cuDNN conv1x1 + relu): {tu:.1f} us   normwise diff {err:.2e}")
# the mono block in factored form (from_normals, default mode): sa_lookup_factored_conv
nl = torch.nn.functional.normalize(torch.randn(b, 3, h, w, device=dev, generator=g), dim=1)
nr = torch.nn.functional.normalize(torch.randn(b, 3, h, w, device=dev, generator=g), dim=1)
del bb
bb = B.from_normals(nl, nr)
tf, of = timeit(fused)
tu, ou = timeit(unfused)
err = float((of[1] - ou[1]).abs().max() / ou[1].abs().max())
print(f"factored mono ({B.mono_mode}): fused {tf:.1f} us   unfused {tu:.1f} us   normwise diff (mono) {err:.2e}")
