#!/usr/bin/env python
"""sa_corr_pack_tf32_half: time at a workload and check the 16-bit packed array against the fp32 one rounded on the host."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import stereoanywhere_b200 as sa
from stereoanywhere_b200 import _lib, ops
lib = _lib.load()
b, c, h, w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD]
dev = torch.device("cuda:0")
_, d = bench.make_inputs(b, c, h, w, dev, seed=0)
rows, nblk = b * h * w, w // 8 + 9
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = sa.CorrBlockB200.from_features(d["fl"], d["fr"], truncate=(d["tdisp"], d["tconf"], 0.9))._packed
for kind, dt in ((1, torch.float16), (2, torch.bfloat16)):
    out = torch.empty((rows, nblk * 32), dtype=dt, device=dev)
    def run():
        rc = lib.sa_corr_pack_tf32_half(d["fl"].data_ptr(), d["fr"].data_ptr(), b, c, h, w, w, ops._divisor(c), 1.0,
                                        d["tdisp"].data_ptr(), d["tconf"].data_ptr(), 0.9, kind, out.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.sa_last_error()
    ts = []
    for i in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    same = torch.equal(out, ref.to(dt))
    print(f"half_kind {kind} ({dt}): {ts[len(ts)//2]:.1f} us; equals the fp32 packed array rounded to {dt}: {same}; "
          f"max |err| / max |ref| = {float((out.float() - ref).abs().max() / ref.abs().max()):.2e}")
