#!/usr/bin/env python
"""Pure-write and copy HBM bandwidth reference (torch fill_ / copy_ on 1.5 GB, CUDA events)."""
import torch
dev = torch.device("cuda:0")
n = 1472 << 20
a = torch.empty(n // 4, dtype=torch.float32, device=dev); b = torch.empty_like(a)
def t(fn, reps=10):
    ts = []
    for i in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
tf = t(lambda: a.fill_(1.0)); tc = t(lambda: b.copy_(a)); tr = t(lambda: a.sum())
print(f"fill_ {n/1e6:.0f} MB: {tf:.1f} us = {n/tf/1e3:.0f} GB/s written | copy_: {tc:.1f} us = {2*n/tc/1e3:.0f} GB/s r+w | sum: {tr:.1f} us = {n/tr/1e3:.0f} GB/s read")
