#!/usr/bin/env python
"""First-light check of the tcgen05 correlation kernel: small + model-size shapes vs fp64."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stereoanywhere_b200 as sa

dev = "cuda:0"
for (b, c, h, w) in [(1, 32, 1, 128), (1, 64, 2, 128), (1, 256, 3, 312), (2, 128, 2, 168), (1, 256, 2, 768), (8, 256, 96, 312)]:
    g = torch.Generator(device=dev).manual_seed(0)
    fl = torch.randn(b, c, h, w, device=dev, generator=g)
    fr = torch.randn(b, c, h, w, device=dev, generator=g)
    v = torch.ops.sa_b200.corr_volume(fl, fr, "tf32", 1.0)
    torch.cuda.synchronize()
    if b * h * w * w < 4e6:
        ref = torch.einsum("bchw,bchv->bhwv", fl.double(), fr.double()) / float(torch.sqrt(torch.tensor(c)))
        d = (v.view_as(ref).double() - ref)
        err = float(d.abs().max() / ref.abs().max())
        bad = (d.abs() > 1e-2 * ref.abs().max()).nonzero()
        print(f"B{b} C{c} H{h} W{w}: normwise err {err:.3e}  bad={len(bad)}  first bad={bad[:4].tolist()}")
        if len(bad):
            print("   got", v.view_as(ref)[tuple(bad[0].tolist())].item(), "want", ref[tuple(bad[0].tolist())].item())
    else:
        idx = torch.randint(0, b * h * w * w, (8192,), device=dev, generator=g)
        bb, hh, w2, w3 = idx // (h * w * w), (idx // (w * w)) % h, (idx // w) % w, idx % w
        ref = (fl[bb, :, hh, w2].double() * fr[bb, :, hh, w3].double()).sum(1) / 16.0
        err = float((v.view(-1)[idx].double() - ref).abs().max() / v.abs().max())
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5):
            torch.ops.sa_b200.corr_volume(fl, fr, "tf32", 1.0)
        t1.record(); torch.cuda.synchronize()
        print(f"B{b} C{c} H{h} W{w}: sampled normwise err {err:.3e}   {t0.elapsed_time(t1)/5*1e3:.1f} us/call")

# bring-up bisect (SA_B200_TC_DEBUG=1: constant column index, 2: operand A channel 0)
mode = int(os.environ.get("SA_B200_TC_DEBUG", "0"))
if mode:
    b, c, h, w = 1, 32, 2, 128
    g = torch.Generator(device=dev).manual_seed(0)
    fl = torch.randn(b, c, h, w, device=dev, generator=g); fr = torch.randn(b, c, h, w, device=dev, generator=g)
    v = torch.ops.sa_b200.corr_volume(fl, fr, "tf32", 1.0).view(b, h, w, w)
    torch.cuda.synchronize()
    if mode == 1:
        want = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, 1, w).expand_as(v)
    else:
        want = fl[:, 0].unsqueeze(-1).expand_as(v)
    print(f"debug mode {mode}: max abs diff {float((v - want).abs().max()):.3e}; v[0,0,:3,:6]=\n{v[0,0,:3,:6]}\nwant=\n{want[0,0,:3,:6]}")
