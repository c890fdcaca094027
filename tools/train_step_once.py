#!/usr/bin/env python
"""Forward + backward of the path at the reference's training shape (train.py: batch 2, 320x640 crops ->
80x160 quarter resolution, C = 256, 12 iterations): CorrBlockB200 with its CUDA backward against
the same op sequence in ATen (einsum / avg_pool2d / grid_sample + autograd) on the same GPU."""
import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stereoanywhere_b200 as sa

dev = torch.device("cuda:0")
b, c, h, w, iters = 2, 256, 80, 160, 12
g = torch.Generator(device=dev).manual_seed(0)
fl = torch.randn(b, c, h, w, device=dev, generator=g); fr = torch.randn(b, c, h, w, device=dev, generator=g)
mono = torch.randn(b, h, w, 1, w, device=dev, generator=g)
x = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, 1, w).expand(b, 1, h, w)
coords = [torch.cat([x - torch.rand(b, 1, h, w, device=dev, generator=g) * (w / 4), torch.zeros(b, 1, h, w, device=dev)], 1) for _ in range(iters)]
wts = [torch.randn(b, 36, h, w, device=dev, generator=g) for _ in range(iters)]


class AtenBlock:  # the reference's op sequence (corr.py:76-132, utils/utils.py:19-35) without the host sync
    def __init__(self, vol, num_levels=4, radius=4):
        bb, hh, w1, _, w2 = vol.shape
        v = vol.reshape(bb * hh * w1, 1, 1, w2)
        self.pyr, self.r, self.n = [v], radius, num_levels
        for _ in range(num_levels):
            v = F.avg_pool2d(v, [1, 2], stride=[1, 2]); self.pyr.append(v)
    @staticmethod
    def corr(a, bm):
        bb, d, hh, ww = a.shape
        return (torch.einsum("aijk,aijh->ajkh", a, bm).reshape(bb, hh, ww, 1, ww) / torch.sqrt(torch.tensor(float(d)))).contiguous()
    def __call__(self, cds):
        bb, _, hh, ww = cds.shape
        cx = cds[:, :1].permute(0, 2, 3, 1)
        out = []
        for i in range(self.n):
            dx = torch.linspace(-self.r, self.r, 2 * self.r + 1, device=cds.device).view(2 * self.r + 1, 1)
            x0 = dx + cx.reshape(bb * hh * ww, 1, 1, 1) / 2 ** i
            wi = self.pyr[i].shape[-1]
            grid = torch.cat([2 * x0 / (wi - 1) - 1, torch.zeros_like(x0)], -1)
            out.append(F.grid_sample(self.pyr[i], grid, align_corners=True).view(bb, hh, ww, -1))
        return torch.cat(out, -1).permute(0, 3, 1, 2).contiguous()


def step(block_cls):
    f2, f3, mv = fl.clone().requires_grad_(True), fr.clone().requires_grad_(True), mono.clone().requires_grad_(True)
    sfn = block_cls(block_cls.corr(f2, f3), num_levels=4, radius=4)
    mfn = block_cls(mv, num_levels=4, radius=4)
    loss = 0
    for k in range(iters):
        loss = loss + (sfn(coords[k]) * wts[k]).sum() + (mfn(coords[k]) * wts[k]).sum()
    return torch.autograd.grad(loss, (f2, f3, mv))


def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out

sa.CorrBlockB200.precision = "fp32"
t_b, gb = timeit(lambda: step(sa.CorrBlockB200))
t_a, ga = timeit(lambda: step(AtenBlock))
err = [float((p - q).abs().max() / q.abs().max()) for p, q in zip(gb, ga)]
print(f"training-shape fwd+bwd (b={b}, {h}x{w}, C={c}, {iters} iters, both blocks): CorrBlockB200 {t_b:.2f} ms | ATen {t_a:.2f} ms | "
      f"normwise grad diff fmapL {err[0]:.1e} fmapR {err[1]:.1e} mono volume {err[2]:.1e}")
