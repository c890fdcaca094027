#!/usr/bin/env python
"""The reference's op sequence for the path (einsum, avg_pool2d, grid_sample, mask products) run by ATen ON THE
B200 - the "real bar" of SURVEY 8d - next to the CUDA path, same inputs, same step definition as bench.py
(2 x corr + truncation + 2 x pyramid + 32 x (stereo + mono lookup)).  `--sync 1` keeps the reference's
`assert torch.unique(ygrid).numel() == 1` (utils/utils.py:26), a device->host sync per level per lookup."""
import argparse, os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import stereoanywhere_b200 as sa

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
ap.add_argument("--sync", type=int, default=1)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
b, c, h, w = bench.WORKLOADS[args.workload]
dev = torch.device("cuda:0")
_, d = bench.make_inputs(b, c, h, w, dev, seed=0)
ITERS = bench.ITERS


def corr(a, bm):  # corr.py:117-132
    bb, dd, hh, ww = a.shape
    v = torch.einsum("aijk,aijh->ajkh", a, bm).reshape(bb, hh, ww, 1, bm.shape[3]).contiguous()
    return v / torch.sqrt(torch.tensor(dd).float())


class AtenBlock:  # corr.py:76-115 + utils/utils.py:19-35
    def __init__(self, vol, num_levels=4, radius=4):
        bb, hh, w1, _, w2 = vol.shape
        v = vol.reshape(bb * hh * w1, 1, 1, w2)
        self.pyr, self.r, self.n = [v], radius, num_levels
        for _ in range(num_levels):
            v = F.avg_pool2d(v, [1, 2], stride=[1, 2]); self.pyr.append(v)

    def __call__(self, cds):
        bb, _, hh, ww = cds.shape
        cx = cds[:, :1].permute(0, 2, 3, 1)
        out = []
        for i in range(self.n):
            dx = torch.linspace(-self.r, self.r, 2 * self.r + 1).view(2 * self.r + 1, 1).to(cds.device)
            x0 = dx + cx.reshape(bb * hh * ww, 1, 1, 1) / 2 ** i
            y0 = torch.zeros_like(x0)
            if args.sync:
                assert torch.unique(y0).numel() == 1
            wi = self.pyr[i].shape[-1]
            grid = torch.cat([2 * x0 / (wi - 1) - 1, y0], -1)
            out.append(F.grid_sample(self.pyr[i], grid, align_corners=True).view(bb, hh, ww, -1))
        return torch.cat(out, -1).permute(0, 3, 1, 2).contiguous().float()


def trunc_mask(disp, conf, g):  # utils/utils.py:216-238
    ww = disp.shape[-1]
    cols = torch.arange(ww, device=disp.device, dtype=disp.dtype)
    arg = (cols.view(1, 1, 1, ww, 1) - disp.unsqueeze(4)) - cols.view(1, 1, 1, 1, ww)
    cc = conf.unsqueeze(4)
    return 1 * (1 - cc) + cc * (torch.sigmoid(arg) * (1 - g) + g)


def aten_step():
    vs = corr(d["fl"], d["fr"]).squeeze(3).unsqueeze(1)
    vm = 1.73 * corr(d["nl"], d["nr"]).squeeze(3).unsqueeze(1)
    t = trunc_mask(d["tdisp"], d["tconf"], 0.9)
    fs = AtenBlock((t * vs).squeeze(1).unsqueeze(3)); fm = AtenBlock(vm.squeeze(1).unsqueeze(3))
    cds = d["coords0"]
    for _ in range(ITERS):
        s, m = fs(cds), fm(cds)
        cds = cds + d["delta"]
    return s, m


def b200_step():
    B = sa.CorrBlockB200
    fs = B.from_features(d["fl"], d["fr"], truncate=(d["tdisp"], d["tconf"], 0.9)); fm = B.from_normals(d["nl"], d["nr"])
    cds = d["coords0"]
    for _ in range(ITERS):
        s, m = B.lookup_pair(fs, fm, cds)
        cds = cds + d["delta"]
    return s, m


def timeit(fn, n):
    with torch.no_grad():
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): out = fn()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out

t_a, (sa_, ma_) = timeit(aten_step, args.steps)
t_b, (sb_, mb_) = timeit(b200_step, args.steps * 3)
es = float((sb_ - sa_).abs().max() / sa_.abs().max()); em = float((mb_ - ma_).abs().max() / ma_.abs().max())
print(f"{args.workload}: ATen on B200 (sync per level = {args.sync}) {t_a:.2f} ms/step = {b / t_a * 1e3:.0f} pairs/s | CorrBlockB200 eager {t_b:.2f} ms/step = "
      f"{b / t_b * 1e3:.0f} pairs/s | x{t_a / t_b:.1f} | normwise diff stereo {es:.1e} mono {em:.1e}")
