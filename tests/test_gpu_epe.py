"""End-to-end gate of the north star: after 32 GRU iterations the disparity driven by the B200
block must stay within 0.05 px EPE of the one driven by the reference path.

The reference's update block (models/stereoanywhere/update.py) cannot travel to the GPU box, so the
loop of stereoanywhere.py:261-280 is reproduced around a random-init surrogate with the same data
flow: shared 1x1 `convc1` on both lookups (update.py:74,81-84), a flow branch, a ConvGRU and a delta
head whose y component is zeroed (stereoanywhere.py:277).  Both runs share weights, features and
initial coords; one uses the oracle's ATen op sequence (on the GPU), the other CorrBlockB200
(TF32 tensor-core correlation, truncation fused, packed lookup).
"""
import pytest
import torch
import torch.nn as nn

from oracle import corr_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class SurrogateUpdate(nn.Module):
    def __init__(self, hidden=64):
        super().__init__()
        self.convc1 = nn.Conv2d(36, 64, 1)
        self.convf1 = nn.Conv2d(2, 32, 7, padding=3)
        self.conv = nn.Conv2d(64 + 64 + 32, 96, 3, padding=1)
        self.convz = nn.Conv2d(hidden + 96, hidden, 3, padding=1)
        self.convr = nn.Conv2d(hidden + 96, hidden, 3, padding=1)
        self.convq = nn.Conv2d(hidden + 96, hidden, 3, padding=1)
        self.head = nn.Sequential(nn.Conv2d(hidden, 64, 3, padding=1), nn.ReLU(), nn.Conv2d(64, 2, 3, padding=1))

    def forward(self, net, stereo_corr, mono_corr, flow):
        cs = torch.relu(self.convc1(stereo_corr))
        cm = torch.relu(self.convc1(mono_corr))
        fl = torch.relu(self.convf1(flow))
        x = torch.relu(self.conv(torch.cat([cs, cm, fl], 1)))
        hx = torch.cat([net, x], 1)
        z = torch.sigmoid(self.convz(hx))
        r = torch.sigmoid(self.convr(hx))
        q = torch.tanh(self.convq(torch.cat([r * net, x], 1)))
        net = (1 - z) * net + z * q
        return net, self.head(net)


def run_loop(update, stereo_fn, mono_fn, coords0, init_disp, iters=32):
    coords1 = coords0.clone()
    coords1[:, :1] = coords0[:, :1] - init_disp
    net = torch.zeros(coords0.shape[0], 64, coords0.shape[2], coords0.shape[3], device=coords0.device)
    for _ in range(iters):
        s, m = stereo_fn(coords1), mono_fn(coords1)
        net, delta = update(net, s, m, coords1 - coords0)
        delta = delta.clone()
        delta[:, 1] = 0.0
        coords1 = coords1 + delta
    return (coords1 - coords0)[:, :1]


@pytest.mark.parametrize("b,c,h,w", [(1, 256, 96, 128), (2, 128, 48, 312)])
def test_epe_after_32_iterations(b, c, h, w):
    import stereoanywhere_b200 as sa

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    update = SurrogateUpdate().to(DEV).eval()
    g = torch.Generator(device=DEV).manual_seed(1)
    fl = torch.randn(b, c, h, w, device=DEV, generator=g)
    fr = torch.roll(fl, -6, dims=3) + 0.3 * torch.randn(b, c, h, w, device=DEV, generator=g)
    nl = torch.nn.functional.normalize(torch.randn(b, 3, h, w, device=DEV, generator=g), dim=1)
    nr = torch.roll(nl, -6, dims=3)
    tdisp = torch.rand(b, 1, h, w, device=DEV, generator=g) * 12
    tconf = torch.rand(b, 1, h, w, device=DEV, generator=g)
    x = torch.arange(w, device=DEV, dtype=torch.float32).view(1, 1, 1, w).expand(b, 1, h, w)
    y = torch.arange(h, device=DEV, dtype=torch.float32).view(1, 1, h, 1).expand(b, 1, h, w)
    coords0 = torch.cat([x, y], 1).contiguous()
    init = torch.rand(b, 1, h, w, device=DEV, generator=g) * 8

    with torch.no_grad():
        # reference path: ATen op sequence of the oracle, fp32, on the GPU
        vs = O.aten_corr_volume(fl, fr).squeeze(3).unsqueeze(1)
        vm = O.aten_mono_corr_volume(nl, nr).squeeze(3).unsqueeze(1)
        t = O.aten_truncation_mask(tdisp, tconf, 0.9)
        ref_s = O.OracleCorrBlock((t * vs).squeeze(1).unsqueeze(3), num_levels=4, radius=4)
        ref_m = O.OracleCorrBlock(vm.squeeze(1).unsqueeze(3), num_levels=4, radius=4)
        d_ref = run_loop(update, ref_s, ref_m, coords0, init)

        B = sa.CorrBlockB200
        assert B.precision == "tf32"
        ours_s = B(B.corr(fl, fr), num_levels=4, radius=4, truncate=(tdisp, tconf, 0.9))
        ours_m = B(B.mono_corr(nl, nr), num_levels=4, radius=4)
        d_b200 = run_loop(update, ours_s, ours_m, coords0, init)

    assert torch.isfinite(d_ref).all() and torch.isfinite(d_b200).all()
    moved = float((d_ref + init).abs().mean())
    epe = float((d_b200 - d_ref).abs().mean())
    worst = float((d_b200 - d_ref).abs().max())
    print(f"EPE after 32 iters: {epe:.2e} px (max {worst:.2e}); mean |update| {moved:.3f} px")
    assert moved > 1e-3, "surrogate did not move the disparity: the test would be vacuous"
    assert epe < 0.05
