"""SURVEY 8f-3 on the GPU: the producer kernels (csrc/producers.cu) against fixtures generated from the reference
(tests/golden/producers.npz, producers_chain.npz) and against the batched host mirror (producers.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
T = torch.from_numpy


@pytest.fixture(scope="module")
def chain():
    return dict(np.load(os.path.join(GOLDEN, "producers_chain.npz")))


@pytest.fixture(scope="module")
def small():
    return dict(np.load(os.path.join(GOLDEN, "producers.npz")))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_mono_inputs_matches_the_reference_chain(chain, tag):
    """resize -> normals -> depth bins in one launch vs F.interpolate / estimate_normals / generate_masks of the
    reference (stereoanywhere.py:109-114,138-139)."""
    from stereoanywhere_b200 import producers as P

    g = chain
    mde = T(g[f"{tag}_mde"]).to(DEV)
    low, normals, masks = P.mono_inputs(mde, n_downsample=2, normal_gain=float(g[f"{tag}_gain"][0]), n_bins=8)
    assert low.shape == g[f"{tag}_low"].shape and normals.shape == g[f"{tag}_normals"].shape
    assert masks.dtype == torch.float16 and masks.shape == g[f"{tag}_masks8"].shape
    assert np.abs(low.cpu().numpy() - g[f"{tag}_low"]).max() < 1e-6
    assert np.abs(normals.cpu().numpy() - g[f"{tag}_normals"]).max() < 5e-6
    assert np.abs(np.linalg.norm(normals.cpu().numpy(), axis=1) - 1).max() < 1e-6
    # bins: exactly generate_masks of OUR resized map; against the fixture they may only differ where the resized
    # depth sits within rounding of a bin edge
    assert torch.equal(masks, P.generate_masks(low, N=8))
    diff = (masks.cpu().numpy() != g[f"{tag}_masks8"]).any(axis=1)
    edge_dist = np.abs(g[f"{tag}_low"][:, 0, :, :, None] - np.arange(9) / 8).min(-1)
    assert (edge_dist[diff] < 1e-6).all()
    assert float(masks[0, :, 0, 0].sum()) == 0.0        # mde == 1.0 falls in no bin (utils/utils.py:51)
    # default gain = W_lowres / 10 (stereoanywhere.py:46,113), and the maskless form
    low2, normals2, none = P.mono_inputs(mde, with_masks=False)
    assert none is None and torch.equal(low2, low)
    assert torch.allclose(normals2, P.estimate_normals(low, (mde.shape[-1] // 4) / 10), atol=5e-6)


def test_mono_inputs_odd_sizes_vs_host_mirror():
    from stereoanywhere_b200 import producers as P

    gen = torch.Generator().manual_seed(4)
    for (b, h, w, nd) in [(1, 37, 53, 2), (3, 384, 1248, 2), (2, 64, 64, 0), (1, 40, 72, 3)]:
        mde = torch.rand(b, 1, h, w, generator=gen).to(DEV)
        low, normals, masks = P.mono_inputs(mde, n_downsample=nd, normal_gain=3.7, n_bins=4)
        want_low = P.lowres(mde, nd) if nd else mde
        assert low.shape == want_low.shape and float((low - want_low).abs().max()) < 1e-6
        assert float((normals - P.estimate_normals(low, 3.7)).abs().max()) < 5e-6
        assert torch.equal(masks, P.generate_masks(low, N=4))


@pytest.mark.parametrize("which", ["small", "chain"])
def test_weighted_lsq_kernel_vs_reference(small, chain, which):
    from stereoanywhere_b200 import producers as P

    g = small if which == "small" else chain
    mono, disp, conf = (T(g[k]).to(DEV) for k in ("wl_mono", "wl_disp", "wl_conf"))
    sc, sh = P.weighted_lsq_b200(mono, disp, conf)
    assert sc.shape == g["wl_scale"].shape and sh.shape == g["wl_shift"].shape
    # the reference solves by QR in fp32 (torch.linalg.lstsq), the kernel by normal equations in double
    assert np.abs(sc.cpu().numpy() - g["wl_scale"]).max() < 1e-4 * np.abs(g["wl_scale"]).max()
    assert np.abs(sh.cpu().numpy() - g["wl_shift"]).max() < 1e-4 * max(1.0, np.abs(g["wl_shift"]).max())
    # the host mirror (torch.quantile + double sums): same window, same sums
    sc2, sh2 = P.weighted_lsq(mono, disp, conf)
    assert float((sc - sc2).abs().max()) < 1e-6 * float(sc2.abs().max())
    assert float((sh - sh2).abs().max()) < 1e-6 * max(1.0, float(sh2.abs().max()))


def test_weighted_lsq_kernel_quantile_window_is_exact():
    """The radix select returns torch.quantile's order statistics exactly (ties at the relu zeros included): feed the
    kernel a problem whose fit is very sensitive to the window and compare with the mirror at model size, batch 64."""
    from stereoanywhere_b200 import producers as P

    gen = torch.Generator().manual_seed(8)
    b, n = 64, 2 * 136 * 240
    mono = torch.rand(b, 2, 136, 240, generator=gen).to(DEV)
    disp = (40 * mono - 12 + 3 * torch.randn(b, 2, 136, 240, generator=gen).to(DEV))
    conf = torch.rand(b, 2, 136, 240, generator=gen).to(DEV)
    sc, sh = P.weighted_lsq_b200(mono, disp, conf)
    sc2, sh2 = P.weighted_lsq(mono, disp, conf)
    assert float(((sc - sc2) / sc2).abs().max()) < 1e-6 and float((sh - sh2).abs().max()) < 1e-5 * float(sh2.abs().max())
