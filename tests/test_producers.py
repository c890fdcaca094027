"""the sync-free producer chain (stereoanywhere_b200/producers.py) against fixtures generated from
the reference (tests/golden/producers.npz).  Plain PyTorch: runs on the CPU here and on the GPU box alike."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "producers.npz")
T = torch.from_numpy


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN))


@pytest.fixture(scope="module")
def P():
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "sa_producers", os.path.join(os.path.dirname(os.path.dirname(__file__)), "stereoanywhere_b200", "producers.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # the module is pure torch: no CUDA library needed to import it
    return mod


@pytest.mark.parametrize("n", [4, 8, 16])
def test_generate_masks_bit_exact(P, g, n):
    m = P.generate_masks(T(g["gm_mde"]), N=n)
    assert m.dtype == torch.float16 and np.array_equal(m.numpy(), g[f"gm_masks{n}"])
    assert float(m[0, :, 0, 0].sum()) == 0.0  # mde == 1.0 is in no bin (utils/utils.py:51)


def test_estimate_normals(P, g):
    n = P.estimate_normals(T(g["en_depth"]), normal_gain=20 / 10)
    assert np.abs(n.numpy() - g["en_normals"]).max() < 1e-6
    assert np.abs(np.linalg.norm(n.numpy(), axis=1) - 1).max() < 1e-6


def test_weighted_lsq_batched(P, g):
    sc, sh = P.weighted_lsq(T(g["wl_mono"]), T(g["wl_disp"]), T(g["wl_conf"]))
    assert sc.shape == g["wl_scale"].shape and sh.shape == g["wl_shift"].shape
    assert np.abs(sc.numpy() - g["wl_scale"]).max() < 1e-4 * np.abs(g["wl_scale"]).max()
    assert np.abs(sh.numpy() - g["wl_shift"]).max() < 1e-4 * max(1.0, np.abs(g["wl_shift"]).max())


def test_lowres_matches_interpolate(P):
    x = torch.rand(1, 1, 32, 64)
    assert P.lowres(x).shape == (1, 1, 8, 16)
