"""Soft-argmax disparities and entropy confidences of an aggregated volume.

Drop-ins for the reference's `estimate_left_disparity`, `estimate_right_disparity`,
`estimate_left_confidence`, `estimate_right_confidence` (models/stereoanywhere/utils/utils.py:112-170), which
`StereoAnywhere.forward` calls on the hourglass outputs (stereoanywhere.py:174-177).  The reference makes four
separate softmax passes over `[B,H,W2,W3]`; here each volume is read from HBM ONCE for both directions
(`csrc/volume_reduce.cu`).  `estimate_disparities` / `estimate_confidences` return the (left, right) pair of one
launch; the four reference-named functions are provided for call-site compatibility (each launches the pair
kernel and returns its half).  Forward only, CUDA only.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import ops  # noqa: F401  (registers torch.ops.sa_b200.*)
from .corr import _no_grad_check

_OPS = torch.ops.sa_b200


def _f32(vol: torch.Tensor) -> torch.Tensor:
    _no_grad_check(vol)
    return vol if vol.dtype == torch.float32 else vol.float()


def estimate_disparities(corr_volume: torch.Tensor, vol_pad: Sequence[int] = (0, 0)) -> Tuple[torch.Tensor, torch.Tensor]:
    """(`estimate_left_disparity(v)`, `estimate_right_disparity(v)`) with one read of `v` = [B,1,H,W2,W3]."""
    dt = corr_volume.dtype
    left, right = _OPS.volume_softargmax(_f32(corr_volume))
    w2, w3 = left.shape[-1], right.shape[-1]
    left, right = left[..., vol_pad[0]: w2 - vol_pad[1]], right[..., vol_pad[0]: w3 - vol_pad[1]]
    return (left, right) if dt == torch.float32 else (left.to(dt), right.to(dt))


def estimate_confidences(corr_volume: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(`estimate_left_confidence(v)`, `estimate_right_confidence(v)`) with one read of `v`."""
    dt = corr_volume.dtype
    left, right = _OPS.volume_entropy_conf(_f32(corr_volume))
    return (left, right) if dt == torch.float32 else (left.to(dt), right.to(dt))


def estimate_left_disparity(corr_volume, vol_pad=(0, 0)):
    return estimate_disparities(corr_volume, vol_pad)[0]


def estimate_right_disparity(corr_volume, vol_pad=(0, 0)):
    return estimate_disparities(corr_volume, vol_pad)[1]


def estimate_left_confidence(corr_volume, logsumexp_eps=1e-3):
    return estimate_confidences(corr_volume)[0]


def estimate_right_confidence(corr_volume, logsumexp_eps=1e-3):
    return estimate_confidences(corr_volume)[1]
