"""The hot path executed by the REFERENCE'S OWN classes (not the restatement in corr_oracle.py).

TEST / BASELINE INFRASTRUCTURE ONLY: used by `bench.py --impl reference`, the `cpu_baseline` leg and tests.
Imports `CorrBlock1D` (models/stereoanywhere/corr.py:75-132) and `truncate_corr_volume_v2`
(utils/utils.py:216-238) unmodified through `ref_shim` - from `/root/reference` in the build container, from the
prebuilt `oracle/_ref/` (recipe `oracle/make_ref.py`) on the GPU box - and runs the call sequence of
stereoanywhere.py:135-136, :203, :253-259, :270-271 for the `use_aggregate_mono_vol=False` wiring, i.e. the same
path `corr_oracle.run_path_cpu` restates and `bench.py` times on the GPU.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import ref_shim


def available() -> bool:
    return ref_shim.reference_available()


def run_path_reference(
    fmap_l: torch.Tensor,
    fmap_r: torch.Tensor,
    normals_l: torch.Tensor,
    normals_r: torch.Tensor,
    coords_seq: Sequence[torch.Tensor],
    trunc: Tuple[torch.Tensor, torch.Tensor, float] | None = None,
    radius: int = 4,
    num_levels: int = 4,
):
    CorrBlock1D, U = ref_shim.import_reference_corr()
    v_s = CorrBlock1D.corr(fmap_l, fmap_r).squeeze(3).unsqueeze(1)                   # stereoanywhere.py:135
    v_m = 1.73 * CorrBlock1D.corr(normals_l, normals_r).squeeze(3).unsqueeze(1)      # :136
    mask = 1
    if trunc is not None:                                                            # :203
        mask = U.truncate_corr_volume_v2(trunc[0], trunc[1], conf_th=None, attenuation_gain=trunc[2]).detach()
    blk_s = CorrBlock1D((mask * v_s).squeeze(1).unsqueeze(3), radius=radius, num_levels=num_levels)   # :253-255
    blk_m = CorrBlock1D(v_m.squeeze(1).unsqueeze(3), radius=radius, num_levels=num_levels)            # :257-259
    outs = None
    for c in coords_seq:                                                             # :270-271
        outs = (blk_s(c), blk_m(c))
    return outs
