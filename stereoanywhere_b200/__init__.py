"""stereoanywhere_b200 - B200-native cost-volume hot path of Stereo Anywhere.

Public surface (mirrors the reference's correlation-block protocol,
models/stereoanywhere/corr.py:75-132 of kei312/stereoanywhere):

    from stereoanywhere_b200 import CorrBlockB200
    vol = CorrBlockB200.corr(fmapL, fmapR)            # [B,H,W,1,W]
    fn  = CorrBlockB200(vol, radius=4, num_levels=4)  # pyramid
    feat = fn(coords)                                 # [B,36,H,W] per GRU iteration

Everything executes in hand-written sm_100a CUDA kernels (stereoanywhere_b200/csrc) reached through
the C ABI of include/sa_b200.h.  Importing this package loads the library and fails loudly if it is
absent and cannot be built - there is no CPU fallback.
"""
from . import _lib

_lib.load()  # raises if the CUDA library is unavailable

from . import ops  # noqa: E402  (registers torch.ops.sa_b200.*)
from .corr import (  # noqa: E402
    CorrBlockB200,
    corrupt_volume,
    lookup_pair_convc1,
    masked_mono_volume,
    masked_volume,
    truncation_mask,
)

from .reductions import (  # noqa: E402
    estimate_confidences,
    estimate_disparities,
    estimate_left_confidence,
    estimate_left_disparity,
    estimate_right_confidence,
    estimate_right_disparity,
)

__all__ = [
    "estimate_disparities",
    "estimate_confidences",
    "estimate_left_disparity",
    "estimate_right_disparity",
    "estimate_left_confidence",
    "estimate_right_confidence",
    "CorrBlockB200",
    "truncation_mask",
    "masked_volume",
    "masked_mono_volume",
    "corrupt_volume",
    "lookup_pair_convc1",
    "ops",
]
__version__ = "0.1.0"
