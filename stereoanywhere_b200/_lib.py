"""ctypes binding of the C-ABI library (include/sa_b200.h).

There is no CPU fallback: if `lib/libsa_b200.so` is missing and cannot be built, importing the
ops raises.  Signatures below mirror include/sa_b200.h one to one.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libsa_b200.so")
#: development only: load another build of the library (kernel experiments; see tools/build_variant.sh)
_LIB_OVERRIDE = os.environ.get("SA_B200_LIB")

_lib = None

c_fp = C.c_void_p  # device pointers travel as integers (tensor.data_ptr())


class SaError(RuntimeError):
    pass


def _declare(lib):
    i, i64, f, d, vp = C.c_int, C.c_int64, C.c_float, C.c_double, C.c_void_p
    lib.sa_abi_version.restype = i
    lib.sa_abi_version.argtypes = []
    lib.sa_last_error.restype = C.c_char_p
    lib.sa_last_error.argtypes = []
    lib.sa_corr_fp32.restype = i
    lib.sa_corr_fp32.argtypes = [vp, vp, vp, i, i, i, i, i, f, f, vp]
    lib.sa_corr_tf32.restype = i
    lib.sa_corr_tf32.argtypes = [vp, vp, vp, i, i, i, i, i, f, f, vp, vp, d, vp, vp, vp, i64, i64, i64, vp]
    lib.sa_pyramid.restype = i
    lib.sa_pyramid.argtypes = [vp, i64, i, i64, i, vp, vp, vp, i64, i64, i64, vp, vp, d, i, vp, vp]
    lib.sa_lookup.restype = i
    lib.sa_lookup.argtypes = [C.POINTER(vp), C.POINTER(i), C.POINTER(i64), i, i, vp, i64, vp, i, i, i, i, i, vp]
    lib.sa_lookup2.restype = i
    lib.sa_lookup2.argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(i), C.POINTER(i64), C.POINTER(i64), i, i,
                               vp, i64, vp, vp, i, i, i, vp]
    lib.sa_packed_row_floats.restype = i64
    lib.sa_packed_row_floats.argtypes = [i]
    lib.sa_pack_pyramid.restype = i
    lib.sa_pack_pyramid.argtypes = [vp, i64, i, vp, vp, d, i, vp, vp]
    lib.sa_pack_pyramid_normals.restype = i
    lib.sa_pack_pyramid_normals.argtypes = [vp, vp, f, f, i, i, i, i, vp, vp]
    lib.sa_lookup_packed.restype = i
    lib.sa_lookup_packed.argtypes = [vp, vp, i, vp, i64, vp, vp, i, i, i, vp]
    lib.sa_corr_pack_tf32.restype = i
    lib.sa_corr_pack_tf32.argtypes = [vp, vp, i, i, i, i, i, f, f, vp, vp, d, vp, vp]
    lib.sa_corr_pack_tf32_half.restype = i
    lib.sa_corr_pack_tf32_half.argtypes = [vp, vp, i, i, i, i, i, f, f, vp, vp, d, i, vp, vp]
    lib.sa_lookup_packed_half.restype = i
    lib.sa_lookup_packed_half.argtypes = [vp, i, i, vp, vp, f, f, i, vp, i64, vp, vp, i, i, i, vp]
    lib.sa_volume_softargmax.restype = i
    lib.sa_volume_softargmax.argtypes = [vp, i64, i, i, vp, vp, vp]
    lib.sa_volume_entropy_conf.restype = i
    lib.sa_volume_entropy_conf.argtypes = [vp, i64, i, i, vp, vp, vp]
    lib.sa_corr_backward_tf32.restype = i
    lib.sa_corr_backward_tf32.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, f, f, vp]
    lib.sa_lookup_backward.restype = i
    lib.sa_lookup_backward.argtypes = [vp, vp, i64, vp, vp, vp, i, i, i, i, i, i, vp]
    lib.sa_pyramid_backward.restype = i
    lib.sa_pyramid_backward.argtypes = [vp, vp, vp, vp, i, i64, vp, vp, d, i, vp]
    lib.sa_lookup_packed_normals.restype = i
    lib.sa_lookup_packed_normals.argtypes = [vp, vp, vp, f, f, i, vp, i64, vp, vp, i, i, i, vp]
    lib.sa_lookup_packed_factored.restype = i
    lib.sa_lookup_packed_factored.argtypes = [vp, vp, vp, f, f, i, vp, i64, vp, vp, i, i, i, vp]
    lib.sa_stitch_tile.restype = i
    lib.sa_stitch_tile.argtypes = [vp, i, i, i, f, i, i, i, i, vp, f, vp, vp]
    lib.sa_stitch_finish.restype = i
    lib.sa_stitch_finish.argtypes = [vp, vp, i, vp, vp, i, i, i, vp]
    lib.sa_lookup_packed_conv.restype = i
    lib.sa_lookup_packed_conv.argtypes = [vp, vp, i, vp, i64, vp, vp, vp, vp, i, i, i, vp]
    lib.sa_lookup_factored_conv.restype = i
    lib.sa_lookup_factored_conv.argtypes = [vp, vp, vp, f, f, i, vp, i64, vp, vp, vp, vp, i, i, i, vp]
    lib.sa_mono_inputs.restype = i
    lib.sa_mono_inputs.argtypes = [vp, i, i, i, i, f, C.POINTER(f), i, vp, vp, vp, vp]
    lib.sa_weighted_lsq.restype = i
    lib.sa_weighted_lsq.argtypes = [vp, vp, vp, i, i, f, f, vp, vp, vp]
    lib.sa_truncate.restype = i
    lib.sa_truncate.argtypes = [vp, vp, vp, d, vp, i64, i, i, vp]
    lib.sa_masked_volume.restype = i
    lib.sa_masked_volume.argtypes = [vp, vp, vp, f, f, vp, vp, C.POINTER(f), i, vp, i, i, i, i, vp]
    lib.sa_corrupt.restype = i
    lib.sa_corrupt.argtypes = [vp, vp, i, i, vp, f, vp, i, i, i, i, vp]


EXPORTS = [
    "sa_abi_version", "sa_last_error", "sa_corr_fp32", "sa_corr_tf32", "sa_pyramid", "sa_lookup", "sa_lookup2",
    "sa_truncate", "sa_masked_volume", "sa_corrupt", "sa_packed_row_floats", "sa_pack_pyramid", "sa_pack_pyramid_normals", "sa_lookup_packed", "sa_lookup_packed_conv", "sa_lookup_factored_conv", "sa_corr_pack_tf32", "sa_corr_pack_tf32_half", "sa_lookup_packed_half", "sa_mono_inputs", "sa_weighted_lsq", "sa_stitch_tile", "sa_stitch_finish", "sa_lookup_packed_normals", "sa_lookup_packed_factored", "sa_corr_backward_tf32", "sa_lookup_backward", "sa_pyramid_backward", "sa_volume_softargmax", "sa_volume_entropy_conf",
]


def load():
    """Load (building first if the .so is absent and nvcc is present). Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    # Bring the .so up to date with csrc/ first (a no-op when the stamp matches; dlopen caches by path, so a stale
    # library cannot be replaced once it has been opened).  Without nvcc an existing, current .so is used as it is.
    from . import build as _build

    try:
        _build.build()
    except Exception as e:  # no nvcc, compile error ...
        if not os.path.exists(LIB_PATH):
            raise SaError(
                f"stereoanywhere_b200: CUDA library {LIB_PATH} is missing and could not be built ({e}). "
                "There is no CPU fallback; run `python -m stereoanywhere_b200.build`."
            ) from e
        # an existing library is only acceptable when it was built from exactly these sources (box without nvcc);
        # a failed rebuild of EDITED sources must not fall back to kernels that no longer match them
        if not _build.is_current():
            raise SaError(
                f"stereoanywhere_b200: {LIB_PATH} was built from different sources than csrc/ and the rebuild "
                f"failed ({e}); fix the build or run `python -m stereoanywhere_b200.build --force`") from e
    try:
        lib = C.CDLL(_LIB_OVERRIDE or LIB_PATH)
        _declare(lib)
    except (OSError, AttributeError) as stale:
        raise SaError(f"stereoanywhere_b200: {LIB_PATH} does not match this binding ({stale}); "
                      "run `python -m stereoanywhere_b200.build --force`") from stale
    if lib.sa_abi_version() != 1:
        raise SaError("stereoanywhere_b200: ABI version mismatch between _lib.py and libsa_b200.so")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().sa_last_error().decode(errors="replace")
        kind = "CUDA error" if rc > 0 else "argument error"
        raise SaError(f"{what}: {kind} {rc}: {msg}")
