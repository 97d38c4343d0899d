"""ctypes binding of ``libsclip.so`` (C ABI: ``include/sclip.h``).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libsclip.so")

ABI_VERSION = 2
SCLIP_F32, SCLIP_BF16 = 0, 1
MATH_F16, MATH_F16X3 = 0, 1


class Problem(Structure):
    _fields_ = [
        ("rows_local", c_int32),
        ("rows_global", c_int32),
        ("row_offset", c_int32),
        ("dim", c_int32),
        ("dtype", c_int32),
        ("math", c_int32),
        ("world", c_int32),
        ("parity", c_int32),
    ]


class Layout(Structure):
    _fields_ = [(name, c_uint64) for name in (
        "total_bytes", "xhat", "xhat_lo", "inv_norm", "row_part", "col_part", "tile_ref", "diag", "lse_row",
        "lse_col_local", "lse_col", "row_inv", "col_sum_local", "col_inv", "loss_part", "grad_tiles", "grad_tiles_lo", "dt_part", "dxhat_row", "dxhat_col",
        "col_contrib", "diag_all", "fac_row", "fac_col", "dot_part", "status", "rowterm_part", "sync")] + [("row_tiles", c_int32), ("col_tiles", c_int32), ("ld_g", c_int32), ("reserved", c_int32)]


class SclipError(RuntimeError):
    pass


_PROTOTYPES = {
    "sclip_abi_version": (c_int, []),
    "sclip_last_error": (c_char_p, []),
    "sclip_kernel_launches": (ctypes.c_longlong, []),
    "sclip_plan": (c_int, [POINTER(Problem), POINTER(Layout)]),
    "sclip_prologue": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "sclip_forward_tiles": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_void_p]),
    "sclip_forward_tiles_cols": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_void_p]),
    "sclip_forward_diag": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_void_p]),
    "sclip_backward_scale": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_void_p, c_void_p]),
    "sclip_forward_reduce": (c_int, [POINTER(Problem), c_void_p, c_void_p]),
    "sclip_forward_loss": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_void_p, c_void_p]),
    "sclip_backward_tiles": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_void_p, c_void_p]),
    "sclip_backward_gemms": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_void_p, c_void_p]),
    "sclip_backward_gemms_role": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "sclip_gemm_converts_stash": (c_int, [POINTER(Problem)]),
    "sclip_backward_factors": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_void_p, c_void_p]),
    "sclip_backward_finish": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "sclip_forward": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                              c_void_p]),
    "sclip_backward": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "sclip_push_shards": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "sclip_wait_shards": (c_int, [POINTER(Problem), c_void_p, c_int, c_void_p]),
    "sclip_forward_loss_peers": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_void_p, c_void_p]),
    "sclip_read_status": (c_int, [POINTER(Problem), c_void_p, POINTER(c_int32), c_void_p]),
    "sclip_pull_reduce_cols": (c_int, [POINTER(Problem), c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "sclip_cosine_logits_scratch": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_uint64)]),
    "sclip_cosine_logits": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                    c_int64, c_void_p]),
    "sclip_gemm_f16": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int,
                               c_int, c_float, c_void_p]),
}

EXPORTS = tuple(_PROTOTYPES)
_lib = None


def load() -> ctypes.CDLL:
    """Load the library (building it first if the shared object is absent and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build

    # incremental: a no-op when the digest of the sources matches the built library; raises when the library is
    # missing or stale and nvcc is not available (there is no other implementation to fall back to)
    _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.sclip_abi_version() != ABI_VERSION:
        raise SclipError("libsclip.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().sclip_last_error().decode(errors="replace")
        raise SclipError(f"{what} failed (status {rc}): {msg}")


def plan(problem: Problem) -> Layout:
    lay = Layout()
    check(load().sclip_plan(byref(problem), byref(lay)), "sclip_plan")
    return lay
