"""ctypes view of include/stencils_b200.h: enums, the sweep descriptor and the library loader.

The product path loads `lib/libstencils_b200.so` (hand-written sm_100a kernels behind a C ABI) and
fails loudly when it is missing — there is no CPU or PyTorch fallback (BASELINE.json north_star).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SB200_LIB: another build of the same library (A/B runs of kernel variants on the GPU box, tools/build_variant.sh)
LIB_PATH = os.environ.get("SB200_LIB") or os.path.join(HERE, "lib", "libstencils_b200.so")

# ---- enums (include/stencils_b200.h) ----
OK, EINVAL, EUNSUPPORTED, ESIZE, ECUDA, ENOMEM = range(6)
BOOL, U8, I32, I64, F32, F64 = range(6)
REMOVE, WRAP, REFLECT, USE = range(4)
(WINDOW, MOORE, VONNEUMANN, CROSS, ANGLEDCROSS, FORWARDSLASH, BACKSLASH, CIRCLE, VERTICAL, HORIZONTAL,
 DIAMOND, ANNULUS, CARDINAL, ORDINAL) = range(14)
SUM, MEAN, MIN, MAX, KERNELDOT, LIFE, DIFFUSION = range(7)
OP_ADD, OP_MAX, OP_MIN = range(3)
SCATTER_WEIGHTS, SCATTER_CENTER_WEIGHTS = range(2)
FLAG_FORCE_GENERIC, FLAG_ZERO_DEST, FLAG_NO_TMA, FLAG_CELLS_01, FLAG_DOUBLE_STEP, FLAG_QUAD_STEP, FLAG_OCT_STEP, FLAG_ALLOW_FMA = 1, 2, 4, 8, 16, 32, 64, 128
FLAG_STEP_MASK = 112
FLAG_SRC_BITS, FLAG_DST_BITS = 256, 512


def flag_gens(n):
    """SB200_FLAG_GENS(n): the *_STEP bits for n generations per launch (1 <= n <= 8)."""
    return {1: 0, 2: 16, 4: 32, 8: 64, 3: 48, 5: 80, 6: 96, 7: 112}[n]


MAX_OFFSETS = 1024
# default of SB200_DIFFUSION_DOUBLE_STEP (two diffusion steps per launch in iterated runs); must agree with
# kDiffusionDoubleStepDefault in csrc/api.cu
DIFFUSION_DOUBLE_STEP_DEFAULT = "1"

ELTYPE_OF_DTYPE = {
    np.dtype(np.bool_): BOOL, np.dtype(np.uint8): U8, np.dtype(np.int32): I32,
    np.dtype(np.int64): I64, np.dtype(np.float32): F32, np.dtype(np.float64): F64,
}
DTYPE_OF_ELTYPE = {v: k for k, v in ELTYPE_OF_DTYPE.items()}


class Desc(C.Structure):
    """struct sb200_desc"""
    _fields_ = [
        ("struct_size", C.c_int32), ("ndim", C.c_int32),
        ("size", C.c_int64 * 3), ("src_ext", C.c_int64 * 3), ("dst_ext", C.c_int64 * 3),
        ("src_off", C.c_int32 * 3), ("dst_off", C.c_int32 * 3), ("boundary", C.c_int32 * 3),
        ("eltype", C.c_int32), ("out_eltype", C.c_int32),
        ("padval_bits", C.c_uint64),
        ("radius", C.c_int32), ("noffsets", C.c_int32),
        ("offsets_host", C.c_void_p),
        ("reducer", C.c_int32), ("scatter_op", C.c_int32), ("scatter_rule", C.c_int32),
        ("born_mask", C.c_uint32), ("survive_mask", C.c_uint32), ("reserved0", C.c_int32),
        ("weights_host", C.c_void_p),
        ("alpha", C.c_double),
        ("region_lo", C.c_int64 * 3), ("region_hi", C.c_int64 * 3),
        ("flags", C.c_int32), ("reserved1", C.c_int32),
        ("mirror_parent", C.c_void_p), ("mirror_lo", C.c_int64), ("mirror_hi", C.c_int64),
    ]


class SB200Error(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"libstencils_b200 status {status}: {msg}")
        self.status = status


class Term(C.Structure):
    """sb200_term (include/stencils_b200.h)"""
    _fields_ = [("desc", C.POINTER(Desc)), ("src_parent", C.c_void_p), ("has_coef", C.c_int32), ("reserved", C.c_int32),
                ("coef", C.c_double)]


class SlabOp(C.Structure):
    """sb200_slab_op (include/stencils_b200.h)"""
    _fields_ = [("kind", C.c_int32), ("gens", C.c_int32), ("mirror", C.c_int32), ("buf", C.c_int32), ("async_", C.c_int32),
                ("first", C.c_int32), ("lo", C.c_int64), ("hi", C.c_int64)]


SLAB_SWEEP, SLAB_PUSH, SLAB_SIGNAL, SLAB_PULL, SLAB_JOIN, SLAB_ENDFILL, SLAB_SWAP = range(1, 8)
SLAB_CUR, SLAB_NXT = 0, 1
SLAB_MIRROR_DOWN, SLAB_MIRROR_UP = 1, 2
PLAN_OVERLAP_OFF, PLAN_OVERLAP_ON, PLAN_SINGLE_STEP, PLAN_FLAGS_SYNC = 1, 2, 4, 8


class ArgumentError(ValueError):
    """Julia's ArgumentError: unsupported user function / eltype / shape, size mismatch."""


_lib = None

_SIGS = {
    "sb200_version": (C.c_int32, []),
    "sb200_last_error": (C.c_char_p, []),
    "sb200_last_kernel": (C.c_char_p, []),
    "sb200_launch_count": (C.c_int64, [C.c_int32]),
    "sb200_stencil_offsets": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                          C.POINTER(C.c_int32)]),
    "sb200_out_eltype": (C.c_int32, [C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "sb200_sizeof": (C.c_size_t, [C.c_int32]),
    "sb200_gather": (C.c_int32, [C.POINTER(Desc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "sb200_update_halo": (C.c_int32, [C.POINTER(Desc), C.c_void_p, C.c_void_p]),
    "sb200_scatter": (C.c_int32, [C.POINTER(Desc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "sb200_gather_multi": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sb200_iterate": (C.c_int32, [C.POINTER(Desc), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "sb200_gather_host": (C.c_int32, [C.POINTER(Desc), C.c_void_p, C.c_void_p]),
    "sb200_iterate_host": (C.c_int32, [C.POINTER(Desc), C.c_void_p, C.c_int32]),
    "sb200_device_count": (C.c_int32, [C.POINTER(C.c_int32)]),
    "sb200_set_device": (C.c_int32, [C.c_int32]),
    "sb200_malloc": (C.c_int32, [C.POINTER(C.c_void_p), C.c_size_t]),
    "sb200_free": (C.c_int32, [C.c_void_p]),
    "sb200_malloc_host": (C.c_int32, [C.POINTER(C.c_void_p), C.c_size_t]),
    "sb200_free_host": (C.c_int32, [C.c_void_p]),
    "sb200_memcpy_h2d": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sb200_memcpy_d2h": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sb200_memcpy_d2d": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sb200_memset": (C.c_int32, [C.c_void_p, C.c_int32, C.c_size_t, C.c_void_p]),
    "sb200_stream_sync": (C.c_int32, [C.c_void_p]),
    "sb200_ipc_export": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "sb200_ipc_import": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "sb200_ipc_close": (C.c_int32, [C.c_void_p]),
    "sb200_push_planes": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p]),
    "sb200_signal_flag": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "sb200_wait_flag": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "sb200_shutdown": (C.c_int32, []),
    "sb200_debug_split_steps": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "sb200_plan_create": (C.c_int32, [C.POINTER(Desc), C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "sb200_plan_create_rank": (C.c_int32, [C.POINTER(Desc), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "sb200_plan_ipc_handle": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "sb200_plan_connect": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "sb200_plan_nslabs": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
    "sb200_plan_slab": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_void_p)]),
    "sb200_plan_mark_dirty": (C.c_int32, [C.c_void_p]),
    "sb200_plan_load_host": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "sb200_plan_store_host": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "sb200_plan_iterate": (C.c_int32, [C.c_void_p, C.c_int32]),
    "sb200_plan_sync": (C.c_int32, [C.c_void_p]),
    "sb200_plan_iterate_timed": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(C.c_float)]),
    "sb200_plan_stats": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64)]),
    "sb200_plan_destroy": (C.c_int32, [C.c_void_p]),
    "sb200_slab_schedule": (C.c_int32, [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.POINTER(SlabOp), C.c_int32, C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32)]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)


def lib() -> C.CDLL:
    """Load libstencils_b200.so (built in-tree by __graft_entry__.build() / csrc/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback for the stencil sweep.")
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int) -> None:
    if status == OK:
        return
    msg = lib().sb200_last_error().decode("utf-8", "replace")
    if status in (EUNSUPPORTED, ESIZE, EINVAL):
        raise ArgumentError(f"[sb200 status {status}] {msg}")
    raise SB200Error(status, msg)
