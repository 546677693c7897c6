"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does (its sweeps fail loudly without the CUDA library).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import stencils_b200  # noqa: F401  (registers the package under its importable name)
from stencils_b200 import _abi as A
from stencils_b200._desc import DescHandle, build_desc

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in ("stencils_oracle.c", "sweep_body.inc", "Makefile")]
    srcs.append(os.path.join(HERE, "..", "include", "stencils_b200.h"))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB)
        l.orc_stencil_offsets.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]
        l.orc_out_eltype.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int32)]
        for name in ("orc_gather", "orc_scatter"):
            getattr(l, name).argtypes = [C.POINTER(A.Desc), C.c_void_p, C.c_void_p]
        l.orc_update_halo.argtypes = [C.POINTER(A.Desc), C.c_void_p]
        l.orc_iterate.argtypes = [C.POINTER(A.Desc), C.c_void_p, C.c_void_p, C.c_int]
        l.orc_set_num_threads.argtypes = [C.c_int]
        _lib = l
    return _lib


class OracleError(RuntimeError):
    def __init__(self, status):
        super().__init__(f"oracle status {status}")
        self.status = status


def _ck(rc):
    if rc:
        raise OracleError(rc)


def offsets(shape: int, radius: int, ndim: int = 2, inner_radius: int = 0) -> list[tuple[int, ...]]:
    n = C.c_int32()
    _ck(lib().orc_stencil_offsets(shape, radius, inner_radius, ndim, None, 0, C.byref(n)))
    buf = np.zeros((n.value, 3), dtype=np.int32)
    _ck(lib().orc_stencil_offsets(shape, radius, inner_radius, ndim, buf.ctypes.data, n.value, C.byref(n)))
    return [tuple(int(v) for v in row[:ndim]) for row in buf]


def out_eltype(reducer: int, eltype: int) -> int:
    o = C.c_int32()
    _ck(lib().orc_out_eltype(reducer, eltype, C.byref(o)))
    return o.value


def threads() -> int:
    return lib().orc_num_threads()


def set_threads(n: int) -> None:
    lib().orc_set_num_threads(n)


def _fptr(a: np.ndarray):
    assert a.flags.f_contiguous, "oracle buffers are column-major (Julia layout)"
    return a.ctypes.data


def gather(h: DescHandle, src_parent: np.ndarray, dst_parent: np.ndarray | None = None) -> np.ndarray:
    """orc_gather on column-major numpy parents. Allocates the dest parent when not given."""
    d = h.desc
    if dst_parent is None:
        dst_parent = np.zeros(tuple(d.dst_ext[a] for a in range(d.ndim)), dtype=A.DTYPE_OF_ELTYPE[d.out_eltype],
                              order="F")
    _ck(lib().orc_gather(h.ptr(), _fptr(src_parent), _fptr(dst_parent)))
    return dst_parent


def update_halo(h: DescHandle, parent: np.ndarray) -> np.ndarray:
    _ck(lib().orc_update_halo(h.ptr(), _fptr(parent)))
    return parent


def scatter(h: DescHandle, src_parent: np.ndarray, dst_parent: np.ndarray) -> np.ndarray:
    _ck(lib().orc_scatter(h.ptr(), _fptr(src_parent), _fptr(dst_parent)))
    return dst_parent


def iterate(h: DescHandle, a: np.ndarray, b: np.ndarray, nsteps: int) -> np.ndarray:
    _ck(lib().orc_iterate(h.ptr(), _fptr(a), _fptr(b), nsteps))
    return a if nsteps % 2 == 0 else b


# ---- convenience: the reference's StencilArray bookkeeping, for tests that read like test/array.jl ----
def stencil_array_sweep(r: np.ndarray, offs, radius: int, boundary: int, padding: str, reducer: int, *,
                        padval=0, weights=None, alpha=0.0, born_mask=1 << 3, survive_mask=0b1100,
                        switching: bool = False) -> np.ndarray:
    """mapstencil(f, StencilArray(r, stencil; boundary, padding)) through the oracle.

    padding: "cond" (Conditional), "out" (Halo{:out}), "in" (Halo{:in}). With `switching` the dest is the
    padded twin buffer (src/gatherstencil.jl:77-83) and the logical window of it is returned.
    """
    r = np.asfortranarray(r)
    nd = r.ndim
    et = A.ELTYPE_OF_DTYPE[r.dtype]
    R = radius
    if padding == "cond":
        parent, size, off = r.copy(order="F"), r.shape, (0,) * nd
    elif padding == "out":  # pad_array(::Halo{:out}), src/padding.jl:104-110 (ring content undefined: poison it)
        parent = np.full(tuple(s + 2 * R for s in r.shape), 77, dtype=r.dtype, order="F")
        parent[tuple(slice(R, R + s) for s in r.shape)] = r
        size, off = r.shape, (R,) * nd
    elif padding == "in":  # pad_array(::Halo{:in}) = parent itself, size shrinks by 2R (src/array.jl:470-473)
        parent, size, off = r.copy(order="F"), tuple(s - 2 * R for s in r.shape), (R,) * nd
    else:
        raise ValueError(padding)
    oet = out_eltype(reducer, et)
    dst_off = off if switching else (0,) * nd
    h = build_desc(size=size, eltype=et, out_eltype=oet, offsets=offs, radius=R, boundary=boundary,
                   reducer=reducer, src_off=off, dst_off=dst_off, src_ext=parent.shape, padval=padval,
                   weights=weights, alpha=alpha, born_mask=born_mask, survive_mask=survive_mask)
    if padding != "cond" and boundary != A.USE:
        update_halo(h, parent)  # gatherstencil! calls update_boundary!(source) first, src/gatherstencil.jl:93
    out = gather(h, parent)
    if switching:
        out = out[tuple(slice(o, o + s) for o, s in zip(dst_off, size))]
    return np.asfortranarray(out)
