"""Builds `sb200_desc` (include/stencils_b200.h) from host-side StencilArray bookkeeping.

The size/offset rules restate src/padding.jl:54-110 and src/array.jl:367-385,470-473:
Conditional keeps the parent size; Halo parents are `size + 2R` on EVERY array axis (even when the
stencil has fewer dimensions than the array) and logical index 0 sits at parent index R.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _abi as A


@dataclass
class DescHandle:
    desc: A.Desc
    keep: list = field(default_factory=list)  # numpy buffers the descriptor points into

    def ptr(self):
        return C.byref(self.desc)

    def copy(self, **updates) -> "DescHandle":
        d = A.Desc()
        C.memmove(C.byref(d), C.byref(self.desc), C.sizeof(A.Desc))
        h = DescHandle(d, list(self.keep))
        for k, v in updates.items():
            if isinstance(v, (list, tuple)):
                arr = getattr(d, k)
                for i, x in enumerate(v):
                    arr[i] = x
            else:
                setattr(d, k, v)
        return h


def padval_bits(padval, eltype: int) -> int:
    dt = A.DTYPE_OF_ELTYPE[eltype]
    raw = np.array([padval]).astype(dt).tobytes()
    return int.from_bytes(raw.ljust(8, b"\0"), "little")


def require_exact(values, dt, what: str) -> None:
    """The reference promotes: `kernelproduct` accumulates hood[i] * kernel[i] (src/stencils/kernel.jl:37-43) and the result
    type follows typeof(padval) (_arg_return_type), so Float64 weights or a Float64 padval on an Int32 / Float32 grid give a
    Float64 result there. The kernels here compute in the source element type, so a value that this type cannot hold exactly
    is refused by the user-facing layer (ops.py) instead of being cast silently (0.1 on an Int32 grid used to become 0).
    build_desc itself stays a plain cast: tests and the slab plans hand it values of the right type."""
    v = np.asarray(values)
    with np.errstate(all="ignore"):
        c = v.astype(dt)
        back = c.astype(v.dtype) if v.dtype.kind in "fiub" else c
        same = (back == v) | ((v != v) & (back != back)) if v.dtype.kind == "f" else (back == v)
    if dt == np.dtype(np.bool_) and v.dtype.kind != "b":
        same = same & ((v == 0) | (v == 1))
    if not np.all(same):
        raise A.ArgumentError(f"{what} {v.reshape(-1)[:4].tolist()} cannot be represented exactly in the source element type "
                              f"{np.dtype(dt).name}; the reference would promote the result type, these kernels compute in the source type "
                              "(convert the array, or use values of its element type)")


def build_desc(*, size, eltype, out_eltype, offsets, radius, boundary, reducer=A.SUM,
               src_off=None, dst_off=None, src_ext=None, dst_ext=None, padval=0, weights=None,
               born_mask=1 << 3, survive_mask=(1 << 2) | (1 << 3), alpha=0.0,
               scatter_op=A.OP_ADD, scatter_rule=A.SCATTER_WEIGHTS, flags=0, region=None) -> DescHandle:
    size = tuple(int(s) for s in size)
    nd = len(size)
    if not 1 <= nd <= 3:
        raise A.ArgumentError(f"arrays must have 1..3 dimensions, got {nd}")
    offs = np.zeros((len(offsets), 3), dtype=np.int32)
    for k, o in enumerate(offsets):
        o = tuple(o) if not np.isscalar(o) else (o,)
        if len(o) > nd:
            raise A.ArgumentError(f"stencil has {len(o)} dimensions but the array has {nd}")
        offs[k, :len(o)] = o
    offs = np.ascontiguousarray(offs)
    bcs = boundary if isinstance(boundary, (list, tuple)) else (boundary,) * nd
    src_off = tuple(src_off) if src_off is not None else (0,) * nd
    dst_off = tuple(dst_off) if dst_off is not None else (0,) * nd
    src_ext = tuple(src_ext) if src_ext is not None else tuple(s + 2 * o for s, o in zip(size, src_off))
    dst_ext = tuple(dst_ext) if dst_ext is not None else tuple(s + 2 * o for s, o in zip(size, dst_off))
    d = A.Desc()
    d.struct_size = C.sizeof(A.Desc)
    d.ndim = nd
    for a in range(3):
        d.size[a] = size[a] if a < nd else 1
        d.src_ext[a] = src_ext[a] if a < nd else 1
        d.dst_ext[a] = dst_ext[a] if a < nd else 1
        d.src_off[a] = src_off[a] if a < nd else 0
        d.dst_off[a] = dst_off[a] if a < nd else 0
        d.boundary[a] = bcs[a] if a < nd else A.REMOVE
    d.eltype = eltype
    d.out_eltype = out_eltype
    d.padval_bits = padval_bits(padval, eltype)
    d.radius = int(radius)
    d.noffsets = len(offsets)
    d.offsets_host = offs.ctypes.data
    d.reducer = reducer
    d.scatter_op = scatter_op
    d.scatter_rule = scatter_rule
    d.born_mask = born_mask
    d.survive_mask = survive_mask
    d.alpha = float(alpha)
    keep = [offs]
    if weights is not None:
        w = np.ascontiguousarray(np.asarray(weights).reshape(-1, order="F").astype(A.DTYPE_OF_ELTYPE[eltype]))
        if w.size != len(offsets):
            raise A.ArgumentError(f"Stencil length {len(offsets)} does not match kernel length {w.size}")
        d.weights_host = w.ctypes.data
        keep.append(w)
    if region is not None:
        lo, hi = region
        for a in range(nd):
            d.region_lo[a] = lo[a]
            d.region_hi[a] = hi[a]
    d.flags = flags
    return DescHandle(d, keep)
