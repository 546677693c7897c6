"""StencilArray / SwitchingStencilArray, boundary conditions and padding — host-side mirror of
src/array.jl:388-631, src/boundary.jl and src/padding.jl. These objects only keep bookkeeping (parent buffer,
stencil, boundary, padding); the sweeps are lowered to the C ABI in ops.py.

Storage. Arrays are column-major like Julia's: the first axis is contiguous. Host parents are F-ordered
NumPy arrays; device parents are torch CUDA tensors viewed with column-major strides (`colmajor_empty`).
"""
from __future__ import annotations

import numpy as np

from . import _abi as A
from .stencils import Stencil, Window


# ---- boundary conditions (src/boundary.jl) ----
class BoundaryCondition:
    enum = None


class Wrap(BoundaryCondition):
    enum = A.WRAP


class Reflect(BoundaryCondition):
    enum = A.REFLECT


class Use(BoundaryCondition):
    enum = A.USE


class Remove(BoundaryCondition):
    enum = A.REMOVE
    _unset = object()

    def __init__(self, padval=_unset):
        self.padval = None if padval is Remove._unset else padval


def padval(x):
    return (x.boundary if hasattr(x, "boundary") else x).padval


# ---- padding (src/padding.jl:7-39) ----
class Padding:
    pass


class Conditional(Padding):
    kind = "cond"


class Halo(Padding):
    def __init__(self, kind="out"):
        kind = str(kind).lstrip(":")
        if kind not in ("in", "out"):
            raise A.ArgumentError(f"Halo must be :in or :out, got {kind}")
        self.kind = kind


# ---- storage helpers ----
def _is_torch(x):
    return type(x).__module__.startswith("torch")


def colmajor_empty(shape, dtype, device):
    """torch tensor of logical `shape` whose first axis is contiguous (Julia layout)."""
    import torch
    t = torch.empty(tuple(reversed(shape)), dtype=dtype, device=device)
    return t.permute(*reversed(range(len(shape))))


def as_colmajor(x):
    """Return x itself when it already is column-major, else a column-major copy."""
    if _is_torch(x):
        nd = x.dim()
        if x.permute(*reversed(range(nd))).is_contiguous():
            return x
        y = colmajor_empty(tuple(x.shape), x.dtype, x.device)
        y.copy_(x)
        return y
    x = np.asarray(x)
    return x if x.flags.f_contiguous else np.asfortranarray(x)


def _np_dtype(x):
    if _is_torch(x):
        import torch
        return np.dtype({torch.bool: np.bool_, torch.uint8: np.uint8, torch.int32: np.int32, torch.int64: np.int64,
                         torch.float32: np.float32, torch.float64: np.float64}[x.dtype])
    return np.dtype(x.dtype)


def _torch_dtype(npdt):
    import torch
    return {np.dtype(np.bool_): torch.bool, np.dtype(np.uint8): torch.uint8, np.dtype(np.int32): torch.int32,
            np.dtype(np.int64): torch.int64, np.dtype(np.float32): torch.float32,
            np.dtype(np.float64): torch.float64}[np.dtype(npdt)]


def similar(x, dtype=None, shape=None):
    """similar(parent, T, dims): same kind of storage (host/device), column-major."""
    shape = tuple(x.shape) if shape is None else tuple(shape)
    npdt = _np_dtype(x) if dtype is None else np.dtype(dtype)
    if _is_torch(x):
        return colmajor_empty(shape, _torch_dtype(npdt), x.device)
    return np.empty(shape, dtype=npdt, order="F")


def data_ptr(x) -> int:
    return x.data_ptr() if _is_torch(x) else x.ctypes.data


def is_device(x) -> bool:
    return _is_torch(x) and x.is_cuda


def _inner_slices(shape, R):
    return tuple(slice(R, s - R) for s in shape)


def pad_array(padding, stencil, parent):
    """pad_array (src/padding.jl:102-110): only Halo{:out} allocates (ring content is undefined)."""
    if isinstance(padding, Halo) and padding.kind == "out":
        R = stencil.radius
        big = similar(parent, shape=tuple(s + 2 * R for s in parent.shape))
        big[_inner_slices(big.shape, R)] = parent
        return big
    return parent


class AbstractStencilArray:
    """abstract type AbstractStencilArray (src/array.jl:7)."""
    stencil: Stencil
    boundary: BoundaryCondition
    padding: Padding

    # -- size bookkeeping: src/array.jl:379, 470-473 --
    @property
    def radius(self):
        return self.stencil.radius

    @property
    def halo(self) -> int:
        return self.stencil.radius if isinstance(self.padding, Halo) else 0

    @property
    def shape(self):
        h = self.halo
        return tuple(s - 2 * h for s in self.parent.shape)

    size = shape

    @property
    def ndim(self):
        return len(self.parent.shape)

    @property
    def dtype(self):
        return _np_dtype(self.parent)

    def inner(self, buf=None):
        """View of the logical cells of a parent buffer (add_halo, src/array.jl:367-377)."""
        buf = self.parent if buf is None else buf
        return buf[_inner_slices(buf.shape, self.halo)] if self.halo else buf

    def __array__(self, dtype=None, copy=None):
        v = self.inner()
        a = v.cpu().numpy() if _is_torch(v) else np.asarray(v)
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, I):
        return self.inner()[I]

    def __setitem__(self, I, v):
        self.inner()[I] = v

    def __eq__(self, other):
        return bool(np.array_equal(np.asarray(self), np.asarray(other)))

    __hash__ = None

    # -- indices / stencil fill on the host (cold path; src/array.jl:22-30, 140-179) --
    def bounded_index(self, I):
        if isinstance(self.padding, Halo) or isinstance(self.boundary, Remove):
            return tuple(I)
        out = []
        for i, s in zip(I, self.shape):
            if isinstance(self.boundary, Wrap):
                out.append(i + s if i < 1 else (i - s if i > s else i))
            else:
                out.append(2 - i if i < 1 else (2 * s - i if i > s else i))
        return tuple(out)

    def indices(self, I, _st=None):
        from .stencils import Layered, indices as st_indices
        st = self.stencil if _st is None else _st
        if isinstance(st, Layered):
            return tuple(self.indices(I, l) for l in st.layers)
        return tuple(self.bounded_index(J) for J in st_indices(st, tuple(I)))

    def neighbors(self, *I, _st=None):
        I = tuple(I[0]) if len(I) == 1 and isinstance(I[0], (tuple, list)) else tuple(I)
        from .stencils import Layered, indices as st_indices
        st = self.stencil if _st is None else _st
        if isinstance(st, Layered):   # neighbors(::Layered, A, I): a tuple per layer (src/array.jl:117-120)
            return tuple(self.neighbors(I, _st=l) for l in st.layers)
        h, par, vals = self.halo, self.parent, []
        for J in st_indices(st, I):
            if h:
                vals.append(par[tuple(j - 1 + h for j in J)])
                continue
            inb = all(1 <= j <= s for j, s in zip(J, self.shape))
            if isinstance(self.boundary, Remove):
                vals.append(par[tuple(j - 1 for j in J)] if inb else self.boundary.padval)
            else:
                vals.append(par[tuple(j - 1 for j in self.bounded_index(J))])
        return tuple(v.item() if hasattr(v, "item") else v for v in vals)

    def stencil_at(self, *I):
        I = tuple(I[0]) if len(I) == 1 and isinstance(I[0], (tuple, list)) else tuple(I)
        c = self.inner()[tuple(i - 1 for i in I)]
        c = c.item() if hasattr(c, "item") else c
        from .stencils import Layered

        def centers(st):   # _center(::Layered, A, I) = map over the layers (src/array.jl:33)
            return tuple(centers(l) for l in st.layers) if isinstance(st, Layered) else c
        return self.stencil.rebuild(self.neighbors(I), centers(self.stencil))


def _check_radius(parent, stencil):
    # src/array.jl:451-453
    if len(parent.shape) < stencil.ndims:
        raise A.ArgumentError(f"stencil has {stencil.ndims} dimensions but the array has {len(parent.shape)}")
    for s in parent.shape:
        if not stencil.radius < s:
            raise A.ArgumentError(f"stencil radius is larger than array axis {s}")


class StencilArray(AbstractStencilArray):
    """StencilArray(A, stencil; boundary=Remove(zero(eltype(A))), padding=Conditional()) (src/array.jl:441-468)."""

    def __init__(self, parent, stencil=None, boundary=None, padding=None, *, _padded=False):
        parent = as_colmajor(parent)
        stencil = Window(1, len(parent.shape)) if stencil is None else stencil
        self.boundary = Remove(np.zeros((), dtype=_np_dtype(parent))[()]) if boundary is None else boundary
        self.padding = Conditional() if padding is None else padding
        self.stencil = stencil
        self.parent = parent if _padded else pad_array(self.padding, stencil, parent)
        _check_radius(self.parent, stencil)

    def similar(self, dtype=None):
        return StencilArray(similar(self.parent, dtype), self.stencil, self.boundary, self.padding, _padded=True)

    def copy(self):
        c = self.similar()
        c.parent[...] = self.parent
        return c


class SwitchingStencilArray(AbstractStencilArray):
    """SwitchingStencilArray(A, stencil; boundary, padding) (src/array.jl:564-611): two same-size buffers.
    For Conditional / Halo{:in} `source` IS the user's array and `dest` a copy (src/array.jl:593-598)."""

    def __init__(self, parent, stencil=None, boundary=None, padding=None, *, _dest=None):
        parent = as_colmajor(parent)
        stencil = Window(1, len(parent.shape)) if stencil is None else stencil
        self.boundary = Remove(np.zeros((), dtype=_np_dtype(parent))[()]) if boundary is None else boundary
        self.padding = Conditional() if padding is None else padding
        self.stencil = stencil
        if _dest is not None:
            self.source, self.dest = parent, _dest
        else:
            self.source = pad_array(self.padding, stencil, parent)
            self.dest = pad_array(self.padding, stencil, parent)
            if self.dest is self.source:
                self.dest = similar(self.source)
                self.dest[...] = self.source
        if tuple(self.source.shape) != tuple(self.dest.shape):
            raise A.ArgumentError(f"source and dest arrays must be the same size, got {tuple(self.source.shape)} "
                                  f"and {tuple(self.dest.shape)}")
        _check_radius(self.source, stencil)

    @property
    def parent(self):
        return self.source

    def switch(self):
        return SwitchingStencilArray(self.dest, self.stencil, self.boundary, self.padding, _dest=self.source)

    def copy(self):
        s, d = similar(self.source), similar(self.dest)
        s[...] = self.source
        d[...] = self.dest
        return SwitchingStencilArray(s, self.stencil, self.boundary, self.padding, _dest=d)


def switch(A_):
    return A_.switch()


def source(A_):
    return A_.source if isinstance(A_, SwitchingStencilArray) else A_.parent


def dest(A_):
    return A_.dest if isinstance(A_, SwitchingStencilArray) else A_.parent


def boundary(A_):
    return A_.boundary


def padding(A_):
    return A_.padding


def stencil(A_, *I):
    """stencil(A) -> the (empty) stencil; stencil(A, I...) -> filled stencil around 1-based index I."""
    return A_.stencil if not I else A_.stencil_at(*I)
