"""mapstencil / gatherstencil / scatterstencil / update_boundary — host-side mirror of
src/gatherstencil.jl and src/scatterstencil.jl. Everything here is bookkeeping; each sweep is ONE call
into libstencils_b200.so (include/stencils_b200.h). A user function outside the reducer menu raises
ArgumentError: there is no KernelAbstractions/PyTorch/CPU fallback.

Python has no `!` in names: the mutating Julia functions `f!` are spelled `f_` (torch convention):
    mapstencil!(f, dest, source) -> mapstencil_(f, dest, source)
"""
from __future__ import annotations

import builtins
import ctypes as C
import operator
import statistics

import numpy as np

from . import _abi as A
from ._desc import DescHandle, build_desc, require_exact
from .array import (AbstractStencilArray, Halo, Remove, StencilArray, SwitchingStencilArray, Use, _is_torch, _np_dtype,
                    data_ptr, is_device, similar)
from .stencils import Kernel, Layered, Stencil, layer


# ---- the reducer menu (BASELINE.json north_star: mean, sum, min, max, Kernel dot-product, Life table; + diffusion) ----
class Reducer:
    enum = None
    params: dict = {}


class _Named(Reducer):
    def __init__(self, enum, name):
        self.enum, self.__name__ = enum, name

    def __repr__(self):
        return self.__name__


sum = _Named(A.SUM, "sum")            # noqa: A001  (mirrors Julia's Base.sum on a Stencil)
mean = _Named(A.MEAN, "mean")
minimum = _Named(A.MIN, "minimum")
maximum = _Named(A.MAX, "maximum")
kernelproduct = _Named(A.KERNELDOT, "kernelproduct")  # src/stencils/kernel.jl:34-43


class Life(Reducer):
    """Life-like rule table over the neighbour count: born/survive are iterables of counts (default B3/S23)."""
    enum = A.LIFE

    def __init__(self, born=(3,), survive=(2, 3)):
        self.born_mask = builtins.sum(1 << int(b) for b in set(born))
        self.survive_mask = builtins.sum(1 << int(s) for s in set(survive))


class Diffusion(Reducer):
    """centre + alpha * (sum(hood) - L * centre), each operation rounded separately."""
    enum = A.DIFFUSION

    def __init__(self, alpha):
        self.alpha = float(alpha)


_ALIASES = {}
for _f, _r in ((builtins.sum, sum), (np.sum, sum), (np.mean, mean), (statistics.mean, mean), (statistics.fmean, mean),
               (builtins.max, maximum), (np.max, maximum), (np.amax, maximum), (builtins.min, minimum),
               (np.min, minimum), (np.amin, minimum)):
    _ALIASES[_f] = _r


def resolve_reducer(f) -> Reducer:
    if isinstance(f, Reducer):
        return f
    try:
        if f in _ALIASES:
            return _ALIASES[f]
    except TypeError:
        pass
    raise A.ArgumentError(
        f"unsupported user function {f!r}: mapstencil lowers only sum, mean, minimum, maximum, kernelproduct, "
        "Life(...) and Diffusion(alpha) to CUDA kernels and has no fallback path")


def out_eltype(reducer: int, eltype: int) -> int:
    o = C.c_int32()
    A.check(A.lib().sb200_out_eltype(reducer, eltype, C.byref(o)))  # _return_type, src/gatherstencil.jl:41-59
    return o.value


def _bc_enum(bc):
    if isinstance(bc, Remove) and bc.padval is None:
        raise A.ArgumentError("Remove() without a padval (padval = nothing) is not a numeric boundary")
    return bc.enum


_raw_stream = None


def _stream():
    """cudaStream_t of torch's current stream on the current device (raw accessor when this torch has it: the Stream object
    round trip costs ~2 us per call, which matters for launch-bound grids like the README's 1000 x 1000)."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", False)
    if _raw_stream:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _desc_for(red: Reducer, src_parent, src_halo: int, dst_parent, dst_halo: int, st: Stencil, bc, *,
              flags=0, region=None, scatter=None) -> DescHandle:
    """Descriptor of one sweep. Small grids are launch-bound (README benchmark: 1000 x 1000), so the descriptor of a
    repeated call is built once and kept on the stencil object (keyed by everything else that enters it)."""
    if scatter is None and region is None:
        key = (type(red), getattr(red, "enum", None), getattr(red, "born_mask", None), getattr(red, "survive_mask", None),
               getattr(red, "alpha", None), tuple(src_parent.shape), src_halo, tuple(dst_parent.shape), dst_halo,
               str(_np_dtype(src_parent)), str(_np_dtype(dst_parent)), type(bc), getattr(bc, "padval", None), flags)
        cache = st.__dict__.setdefault("_desc_cache", {})
        h = cache.get(key)
        if h is None:
            h = cache[key] = _desc_build(red, src_parent, src_halo, dst_parent, dst_halo, st, bc, flags=flags)
            if len(cache) > 64:
                cache.pop(next(iter(cache)))
        return h
    return _desc_build(red, src_parent, src_halo, dst_parent, dst_halo, st, bc, flags=flags, region=region, scatter=scatter)


def _desc_build(red: Reducer, src_parent, src_halo: int, dst_parent, dst_halo: int, st: Stencil, bc, *,
                flags=0, region=None, scatter=None) -> DescHandle:
    et = A.ELTYPE_OF_DTYPE.get(_np_dtype(src_parent))
    if et is None:
        raise A.ArgumentError(f"unsupported element type {_np_dtype(src_parent)}")
    nd = len(src_parent.shape)
    size = tuple(s - 2 * src_halo for s in src_parent.shape)
    kw = {}
    if scatter is not None:
        oet = et
        kw.update(scatter)
        enum = A.SUM
    else:
        enum = red.enum
        oet = out_eltype(enum, et)
    if _np_dtype(dst_parent) != A.DTYPE_OF_ELTYPE[oet]:
        raise A.ArgumentError(f"dest eltype {_np_dtype(dst_parent)} does not match the result type "
                              f"{A.DTYPE_OF_ELTYPE[oet]} of {red!r}")
    dsize = tuple(s - 2 * dst_halo for s in dst_parent.shape)
    if dsize != size:  # _checksizes, src/gatherstencil.jl:118-124
        raise A.ArgumentError(f"Source array sizes must match. Found: {dsize} and {size}")
    weights = kw.pop("weights", None)
    if enum == A.KERNELDOT:
        if not isinstance(st, Kernel):
            raise A.ArgumentError("kernelproduct needs a Kernel stencil")
        weights = st.kernel
    if isinstance(red, Life):
        kw.update(born_mask=red.born_mask, survive_mask=red.survive_mask)
    if isinstance(red, Diffusion):
        kw.update(alpha=red.alpha)
    pv = bc.padval if isinstance(bc, Remove) else 0
    # the reference promotes on mixed types (kernel.jl:37-43, _arg_return_type); here values must fit the source element type
    require_exact(np.array([pv]), A.DTYPE_OF_ELTYPE[et], "Remove padval")
    if weights is not None and et not in (A.BOOL, A.U8):
        require_exact(np.asarray(weights), A.DTYPE_OF_ELTYPE[et], "kernel / scatter weights")
    return build_desc(size=size, eltype=et, out_eltype=oet, offsets=st.offsets(), radius=st.radius,
                      boundary=_bc_enum(bc), reducer=enum, src_off=(src_halo,) * nd, dst_off=(dst_halo,) * nd,
                      src_ext=tuple(src_parent.shape), dst_ext=tuple(dst_parent.shape), padval=pv, weights=weights,
                      flags=flags, region=region, **kw)


def _same_place(a, b):
    if is_device(a) != is_device(b):
        raise A.ArgumentError("source and dest must both be host arrays or both be device tensors")


def update_boundary_(A_: AbstractStencilArray, buf=None):
    """update_boundary!(A) (src/array.jl:195-239): refresh the halo ring of the source parent."""
    if not isinstance(A_.padding, Halo) or isinstance(A_.boundary, Use):
        return A_
    par = A_.parent if buf is None else buf
    # the halo refresh has no reducer and no dest: `minimum` keeps the element type for every eltype (sum(Bool) is Int64, which
    # tripped the dest-eltype check for Bool parents: ADVICE r1)
    h = _desc_for(minimum, par, A_.halo, par, A_.halo, A_.stencil, A_.boundary)
    if is_device(par):
        A.check(A.lib().sb200_update_halo(h.ptr(), data_ptr(par), _stream()))
    else:  # host parent: one round trip of the parent through the halo kernel
        import torch
        t = torch.from_numpy(np.ascontiguousarray(par.T)).cuda()
        A.check(A.lib().sb200_update_halo(h.ptr(), t.data_ptr(), _stream()))
        par[...] = t.cpu().numpy().T
    return A_


def _gather_into(red: Reducer, dst_parent, dst_halo, src_parent, src_halo, st, bc, flags=0):
    # Repeated call with the very same objects (the launch-bound case: README benchmark, 1000 x 1000): everything that was
    # checked and built the first time is remembered on the stencil object, keyed by object identity.
    fast = st.__dict__.get("_fast_call")
    if (fast is not None and fast[0] is red and fast[1]() is src_parent and fast[2]() is dst_parent and fast[3] is bc and fast[4] == flags
            and fast[5] == (src_halo, dst_halo, src_parent.shape, dst_parent.shape)):
        rc = fast[7](fast[6], src_parent.data_ptr(), dst_parent.data_ptr(), _stream())
        if rc:
            A.check(rc)
        return dst_parent
    _same_place(src_parent, dst_parent)
    h = _desc_for(red, src_parent, src_halo, dst_parent, dst_halo, st, bc, flags=flags)
    l = A.lib()
    if is_device(src_parent) and not (src_halo and not isinstance(bc, Use)):
        import weakref   # weak: a long-lived stencil object must not keep multi-GiB parents alive
        st.__dict__["_fast_call"] = (red, weakref.ref(src_parent), weakref.ref(dst_parent), bc, flags,
                                     (src_halo, dst_halo, src_parent.shape, dst_parent.shape), h.ptr(), l.sb200_gather, h)
    if is_device(src_parent):
        if src_halo and not isinstance(bc, Use):
            A.check(l.sb200_update_halo(h.ptr(), data_ptr(src_parent), _stream()))   # src/gatherstencil.jl:93
        A.check(l.sb200_gather(h.ptr(), data_ptr(src_parent), data_ptr(dst_parent), _stream()))
    else:
        if _is_torch(src_parent):
            raise A.ArgumentError("CPU torch tensors are not supported; pass a NumPy array or a CUDA tensor")
        A.check(l.sb200_gather_host(h.ptr(), data_ptr(src_parent), data_ptr(dst_parent)))  # refreshes the ring too
    return dst_parent


class LinearCombination(Reducer):
    """f(hood_1, hood_2, ...) = c_1*g_1(hood_1) + c_2*g_2(hood_2) + ... evaluated left to right, every operation
    rounded separately — the multi-array user functions of test/array.jl:312-383, e.g.
    `center(a) + 0.1 * sum(neighbors(b))` is LinearCombination(center, (0.1, sum)).
    One term per array argument, in order: `g` or `(coef, g)` with g in {center, sum, mean, minimum, maximum,
    kernelproduct}. A plain (non-stencil) array argument is indexed, not stencilled (src/gatherstencil.jl:112-113):
    its term must be `center`."""

    def __init__(self, *terms):
        from .stencils import center as _center
        self.terms = []
        self.layer_keys = []
        for t in terms:
            coef, g = (None, t) if not isinstance(t, (tuple, list)) else (float(t[0]), t[1])
            key = None
            if isinstance(g, layer):   # a term over one layer of a Layered array (src/stencils/layered.jl)
                key, g = g.key, g.g
            self.layer_keys.append(key)
            if g is _center or g == "center":
                g = "center"
            else:
                g = resolve_reducer(g)
                if g.enum not in (A.SUM, A.MEAN, A.MIN, A.MAX, A.KERNELDOT):
                    raise A.ArgumentError(f"{g!r} cannot be a term of a LinearCombination")
            self.terms.append((coef, g))

    def __repr__(self):
        return "LinearCombination(" + ", ".join(f"{c}*{g}" if c is not None else f"{g}" for c, g in self.terms) + ")"


def gather_multi_(f: LinearCombination, dst, *srcs):
    """gatherstencil!(f, dest, A1, A2, ...) with several array arguments -> sb200_gather_multi."""
    keys = getattr(f, "layer_keys", [None] * len(f.terms))
    if len(srcs) == 1 and isinstance(srcs[0], AbstractStencilArray) and isinstance(srcs[0].stencil, Layered):
        # one Layered array: every term reads the same parent through the table of its layer
        if any(k is None for k in keys):
            raise A.ArgumentError("every term over a Layered array must name its layer: layer(key, g)")
        srcs = srcs * len(f.terms)
    elif any(k is not None for k in keys):
        raise A.ArgumentError("layer(key, g) terms need a single StencilArray with a Layered stencil")
    if len(srcs) != len(f.terms):
        raise A.ArgumentError(f"{f!r} has {len(f.terms)} terms but {len(srcs)} array arguments were passed")
    if isinstance(dst, AbstractStencilArray):
        dst_parent, dst_halo = dst.parent, dst.halo
    else:
        dst_parent, dst_halo = dst, 0
    terms = (A.Term * len(srcs))()
    keep = []
    for j, (src, (coef, g)) in enumerate(zip(srcs, f.terms)):
        if isinstance(src, AbstractStencilArray):
            par, halo, st, bc = src.parent, src.halo, src.stencil, src.boundary
            if keys[j] is not None:
                try:
                    st = st[keys[j]]
                except (KeyError, IndexError, TypeError):
                    raise A.ArgumentError(f"the Layered stencil has no layer {keys[j]!r}") from None
            if isinstance(st, Layered):
                raise A.ArgumentError("a term needs one stencil: pick a leaf layer with layer(key, g) (a tuple key walks nested layers)")
        else:
            if g != "center":
                raise A.ArgumentError("a plain array argument is indexed, not stencilled: its term must be `center`")
            from .stencils import Positional
            par, halo, st, bc = src, 0, Positional(*[(0,) * len(src.shape)]), Remove(0)
        if not is_device(par) or not is_device(dst_parent):
            raise A.ArgumentError("multi-array gathers need device arrays")
        if g == "center":
            from .stencils import Positional
            one = Positional(*[(0,) * st.ndims])
            one.radius = st.radius          # the parent's ring thickness is the array's stencil radius
            h = _desc_build(sum, par, halo, dst_parent, dst_halo, one, bc)
        else:
            h = _desc_build(g, par, halo, dst_parent, dst_halo, st, bc)
        keep.append(h)
        terms[j].desc = C.pointer(h.desc)
        terms[j].src_parent = data_ptr(par)
        terms[j].has_coef = 0 if coef is None else 1
        terms[j].coef = 0.0 if coef is None else coef
    A.check(A.lib().sb200_gather_multi(C.cast(terms, C.c_void_p), len(srcs), data_ptr(dst_parent), None, _stream()))
    return dst


def gatherstencil_(f, *args, flags=0):
    """gatherstencil!(f, dest, source) / gatherstencil!(f, A::SwitchingStencilArray) (src/gatherstencil.jl:77-103).
    Returns dest, or the switched array for a SwitchingStencilArray (the caller must rebind it)."""
    if isinstance(f, LinearCombination):
        if isinstance(args[0], SwitchingStencilArray) and len(args) == len(f.terms):
            S = args[0]   # gatherstencil!(f, A::SwitchingStencilArray, args...): source(A) is the first argument
            first = StencilArray(S.source, S.stencil, S.boundary, S.padding, _padded=True)
            gather_multi_(f, StencilArray(S.dest, S.stencil, S.boundary, S.padding, _padded=True), first, *args[1:])
            return S.switch()
        return gather_multi_(f, args[0], *args[1:])
    red = resolve_reducer(f)
    if len(args) == 1 and isinstance(args[0], SwitchingStencilArray):
        S = args[0]
        _gather_into(red, S.dest, S.halo, S.source, S.halo, S.stencil, S.boundary, flags)
        return S.switch()
    if len(args) != 2 or not isinstance(args[1], AbstractStencilArray):
        raise A.ArgumentError("gatherstencil!(f, dest, source): extra array arguments are not supported by the "
                              "CUDA reducer menu (SURVEY §8f.1)")
    dst, src = args
    if isinstance(dst, AbstractStencilArray):
        _gather_into(red, dst.parent, dst.halo, src.parent, src.halo, src.stencil, src.boundary, flags)
    else:
        fast = src.stencil.__dict__.get("_fast_call")
        if fast is None or fast[2]() is not dst:   # a dest that went through _gather_into before was checked then
            from .array import as_colmajor
            if as_colmajor(dst) is not dst:
                raise A.ArgumentError("dest must be column-major (first axis contiguous)")
        _gather_into(red, dst, 0, src.parent, src.halo, src.stencil, src.boundary, flags)
    return dst


def gatherstencil(f, *args, boundary=None, padding=None, flags=0):
    """gatherstencil(f, A::StencilArray) / gatherstencil(f, stencil, A; boundary, padding) (src/gatherstencil.jl:14-39)."""
    if isinstance(f, LinearCombination):
        first = args[0]
        if not isinstance(first, AbstractStencilArray):
            raise A.ArgumentError("the first argument of a multi-array gather must be a StencilArray")
        et = A.ELTYPE_OF_DTYPE.get(first.dtype)
        dst = similar(first.parent, A.DTYPE_OF_ELTYPE[et], first.shape)
        return gather_multi_(f, dst, *args)
    red = resolve_reducer(f)
    if isinstance(args[0], Stencil):
        if len(args) != 2:
            raise A.ArgumentError("extra array arguments are not supported by the CUDA reducer menu (SURVEY §8f.1)")
        src = StencilArray(args[1], args[0], boundary, padding)
    else:
        if len(args) != 1 or not isinstance(args[0], AbstractStencilArray):
            raise A.ArgumentError("extra array arguments are not supported by the CUDA reducer menu (SURVEY §8f.1)")
        src = args[0]
    et = A.ELTYPE_OF_DTYPE.get(src.dtype)
    if et is None:
        raise A.ArgumentError(f"unsupported element type {src.dtype}")
    oet = out_eltype(red.enum, et)
    dst = similar(src.parent, A.DTYPE_OF_ELTYPE[oet], src.shape)  # similar(parent(source), T_return, size(source))
    _gather_into(red, dst, 0, src.parent, src.halo, src.stencil, src.boundary, flags)
    return dst


mapstencil = gatherstencil    # deprecated aliases in the reference, src/gatherstencil.jl:127-128
mapstencil_ = gatherstencil_


def iterate_(f, S: SwitchingStencilArray, nsteps: int) -> SwitchingStencilArray:
    """`for _ in 1:nsteps; A = mapstencil!(f, A); end` as one stream-ordered C-ABI call (sb200_iterate):
    no host synchronisation between steps. Returns the array whose `source` holds the final state."""
    red = resolve_reducer(f)
    h = _desc_for(red, S.source, S.halo, S.dest, S.halo, S.stencil, S.boundary)
    l = A.lib()
    if is_device(S.source):
        A.check(l.sb200_iterate(h.ptr(), data_ptr(S.source), data_ptr(S.dest), int(nsteps), _stream()))
        return S if nsteps % 2 == 0 else S.switch()
    A.check(l.sb200_iterate_host(h.ptr(), data_ptr(S.source), int(nsteps)))
    return S


# ---- scatterstencil! (src/scatterstencil.jl) ----
class ScatterRule:
    """Fixed menu for the user function of scatterstencil!: the value sent to offset k."""


class ScatterWeights(ScatterRule):
    """val_k = w_k (e.g. `map(_ -> 0.1, neighbors(hood))`, test/array.jl:391-393)."""
    enum = A.SCATTER_WEIGHTS

    def __init__(self, weights):
        self.weights = weights


class ScatterCenterWeights(ScatterRule):
    """val_k = center(hood) * w_k (w = 1: `map(_ -> center(hood), neighbors(hood))`, test/array.jl:412-415)."""
    enum = A.SCATTER_CENTER_WEIGHTS

    def __init__(self, weights=1.0):
        self.weights = weights


_OPS = {operator.add: A.OP_ADD, "+": A.OP_ADD, builtins.sum: A.OP_ADD, builtins.max: A.OP_MAX, max: A.OP_MAX,
        "max": A.OP_MAX, builtins.min: A.OP_MIN, "min": A.OP_MIN, np.add: A.OP_ADD, np.maximum: A.OP_MAX,
        np.minimum: A.OP_MIN}


def scatterstencil_(f, op, dest_or_switching, src=None, flags=0):
    """scatterstencil!(f, op, dest, source) and the SwitchingStencilArray forms (src/scatterstencil.jl:36-45,115-133)."""
    if not isinstance(f, ScatterRule):
        raise A.ArgumentError(f"unsupported scatter function {f!r}: use ScatterWeights(w) or ScatterCenterWeights(w)")
    if op not in _OPS:
        raise A.ArgumentError(f"unsupported scatter op {op!r}: use +, max or min")
    switching = isinstance(dest_or_switching, SwitchingStencilArray)
    if switching:
        S = dest_or_switching
        dst = S.dest
        source_arr = src if src is not None else StencilArray(S.source, S.stencil, S.boundary,
                                                              Halo("in") if S.halo else S.padding, _padded=True)
        flags |= A.FLAG_ZERO_DEST
        dst_halo = 0
        if S.halo:  # src/scatterstencil.jl:125-133 trips _checksizes here
            raise A.ArgumentError("Source array sizes must match: Switching + Halo scatter passes the padded dest")
    else:
        dst, source_arr, dst_halo = dest_or_switching, src, 0
    st = source_arr.stencil
    require_exact(np.asarray(f.weights), np.dtype(source_arr.dtype), "scatter weights")
    w = np.broadcast_to(np.asarray(f.weights, dtype=source_arr.dtype), (len(st),)).copy()
    _same_place(source_arr.parent, dst)
    h = _desc_for(sum, source_arr.parent, source_arr.halo, dst, dst_halo, st, source_arr.boundary, flags=flags,
                  scatter=dict(weights=w, scatter_op=_OPS[op], scatter_rule=f.enum))
    l = A.lib()
    refresh = source_arr.halo and not isinstance(source_arr.boundary, Use)   # update_boundary!(source), src/scatterstencil.jl:39
    if is_device(dst):
        if refresh:
            A.check(l.sb200_update_halo(h.ptr(), data_ptr(source_arr.parent), _stream()))
        A.check(l.sb200_scatter(h.ptr(), data_ptr(source_arr.parent), data_ptr(dst), _stream()))
    else:
        import torch
        ts = torch.from_numpy(np.ascontiguousarray(source_arr.parent.T)).cuda()
        td = torch.from_numpy(np.ascontiguousarray(dst.T)).cuda()
        if refresh:
            A.check(l.sb200_update_halo(h.ptr(), ts.data_ptr(), _stream()))
        A.check(l.sb200_scatter(h.ptr(), ts.data_ptr(), td.data_ptr(), _stream()))
        dst[...] = td.cpu().numpy().T
        if refresh:
            source_arr.parent[...] = ts.cpu().numpy().T   # the reference refreshes the caller's ring in place
    return dest_or_switching.switch() if switching else dst
