"""Stencil algebra: the host-side mirror of src/stencil.jl and src/stencils/*.jl.

A stencil is an ordered table of offsets around a centre cell. Named shapes get their table from the C ABI
(`sb200_stencil_offsets`, which restates the reference's generators: box (-R:R)^N with the first axis
fastest, filtered by the shape predicate); Positional / NamedStencil / Rectangle carry user offsets.
Index tuples follow Julia: `(o1, o2[, o3])`, o1 along the contiguous (first) axis, 1-based array indices
in `indices`.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _abi as A


class Stencil:
    """abstract type Stencil{R,N,L,T} (src/stencil.jl:17). `neighbors`/`center` are filled by `stencil(A, I)`."""
    shape_enum: int | None = None

    def __init__(self, offsets, radius, ndims, neighbors=None, center=None):
        self._offsets = tuple(tuple(int(v) for v in o) for o in offsets)
        self.radius = int(radius)
        self.ndims = int(ndims)
        self.neighbors = neighbors
        self.center = center

    # -- reference API (src/stencil.jl:38-129) --
    def offsets(self):
        return self._offsets

    def __len__(self):
        return len(self._offsets)

    def __iter__(self):
        return iter(self.neighbors)

    def __getitem__(self, i):
        return self.neighbors[i]

    def rebuild(self, neighbors, center):
        import copy
        s = copy.copy(self)
        s.neighbors, s.center = tuple(neighbors), center
        return s

    def __eq__(self, other):
        return isinstance(other, Stencil) and self._offsets == other._offsets and self.radius == other.radius

    def __hash__(self):
        return hash((self._offsets, self.radius))

    def __repr__(self):
        return f"{type(self).__name__}{{R={self.radius},N={self.ndims},L={len(self)}}}"


def _named(shape_enum, radius, ndims, inner=0):
    n = C.c_int32()
    l = A.lib()
    A.check(l.sb200_stencil_offsets(shape_enum, radius, inner, ndims, None, 0, C.byref(n)))
    buf = np.zeros((max(n.value, 1), 3), dtype=np.int32)
    A.check(l.sb200_stencil_offsets(shape_enum, radius, inner, ndims, buf.ctypes.data, n.value, C.byref(n)))
    return [tuple(int(v) for v in row[:ndims]) for row in buf[:n.value]]


def _make_shape(name, enum, doc):
    def __init__(self, radius=1, ndims=2):
        Stencil.__init__(self, _named(enum, radius, ndims), radius, ndims)
    return type(name, (Stencil,), {"__init__": __init__, "__doc__": doc, "shape_enum": enum})


Window = _make_shape("Window", A.WINDOW, "Radius-R box including the centre (src/stencils/window.jl).")
Moore = _make_shape("Moore", A.MOORE, "Radius-R box without the centre (src/stencils/moore.jl).")
VonNeumann = _make_shape("VonNeumann", A.VONNEUMANN, "Manhattan distance 1..R, no centre (src/stencils/vonneumman.jl).")
Cross = _make_shape("Cross", A.CROSS, "Offsets with zeros on at least N-1 axes (src/stencils/shapes.jl:2-10).")
AngledCross = _make_shape("AngledCross", A.ANGLEDCROSS, "All diagonals (src/stencils/shapes.jl:13-25).")
ForwardSlash = _make_shape("ForwardSlash", A.FORWARDSLASH, "Forward diagonal (src/stencils/shapes.jl:28-40).")
BackSlash = _make_shape("BackSlash", A.BACKSLASH, "Backward diagonal (src/stencils/shapes.jl:43-55).")
Circle = _make_shape("Circle", A.CIRCLE, "Cells whose centre is within R+0.5 (src/stencils/shapes.jl:58-67).")
Vertical = _make_shape("Vertical", A.VERTICAL, "Vertical bar or plane (src/stencils/shapes.jl:70-79).")
Horizontal = _make_shape("Horizontal", A.HORIZONTAL, "Horizontal bar or plane (src/stencils/shapes.jl:82-91).")
Diamond = _make_shape("Diamond", A.DIAMOND, "Manhattan distance 0..R (src/stencils/shapes.jl:94-104).")
Cardinal = _make_shape("Cardinal", A.CARDINAL, "N,S,W,E at distance R (src/stencils/shapes.jl:154-163).")
Ordinal = _make_shape("Ordinal", A.ORDINAL, "NE,SE,SW,NW at distance R (src/stencils/shapes.jl:166-175).")


class Annulus(Stencil):
    """Annulus{RO,RI,N}: RI+0.5 <= dist < RO+0.5 (src/stencils/shapes.jl:116-151)."""
    shape_enum = A.ANNULUS

    def __init__(self, outer_radius=2, inner_radius=None, ndims=2):
        inner_radius = outer_radius - 1 if inner_radius is None else inner_radius
        super().__init__(_named(A.ANNULUS, outer_radius, ndims, inner_radius), outer_radius, ndims)
        self.inner_radius = inner_radius


class Positional(Stencil):
    """Positional(offsets...): arbitrary offsets in user order (src/stencils/positional.jl:31-66).
    The radius is the largest |offset| (the reference takes the largest signed offset)."""

    def __init__(self, *offsets):
        if len(offsets) == 1 and offsets[0] and isinstance(offsets[0][0], (tuple, list)):
            offsets = tuple(offsets[0])
        offsets = [tuple(o) if isinstance(o, (tuple, list)) else (o,) for o in offsets]
        n = len(offsets[0])
        if any(len(o) != n for o in offsets):
            raise A.ArgumentError(f"All offsets must be the length `N` of {n}, got {offsets}")
        super().__init__(offsets, max(abs(v) for o in offsets for v in o), n)


class NamedStencil(Positional):
    """NamedStencil(; name=offset...) / NamedStencil(names, stencil) (src/stencils/named.jl:43-92)."""

    def __init__(self, *args, **named):
        if args and isinstance(args[0], Cardinal) and len(args) == 1:
            args = (("E", "S", "N", "W"), args[0])      # src/stencils/named.jl:91
        elif args and isinstance(args[0], Ordinal) and len(args) == 1:
            args = (("SE", "NE", "SW", "NW"), args[0])  # src/stencils/named.jl:92
        if args:
            names, st = args
            offs = st.offsets() if isinstance(st, Stencil) else tuple(st)
            if len(names) != len(offs):
                raise A.ArgumentError("Length of keys must match length of offsets and L parameter")
        else:
            names, offs = tuple(named), tuple(named.values())
        super().__init__(*offs)
        self.names = tuple(names)

    def __getattr__(self, name):
        names = self.__dict__.get("names", ())
        if name in names and self.neighbors is not None:
            return self.neighbors[names.index(name)]
        raise AttributeError(name)


class Rectangle(Stencil):
    """Rectangle((lo1,hi1),(lo2,hi2)...): per-axis ranges, first axis fastest (src/stencils/rectangle.jl:13-48)."""

    def __init__(self, *ranges):
        if len(ranges) == 1 and isinstance(ranges[0][0], (tuple, list)):
            ranges = tuple(ranges[0])
        if any(len(r) != 2 for r in ranges):
            raise A.ArgumentError(f"All offset tuples must have length `2`, got {ranges}")
        out = []
        n = len(ranges)

        # CartesianIndices(map(splat(:), O)): the first axis varies fastest

        def rec(axis, cur):
            if axis < 0:
                out.append(tuple(cur))
                return
            lo, hi = ranges[axis]
            for v in range(lo, hi + 1):
                cur[axis] = v
                rec(axis - 1, cur)
        rec(n - 1, [0] * n)
        super().__init__(out, max(abs(v) for r in ranges for v in r), n)


class Kernel(Stencil):
    """Kernel(stencil, weights) / Kernel(weights) / Kernel(f, stencil) (src/stencils/kernel.jl:93-112).
    `weights` are indexed linearly in column-major order, matching the offset order of Window."""

    def __init__(self, *args):
        if len(args) == 1:
            w = np.asarray(args[0])
            st = Window(w.shape[0] // 2, w.ndim)  # src/stencils/kernel.jl:111
        elif callable(args[0]) and not isinstance(args[0], Stencil):
            st = args[1]
            w = np.array([args[0](d) for d in distances(st)])  # src/stencils/kernel.jl:98-106
        else:
            st, w = args[0], np.asarray(args[1])
        w = np.asarray(w).reshape(-1, order="F")
        if len(st) != w.size:
            raise A.ArgumentError(f"Stencil length {len(st)} does not match kernel length {w.size}")
        super().__init__(st.offsets(), st.radius, st.ndims)
        self.stencil = st
        self.kernel = w
        self.shape_enum = st.shape_enum


class Layered:
    """Layered(layers...) / Layered(name=stencil, ...) / Layered(tuple_or_dict) (src/stencils/layered.jl:13-57): stencils that
    are used together on one array. `radius` is the largest layer radius (it sizes the Halo ring), `len` and the accessor
    functions map over the layers; layers may be Layered themselves. On the GPU a Layered array feeds multi-table gathers:
    `LinearCombination(layer(0, sum), (-1.0, layer(1, sum)))` is the reference test's `sum(l[1]) - sum(l[2])`
    (test/stencils.jl:242-264), one sweep of the same parent per referenced layer."""

    def __init__(self, *layers, **named):
        if named:
            if layers:
                raise A.ArgumentError("Layered takes positional layers or keyword layers, not both")
            names, layers = tuple(named), tuple(named.values())
        else:
            names = None
            if len(layers) == 1 and isinstance(layers[0], dict):
                names, layers = tuple(layers[0]), tuple(layers[0].values())
            elif len(layers) == 1 and isinstance(layers[0], (tuple, list)):
                layers = tuple(layers[0])
        if not layers or not all(isinstance(l, (Stencil, Layered)) for l in layers):
            raise A.ArgumentError("Layered needs one or more stencils")
        self.layers = tuple(layers)
        self.names = names
        self.ndims = layers[0].ndims                       # ndimensions(first(layers))
        self.radius = max(l.radius for l in layers)        # maximum(map(radius, layers))

    def __len__(self):
        raise TypeError("length of a Layered is a tuple: use lengths()")

    def lengths(self):
        """Base.length(l::Layered) = map(length, layers(l))"""
        return tuple(l.lengths() if isinstance(l, Layered) else len(l) for l in self.layers)

    def _index(self, key):
        if isinstance(key, str):
            if self.names is None or key not in self.names:
                raise KeyError(key)
            return self.names.index(key)
        return int(key)

    def __getitem__(self, key):
        if isinstance(key, tuple):   # a path through nested layers
            cur = self
            for k in key:
                cur = cur[k]
            return cur
        return self.layers[self._index(key)]

    def __getattr__(self, name):
        names = self.__dict__.get("names")
        if names and name in names:
            return self.__dict__["layers"][names.index(name)]
        raise AttributeError(name)

    def __iter__(self):
        return iter(self.layers)

    def offsets(self):
        return tuple(l.offsets() for l in self.layers)

    @property
    def neighbors(self):
        return tuple(l.neighbors for l in self.layers)

    @property
    def center(self):
        return tuple(l.center for l in self.layers)

    def rebuild(self, layerneighbors, centers):
        out = Layered(*[l.rebuild(n, c) for l, n, c in zip(self.layers, layerneighbors, centers)])
        out.names = self.names
        return out

    def __eq__(self, other):
        return isinstance(other, Layered) and self.layers == other.layers and self.names == other.names

    def __hash__(self):
        return hash((self.layers, self.names))

    def __repr__(self):
        inner = ", ".join((f"{n}=" if self.names else "") + repr(l) for n, l in zip(self.names or [None] * len(self.layers), self.layers))
        return f"Layered{{R={self.radius},N={self.ndims}}}({inner})"


class layer:
    """A term of a LinearCombination over a Layered array: g applied to layer `key` (index, name, or a tuple path through
    nested layers) — `sum(l[1])` is layer(0, sum), `sum(l.l1.b)` is layer(("l1", "b"), sum)."""

    def __init__(self, key, g):
        self.key, self.g = key, g


# ---- free functions of the reference API ----
def offsets(s):
    return s.offsets()


def radius(s):
    return s.radius


def diameter(s):
    return 2 * (s if isinstance(s, int) else s.radius) + 1


def neighbors(s, *I):
    if isinstance(s, (Stencil, Layered)):
        return s.neighbors
    return s.neighbors(*I)  # StencilArray


def center(s):
    return s.center


def distances(s):
    if isinstance(s, Layered):
        return tuple(distances(l) for l in s.layers)
    return tuple(math.sqrt(sum(v * v for v in o)) for o in s.offsets())  # src/stencil.jl:116-120


def distance_zones(s):
    if isinstance(s, Layered):
        return tuple(distance_zones(l) for l in s.layers)
    return tuple(sum(abs(v) for v in o) for o in s.offsets())  # src/stencil.jl:127-129


def indices(s, I):
    """indices(hood, I) (src/stencil.jl:90-96) or indices(A::StencilArray, I) with the array's boundary."""
    if isinstance(s, Layered):
        return tuple(indices(l, I) for l in s.layers)   # indices(layered, I) = map(l -> indices(l, I), layered)
    if not isinstance(s, Stencil):
        return s.indices(I)
    I = tuple(I)
    n = s.ndims
    return tuple(tuple(o[a] + I[a] for a in range(n)) + I[n:] for o in s.offsets())


def merge(a, *rest):
    """merge(stencils...) -> Positional with sorted unique offsets (src/stencil.jl:158-167)."""
    offs = set(a.offsets())
    for b in rest:
        if b.ndims != a.ndims:
            raise A.ArgumentError(f"Stencils must have the same dimensionality to merge. Got {a.ndims} and {b.ndims}")
        offs |= set(b.offsets())
    if not rest:
        return a
    return Positional(*sorted(offs))
