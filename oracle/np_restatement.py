"""Second, independent restatement of the reference semantics in NumPy — TEST INFRASTRUCTURE ONLY.

Formulated differently from oracle/stencils_oracle.c on purpose (pad-then-shift instead of per-neighbour
index arithmetic; per-destination sorted fold instead of the pass loops for scatter) so that the two
restatements check each other. Slow; tiny shapes only.
"""
from __future__ import annotations

import itertools
import math

import numpy as np

# ---- offsets: src/stencils/*.jl — CartesianIndices((-R:R)^N) with the first axis fastest ----
_PRED = {
    "Window": lambda t, R, N, RI: True,
    "Moore": lambda t, R, N, RI: any(t),
    "VonNeumann": lambda t, R, N, RI: 1 <= sum(map(abs, t)) <= R,
    "Cross": lambda t, R, N, RI: sum(x == 0 for x in t) >= N - 1,
    "AngledCross": lambda t, R, N, RI: sum(abs(x) == abs(t[0]) for x in t[1:]) == N - 1,
    "ForwardSlash": lambda t, R, N, RI: sum(x == -t[0] for x in t[1:]) == N - 1,
    "BackSlash": lambda t, R, N, RI: sum(x == t[0] for x in t[1:]) == N - 1,
    "Circle": lambda t, R, N, RI: math.sqrt(sum(x * x for x in t)) < R + 0.5,
    "Vertical": lambda t, R, N, RI: (N > 1 and t[1] == 0) or (N == 1 and t[0] == 0),
    "Horizontal": lambda t, R, N, RI: N > 1 and t[0] == 0,
    "Diamond": lambda t, R, N, RI: sum(map(abs, t)) <= R,
    "Annulus": lambda t, R, N, RI: RI + 0.5 <= math.sqrt(sum(x * x for x in t)) < R + 0.5,
    "Cardinal": lambda t, R, N, RI: sum(map(abs, t)) == R and max(map(abs, t)) == R,
    "Ordinal": lambda t, R, N, RI: sum(map(abs, t)) == R * N and max(map(abs, t)) == R,
}
SHAPE_NAMES = list(_PRED)  # index == sb200_shape enum value


def offsets(name: str, R: int, N: int = 2, RI: int = 0):
    out = []
    for rev in itertools.product(range(-R, R + 1), repeat=N):  # last varies fastest ...
        t = rev[::-1]                                          # ... so reverse: first axis fastest
        if _PRED[name](t, R, N, RI):
            out.append(t)
    return out


def _jl_max(a, b):
    if a.dtype.kind != "f":
        return np.maximum(a, b)
    r = np.where(a > b, a, np.where(a < b, b, np.where(np.signbit(a), b, a)))
    return np.where(np.isnan(a) | np.isnan(b), np.nan, r).astype(a.dtype)


def _jl_min(a, b):
    if a.dtype.kind != "f":
        return np.minimum(a, b)
    r = np.where(a < b, a, np.where(a > b, b, np.where(np.signbit(a), a, b)))
    return np.where(np.isnan(a) | np.isnan(b), np.nan, r).astype(a.dtype)


def padded(inner: np.ndarray, R: int, boundary: str, padval=0) -> np.ndarray:
    """What a Halo{:out} parent holds after update_boundary! (src/array.jl:202-239)."""
    if boundary == "remove":
        return np.pad(inner, R, mode="constant", constant_values=np.array(padval).astype(inner.dtype))
    if boundary == "wrap":
        return np.pad(inner, R, mode="wrap")
    if boundary == "reflect":
        return np.pad(inner, R, mode="reflect")  # mirror without repeating the edge == 2-i / 2s-i
    raise ValueError(boundary)


def gather(r: np.ndarray, offs, R: int, boundary: str, padding: str, reducer: str, *, padval=0,
           weights=None, alpha=0.0, born_mask=1 << 3, survive_mask=0b1100) -> np.ndarray:
    """mapstencil(f, StencilArray(r, stencil; boundary, padding)) -> plain array of size(A)."""
    r = np.asarray(r)
    nd = r.ndim
    if padding == "in":
        inner = r[tuple(slice(R, s - R) for s in r.shape)]
        P = r if boundary == "use" else padded(inner, R, boundary, padval)
    else:
        inner = r
        P = padded(inner, R, boundary, padval)
    size = inner.shape
    L = len(offs)

    def shifted(o):
        o = tuple(o) + (0,) * (nd - len(o))
        return P[tuple(slice(R + oa, R + oa + s) for oa, s in zip(o, size))]

    T = r.dtype
    with np.errstate(over="ignore", invalid="ignore"):
        if reducer in ("sum", "mean", "diffusion"):
            acc = shifted(offs[0]).astype(np.int64) if T == np.bool_ else shifted(offs[0]).copy()
            for o in offs[1:]:
                acc = (acc + shifted(o)).astype(acc.dtype)
            if reducer == "sum":
                return acc
            if reducer == "mean":
                if T.kind == "f":
                    return (acc / T.type(L)).astype(T)
                return acc.astype(np.float64) / np.float64(L)
            c = inner
            lc = (T.type(L) * c).astype(T)
            u = (acc - lc).astype(T)
            v = (T.type(alpha) * u).astype(T)
            return (c + v).astype(T)
        if reducer in ("max", "min"):
            f = _jl_max if reducer == "max" else _jl_min
            acc = shifted(offs[0]).copy()
            for o in offs[1:]:
                acc = f(acc, shifted(o))
            return acc.astype(T)
        if reducer == "kerneldot":
            w = np.asarray(weights).reshape(-1, order="F").astype(T)
            acc = np.zeros(size, dtype=T)
            for k, o in enumerate(offs):
                p = (shifted(o) * w[k]).astype(T)
                acc = (acc + p).astype(T)
            return acc
        if reducer == "life":
            cnt = np.zeros(size, dtype=np.int64)
            for o in offs:
                cnt += shifted(o) != 0
            alive = inner != 0
            bit = np.where(alive, (survive_mask >> cnt) & 1, (born_mask >> cnt) & 1)
            return bit.astype(T)
    raise ValueError(reducer)


def scatter(src: np.ndarray, dest0: np.ndarray, offs, R: int, boundary: str, op: str, rule: str, weights):
    """scatterstencil!(f, op, dest, source) as a per-destination fold: for every dest cell collect the
    contributions (pass = mod1(j_src, 2R+1), j_src, i_src, k), sort, and fold into dest's prior value.
    This is the "gather-transpose" order SURVEY §3.3 derives from src/scatterstencil.jl:56-71."""
    ny, nx = src.shape
    T = src.dtype
    S = 2 * R + 1
    w = np.asarray(weights).astype(T)
    contrib = {}
    for j in range(1, nx + 1):
        for i in range(1, ny + 1):
            c = src[i - 1, j - 1]
            for k, (o1, o2) in enumerate(offs):
                t = [i + o1, j + o2]
                n = [ny, nx]
                skip = False
                for a in range(2):
                    if boundary == "wrap":
                        t[a] = (t[a] - 1) % n[a] + 1
                    elif boundary == "reflect":
                        t[a] = 2 - t[a] if t[a] < 1 else (2 * n[a] - t[a] if t[a] > n[a] else t[a])
                    elif t[a] < 1 or t[a] > n[a]:
                        skip = True
                if skip:
                    continue
                with np.errstate(over="ignore"):
                    val = w[k] if rule == "weights" else T.type(c * w[k])
                contrib.setdefault((t[0], t[1]), []).append((((j - 1) % S) + 1, j, i, k, val))
    dest = dest0.copy()
    for (ti, tj), lst in contrib.items():
        acc = dest[ti - 1, tj - 1]
        for _, _, _, _, val in sorted(lst, key=lambda e: e[:4]):
            with np.errstate(over="ignore"):
                if op == "add":
                    acc = T.type(acc + val)
                elif op == "max":
                    acc = _jl_max(np.array(acc), np.array(val))[()]
                else:
                    acc = _jl_min(np.array(acc), np.array(val))[()]
        dest[ti - 1, tj - 1] = acc
    return dest
