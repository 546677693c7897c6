"""Deterministic synthetic fields for benchmarks and parity tests: the value at linear
(column-major) index i is hash-to-uniform(splitmix64(seed ^ i)). The NumPy and the torch (device) generators
produce bit-identical values, so tiles of a full-size device field can be regenerated on the host."""
from __future__ import annotations

import numpy as np

_C1, _C2, _C3 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB
LIFE_DENSITY = 0.35


def _uniform_np(lo: int, n: int, seed: int) -> np.ndarray:
    i = np.arange(lo, lo + n, dtype=np.uint64) ^ np.uint64(seed)
    with np.errstate(over="ignore"):
        z = i + np.uint64(_C1)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_C2)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_C3)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def _convert_np(u, dtype):
    dtype = np.dtype(dtype)
    if dtype.kind in "ub":
        return (u < LIFE_DENSITY).astype(dtype)
    return u.astype(dtype)


def synth_np(shape, dtype, seed: int, lo: int = 0) -> np.ndarray:
    """Column-major NumPy array of `shape`; `lo` offsets the linear index (for slabs of a larger field)."""
    n = int(np.prod(shape))
    out = np.empty(n, dtype=dtype)
    chunk = 1 << 22
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        out[s:s + m] = _convert_np(_uniform_np(lo + s, m, seed), dtype)
    return out.reshape(shape, order="F")


def _s64(c: int) -> int:
    return c - (1 << 64) if c >= (1 << 63) else c


def synth_torch(shape, dtype, seed: int, device, lo: int = 0):
    """Same field generated on the device; returns a column-major torch tensor of logical `shape`."""
    import torch
    from .array import _torch_dtype, colmajor_empty
    n = int(np.prod(shape))
    out = colmajor_empty(tuple(shape), _torch_dtype(dtype), device)
    flat = out.permute(*reversed(range(len(shape)))).reshape(-1)  # memory order == column-major linear index

    def lsr(x, k):
        return (x >> k) & ((1 << (64 - k)) - 1)

    chunk = 1 << 26
    kind = np.dtype(dtype).kind
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        z = torch.arange(lo + s, lo + s + m, dtype=torch.int64, device=device) ^ _s64(seed)
        z = z + _s64(_C1)
        z = (z ^ lsr(z, 30)) * _s64(_C2)
        z = (z ^ lsr(z, 27)) * _s64(_C3)
        z = z ^ lsr(z, 31)
        u = lsr(z, 11).to(torch.float64) * (1.0 / (1 << 53))
        flat[s:s + m] = (u < LIFE_DENSITY).to(flat.dtype) if kind in "ub" else u.to(flat.dtype)
    return out
