"""Helpers for the GPU parity tests: run the SAME sb200_desc through the CPU oracle and through the C ABI."""
import numpy as np

from stencils_b200 import _abi as A


def to_dev(a: np.ndarray):
    """Column-major NumPy parent -> CUDA tensor with the same memory layout (first axis contiguous)."""
    import torch
    a = np.asfortranarray(a)
    t = torch.from_numpy(np.ascontiguousarray(a.T)).cuda()  # C-order of the transpose == F-order bytes
    return t


def to_host(t, shape, dtype) -> np.ndarray:
    out = t.cpu().numpy().astype(dtype, copy=False)
    return np.asfortranarray(out.T).reshape(shape, order="F")


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def sync():
    import torch
    torch.cuda.synchronize()


def dst_like(h, fill=None):
    d = h.desc
    shape = tuple(d.dst_ext[a] for a in range(d.ndim))
    dt = A.DTYPE_OF_ELTYPE[d.out_eltype]
    if fill is None:
        return np.zeros(shape, dtype=dt, order="F")
    return np.full(shape, fill, dtype=dt, order="F")


def gpu_gather(h, src_parent: np.ndarray, dst_parent: np.ndarray | None = None, halo=False):
    """[sb200_update_halo] + sb200_gather with device buffers; returns (dst_parent, src_parent_after)."""
    l = A.lib()
    dst_parent = dst_like(h) if dst_parent is None else dst_parent
    ts, td = to_dev(src_parent), to_dev(dst_parent)
    if halo:
        A.check(l.sb200_update_halo(h.ptr(), ts.data_ptr(), stream()))
    A.check(l.sb200_gather(h.ptr(), ts.data_ptr(), td.data_ptr(), stream()))
    sync()
    return to_host(td, dst_parent.shape, dst_parent.dtype), to_host(ts, src_parent.shape, src_parent.dtype)


def gpu_scatter(h, src_parent, dst_parent):
    l = A.lib()
    ts, td = to_dev(src_parent), to_dev(dst_parent)
    A.check(l.sb200_scatter(h.ptr(), ts.data_ptr(), td.data_ptr(), stream()))
    sync()
    return to_host(td, dst_parent.shape, dst_parent.dtype)


def bits_equal(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape, (a.dtype, b.dtype, a.shape, b.shape)
    if a.dtype.kind == "f":
        u = f"u{a.itemsize}"
        same = a.view(u) == b.view(u)
        both_nan = np.isnan(a) & np.isnan(b)  # NaN payloads are not part of the contract
        bad = ~(same | both_nan)
    else:
        bad = a != b
    if bad.any():
        idx = np.argwhere(bad)[:5]
        raise AssertionError(f"{bad.sum()} of {a.size} cells differ; first at {idx.tolist()}: "
                             f"{[a[tuple(i)] for i in idx]} vs {[b[tuple(i)] for i in idx]}")


def max_ulp(a, b):
    a, b = np.asarray(a), np.asarray(b)
    it = {4: np.int32, 8: np.int64}[a.itemsize]
    ai, bi = a.view(it).astype(np.int64), b.view(it).astype(np.int64)
    ai = np.where(ai < 0, np.iinfo(it).min - ai, ai)
    bi = np.where(bi < 0, np.iinfo(it).min - bi, bi)
    return int(np.abs(ai - bi).max())
