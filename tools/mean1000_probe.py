#!/usr/bin/env python
"""One-GPU probe for BASELINE configs[0] at the README's size (mean, Window(1), Float64 1000 x 1000, Remove(0)): where do the
8.2 us per sweep go? (a) sb.mapstencil_ in a Python loop (what bench.py times), (b) bare ctypes sb200_gather calls, (c) the same
launches captured in a CUDA graph and replayed (no host cost per launch).   tools/mean1000_probe.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import stencils_b200 as sb  # noqa: E402
from stencils_b200 import _abi as A  # noqa: E402
from stencils_b200._desc import build_desc  # noqa: E402
from stencils_b200.synth import synth_torch  # noqa: E402

dev = torch.device("cuda", 0)
shape = (1000, 1000)
a = synth_torch(shape, np.float64, 0x5EED0001, dev)
dst = torch.empty_like(a)
lib = A.lib()
sa = sb.StencilArray(a, sb.Window(1))
K = 400


def timed(fn, reps=15):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) / K * 1e3, min(ts) / K * 1e3


def py_loop():
    for _ in range(K):
        sb.mapstencil_(sb.mean, dst, sa)


print("sb.mapstencil_ loop      : %.2f us per sweep (best %.2f), kernel %s" % (*timed(py_loop), lib.sb200_last_kernel().decode()), flush=True)
for flags, name in ((0, "default"), (A.FLAG_FORCE_GENERIC, "FORCE_GENERIC"), (A.FLAG_NO_TMA, "NO_TMA")):
    d = build_desc(size=shape, eltype=A.F64, out_eltype=A.F64, offsets=sb.Window(1).offsets(), radius=1, boundary=A.REMOVE, reducer=A.MEAN, padval=0.0, flags=flags)
    st = torch.cuda.current_stream().cuda_stream

    def c_loop():
        for _ in range(K):
            lib.sb200_gather(d.ptr(), a.data_ptr(), dst.data_ptr(), st)
    print("ctypes sb200_gather loop, %-14s: %.2f us per sweep (best %.2f), kernel %s" % (name, *timed(c_loop), lib.sb200_last_kernel().decode()), flush=True)
    s2 = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s2):
        lib.sb200_gather(d.ptr(), a.data_ptr(), dst.data_ptr(), s2.cuda_stream)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s2):
            for _ in range(K):
                lib.sb200_gather(d.ptr(), a.data_ptr(), dst.data_ptr(), s2.cuda_stream)
    print("CUDA graph of %d launches, %-14s: %.2f us per sweep (best %.2f)" % (K, name, *timed(g.replay)), flush=True)
