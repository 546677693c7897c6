#!/usr/bin/env python
"""One-GPU probe: duration of ONE launch of the Life kernels by generations per launch (SB200_FLAG_GENS(n), n = 1 .. 8) on the
BASELINE grid (16384^2 UInt8, Wrap), for 0/1 cells (the steady state of sb200_iterate) and for raw UInt8 cells (its first launch).
The relative times are the kLifeCost table of split_steps in csrc/api.cu. Also times sb200_iterate for a few step counts with the
default split and with SB200_POW2_STEPS=1 (sizes 1 / 2 / 4 / 8 only).   tools/life_gens_probe.py"""
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
shape = (16384, 16384)
cells = shape[0] * shape[1]
lib = A.lib()
a = synth_torch(shape, np.uint8, 0x5EED0005, dev)
b = torch.empty_like(a)
moore = sb.Moore(1).offsets()
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=12):
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
    return float(np.median(ts)), float(min(ts))


for cells01 in (1, 0):
    for n in range(1, 9):
        d = build_desc(size=shape, eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1, boundary=A.WRAP, reducer=A.LIFE,
                       flags=A.flag_gens(n) | (A.FLAG_CELLS_01 if cells01 else 0))

        def one():
            for _ in range(4):   # four back-to-back launches per sample (a -> b, b -> a, ...): the state stays 0/1
                A.check(lib.sb200_gather(d.ptr(), a.data_ptr(), b.data_ptr(), st))
                A.check(lib.sb200_gather(d.ptr(), b.data_ptr(), a.data_ptr(), st))
        med, mn = timed(one)
        us = med / 8 * 1e3
        print(f"gens {n} cells01={cells01}: {us:8.1f} us per launch (min {mn / 8 * 1e3:.1f}), {cells * n / us / 1e6:9.1f} Gcell-updates/s, "
              f"kernel {lib.sb200_last_kernel().decode()}", flush=True)

# packed -> packed launches (SB200_FLAG_SRC_BITS | SB200_FLAG_DST_BITS): the two buffers hold the packed grids (garbage bits are as good as any)
for n in range(1, 9):
    d = build_desc(size=shape, eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1, boundary=A.WRAP, reducer=A.LIFE,
                   flags=A.flag_gens(n) | A.FLAG_SRC_BITS | A.FLAG_DST_BITS)

    def one_pk():
        for _ in range(4):
            A.check(lib.sb200_gather(d.ptr(), a.data_ptr(), b.data_ptr(), st))
            A.check(lib.sb200_gather(d.ptr(), b.data_ptr(), a.data_ptr(), st))
    med, mn = timed(one_pk)
    us = med / 8 * 1e3
    print(f"gens {n} packed -> packed: {us:8.1f} us per launch (min {mn / 8 * 1e3:.1f}), {cells * n / us / 1e6:9.1f} Gcell-updates/s, "
          f"kernel {lib.sb200_last_kernel().decode()}", flush=True)
a.copy_(synth_torch(shape, np.uint8, 0x5EED0005, dev))

h1 = build_desc(size=shape, eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1, boundary=A.WRAP, reducer=A.LIFE)
for gm in ("5", "6", "7", "8"):
    os.environ["SB200_LIFE_PACKED_GENS"] = gm
    for n in (20, 100, 1000):
        def run_pk():
            A.check(lib.sb200_iterate(h1.ptr(), a.data_ptr(), b.data_ptr(), n, st))
        med, mn = timed(run_pk, reps=10 if n < 1000 else 5)
        print(f"sb200_iterate {n:5d} steps, SB200_LIFE_PACKED_GENS={gm}: {cells * n / med / 1e6:9.1f} Gcell-updates/s (best {cells * n / mn / 1e6:.1f})", flush=True)
del os.environ["SB200_LIFE_PACKED_GENS"]
for pow2 in ("0", "1"):
    os.environ["SB200_POW2_STEPS"] = pow2
    for n in (10, 20, 21, 30, 50, 100, 1000):
        def run():
            A.check(lib.sb200_iterate(h1.ptr(), a.data_ptr(), b.data_ptr(), n, st))
        med, mn = timed(run, reps=10 if n < 1000 else 5)
        print(f"sb200_iterate {n:5d} steps, SB200_POW2_STEPS={pow2}: {cells * n / med / 1e6:9.1f} Gcell-updates/s (best {cells * n / mn / 1e6:.1f})", flush=True)
