#!/usr/bin/env python
"""One-GPU probe: a two-argument gather  mean(Window(1) of A) + 0.5 * sum(VonNeumann(1) of B)  on 16384^2 Float32 through
sb200_gather_multi — the opt-in single-pass kernel (csrc/multi_tile.cu, SB200_MULTI_SINGLE_PASS=1) against the default sweep-per-argument path.
Algorithmic traffic: two reads + one write = 12 bytes per cell.   tools/multi_probe.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import stencils_b200 as sb  # noqa: E402
from stencils_b200 import _abi as A  # noqa: E402
from stencils_b200.synth import synth_torch  # noqa: E402

dev = torch.device("cuda", 0)
shape = (16384, 16384)
cells = shape[0] * shape[1]
a = synth_torch(shape, np.float32, 0x5EED0001, dev)
b = synth_torch(shape, np.float32, 0x5EED0002, dev)
sa = sb.StencilArray(a, sb.Window(1), boundary=sb.Remove(np.float32(0)))
sbb = sb.StencilArray(b, sb.VonNeumann(1), boundary=sb.Wrap())
f = sb.LinearCombination(sb.mean, (0.5, sb.sum))
lib = A.lib()
peak = 6547.0
for mode in ("1", "0"):
    os.environ["SB200_MULTI_SINGLE_PASS"] = mode
    out = sb.mapstencil(f, sa, sbb)
    torch.cuda.synchronize()
    ts = []
    for _ in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = sb.mapstencil(f, sa, sbb)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    med = float(np.median(ts))
    print(f"SB200_MULTI_SINGLE_PASS={mode}: {med:.3f} ms per call (min {min(ts):.3f}), {cells / med / 1e6:.1f} Gcell/s, "
          f"{cells * 12 / med / 1e6 / peak:.3f} of the HBM roofline at 12 B per cell, last kernel {lib.sb200_last_kernel().decode()}", flush=True)
    ref = out if mode == "1" else ref
    if mode == "0":
        print("bit-identical:", bool(torch.equal(out.view(torch.int32), ref.view(torch.int32))))
