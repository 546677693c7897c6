#!/usr/bin/env python
"""One-GPU probe: sb200_iterate Life runs of 1000 generations on small grids — byte launches replayed as CUDA graphs (default up to
4 Mi cells) against packed runs (SB200_LIFE_PACKED=1), on the default stream and on a real stream.   tools/life_small_probe.py"""
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
lib = A.lib()
moore = sb.Moore(1).offsets()
N = 1000
for shape in ((1024, 1024), (2048, 2048), (4096, 4096), (8192, 8192)):
    a = synth_torch(shape, np.uint8, 0x5EED0005, dev)
    b = torch.empty_like(a)
    cells = shape[0] * shape[1]
    h = build_desc(size=shape, eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1, boundary=A.WRAP, reducer=A.LIFE)
    s2 = torch.cuda.Stream()
    for packed in ("0", "1"):
        os.environ["SB200_LIFE_PACKED"] = packed
        for st, name in ((None, "default stream"), (s2.cuda_stream, "real stream")):
            def run():
                A.check(lib.sb200_iterate(h.ptr(), a.data_ptr(), b.data_ptr(), N, st))
            run()
            torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                torch.cuda.synchronize()
                import time
                t0 = time.perf_counter()
                run()
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
            med = float(np.median(ts))
            print(f"{shape[0]}x{shape[1]} SB200_LIFE_PACKED={packed} {name:14s}: {med:8.3f} ms per 1000 generations, {cells * N / med / 1e6:9.1f} Gcell-updates/s, "
                  f"last kernel {lib.sb200_last_kernel().decode()}", flush=True)
