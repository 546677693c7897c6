#!/usr/bin/env python
"""One-GPU probe: the slab plan with ONE slab (its own ring neighbour) against sb200_iterate on the same box — separates the cost of
the plan's sweep form + exchange from multi-GPU effects (power, skew). tools/plan_probe.py [diffusion|life]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import stencils_b200 as sb  # noqa: E402
from stencils_b200 import _abi as A  # noqa: E402
from stencils_b200.slab import SlabPlan  # noqa: E402
from stencils_b200.synth import synth_torch  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "diffusion"
dev = torch.device("cuda", 0)
if wl == "diffusion":
    shape, dt, st, red, kw, et, bcs, steps = (1024, 1024, 1024), np.float32, sb.VonNeumann(1, 3), A.DIFFUSION, dict(alpha=0.1), A.F32, (A.WRAP,) * 3, 100
    f = sb.Diffusion(0.1)
else:
    shape, dt, st, red, kw, et, bcs, steps = (16384, 16384), np.uint8, sb.Moore(1), A.LIFE, dict(born_mask=8, survive_mask=12), A.U8, (A.WRAP,) * 2, 512
    f = sb.Life()
cells = int(np.prod(shape))
field = synth_torch(shape, dt, 0x5EED0005, dev)
lib = A.lib()


def med(xs):
    return float(np.median(xs))


S = sb.SwitchingStencilArray(field, st, boundary=sb.Wrap())
S = sb.iterate_(f, S, 8)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    S = sb.iterate_(f, S, steps)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"sb200_iterate            : {cells * steps / med(ts) / 1e6:9.1f} Gcell-updates/s  ({med(ts) / steps:.5f} ms/step, min {min(ts) / steps:.5f})", flush=True)
del S
torch.cuda.empty_cache()
for ghost, flags, name in ((0, 0, "default"), (0, A.PLAN_OVERLAP_OFF, "no overlap"), (8, 0, "G=8"), (0, A.PLAN_SINGLE_STEP, "single step")):
    if wl != "diffusion" and name == "G=8":
        ghost = 32
    plan = SlabPlan(shape, offsets=st.offsets(), radius=1, reducer=red, boundary=bcs, eltype=et, ghost=ghost, devices=[0], reducer_kwargs=kw,
                    plan_flags=flags)
    lo, hi, _, ptr = plan.slab(0)
    A.check(lib.sb200_memcpy_d2d(ptr, field.data_ptr(), cells * field.element_size(), None))
    A.check(lib.sb200_stream_sync(None))
    plan.mark_dirty()
    plan.iterate(16)
    plan.sync()
    ts = [plan.iterate_timed(steps) for _ in range(10)]
    print(f"plan, one slab, {name:12s}: {cells * steps / med(ts) / 1e6:9.1f} Gcell-updates/s  ({med(ts) / steps:.5f} ms/step, min {min(ts) / steps:.5f})  {plan.stats()}",
          flush=True)
    plan.close()
