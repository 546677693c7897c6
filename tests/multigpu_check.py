#!/usr/bin/env python
"""N-GPU == 1-GPU bitwise check for the slab iterator (run under torchrun on a multi-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py

Every rank also runs the whole domain on its own GPU through sb200_iterate and compares its slab bit for bit.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import stencils_b200 as sb  # noqa: E402
from stencils_b200 import _abi as A  # noqa: E402
from stencils_b200.slab import SlabIterator, split_axis_last  # noqa: E402
from stencils_b200.synth import synth_torch  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    dev = torch.device("cuda", torch.cuda.current_device())
    bad_total = 0
    cases = [
        ("life", (4096, 1024 * world), np.uint8, (A.WRAP, A.WRAP), 16, 40),
        ("life", (2048, 512 * world + 3), np.uint8, (A.REFLECT, A.REMOVE), 2, 9),
        ("diffusion", (256, 192, 64 * world), np.float32, (A.WRAP, A.WRAP, A.WRAP), 4, 22),
        ("diffusion", (128, 100, 40 * world + 1), np.float32, (A.REMOVE, A.WRAP, A.REFLECT), 2, 7),
    ]
    cases = [c + (ex,) for ex in ("p2p-fused", "p2p", "nccl") for c in cases]
    for name, shape, dt, bcs, ghost, nsteps, ex in cases:
        full = synth_torch(shape, dt, 0xABC, dev)                      # logical (column-major) view
        if name == "life":
            st, red, kw, et, f = sb.Moore(1), A.LIFE, dict(born_mask=8, survive_mask=12), A.U8, sb.Life()
        else:
            st, red, kw, et, f = sb.VonNeumann(1, 3), A.DIFFUSION, dict(alpha=0.1), A.F32, sb.Diffusion(0.1)
        bc_cls = {A.WRAP: sb.Wrap, A.REFLECT: sb.Reflect}
        lo, hi = split_axis_last(shape, world, rank)
        tfull = full.permute(*reversed(range(len(shape))))             # split axis first, C-contiguous
        it = SlabIterator(tfull[lo:hi].contiguous(), offsets=st.offsets(), radius=1, reducer=red, boundary=bcs, eltype=et,
                          ghost=ghost, rank=rank, world=world, reducer_kwargs=kw, padval=0, exchange=ex.split("-")[0])
        it.fused = ex == "p2p-fused"
        it.step(nsteps)
        # single-domain reference on this GPU (per-axis boundaries -> descriptor directly)
        from stencils_b200._desc import build_desc
        h = build_desc(size=shape, eltype=et, out_eltype=et, offsets=st.offsets(), radius=1, boundary=bcs, reducer=red,
                       padval=0, **kw)
        a = tfull.contiguous().clone()
        b = torch.empty_like(a)
        A.check(A.lib().sb200_iterate(h.ptr(), a.data_ptr(), b.data_ptr(), nsteps, torch.cuda.current_stream().cuda_stream))
        ref = a if nsteps % 2 == 0 else b
        torch.cuda.synchronize()
        bad = (it.state.view(torch.uint8) != ref[lo:hi].view(torch.uint8)).sum()
        dist.all_reduce(bad)
        if rank == 0:
            print(f"{name} {shape} bcs={bcs} ghost={ghost} steps={nsteps} exchange={ex}: mismatching bytes = {int(bad)}", flush=True)
        bad_total += int(bad)
    dist.destroy_process_group()
    if bad_total:
        sys.exit(1)
    if rank == 0:
        print("MULTIGPU_OK", flush=True)


if __name__ == "__main__":
    main()
