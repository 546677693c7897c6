#!/usr/bin/env python
"""N-GPU == 1-GPU bitwise check for the slab iterator (run under torchrun on a multi-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py [--quick]

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
from stencils_b200.slab import (SlabIterator, slab_gather, slab_scatter, split_axis_last,  # noqa: E402
                                 split_columns_for_scatter)
from stencils_b200.synth import synth_torch  # noqa: E402


def one_shot_cases(rank, world, dev):
    """slab_gather / slab_scatter (one exchange + one sweep) against the single-domain call on this GPU."""
    from stencils_b200._desc import build_desc
    bad_total = 0
    stream = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    # gathers: configs[0], [2], [3]a shapes at reduced size
    w7 = np.random.default_rng(3).random(49).astype(np.float32)
    for name, shape, dt, et, stn, red, bcs, kw in [
            ("mean Window(1) F64", (2048, 300 * world + 1), np.float64, A.F64, sb.Window(1), A.MEAN, (A.REMOVE, A.REMOVE), {}),
            ("kernelproduct Window(3) F32", (2048, 256 * world), np.float32, A.F32, sb.Window(3), A.KERNELDOT, (A.REMOVE, A.REMOVE), dict(weights=w7)),
            ("maximum Circle(4) F32", (4096, 200 * world), np.float32, A.F32, sb.Circle(4), A.MAX, (A.WRAP, A.WRAP), {})]:
        full = synth_torch(shape, dt, 0xDEF, dev)
        tfull = full.permute(1, 0)
        lo, hi = split_axis_last(shape, world, rank)
        got = slab_gather(tfull[lo:hi].contiguous(), offsets=stn.offsets(), radius=stn.radius, reducer=red, boundary=bcs, eltype=et,
                          rank=rank, world=world, reducer_kwargs=kw, padval=0)
        h = build_desc(size=shape, eltype=et, out_eltype=et, offsets=stn.offsets(), radius=stn.radius, boundary=bcs, reducer=red, padval=0, **kw)
        a = tfull.contiguous()
        ref = torch.empty_like(a)
        A.check(A.lib().sb200_gather(h.ptr(), a.data_ptr(), ref.data_ptr(), stream()))
        torch.cuda.synchronize()
        bad = (got.view(torch.uint8) != ref[lo:hi].view(torch.uint8)).sum()
        dist.all_reduce(bad)
        if rank == 0:
            print(f"slab_gather {name} {shape} bcs={bcs}: mismatching bytes = {int(bad)}", flush=True)
        bad_total += int(bad)
    # scatter: configs[3]b shape at reduced size
    offs = [(-1, 1), (-2, -1), (1, 0), (-2, 2)]
    w = np.array([0.4, 0.3, 0.2, 0.1], dtype=np.float32)
    for shape, bcs in [((4096, 300 * world + 2), (A.REMOVE, A.REMOVE)), ((2048, 250 * world), (A.WRAP, A.WRAP)),
                       ((1000, 123 * world), (A.REFLECT, A.REFLECT))]:
        tsrc = synth_torch(shape, np.float32, 0x123, dev).permute(1, 0).contiguous()
        tdst = synth_torch(shape, np.float32, 0x456, dev).permute(1, 0).contiguous()
        lo, hi = split_columns_for_scatter(shape[1], world, rank, 2)
        mine = tdst[lo:hi].clone()
        slab_scatter(tsrc[lo:hi].contiguous(), mine, ncols_global=shape[1], offsets=offs, radius=2, weights=w, boundary=bcs,
                     eltype=A.F32, rank=rank, world=world)
        h = build_desc(size=shape, eltype=A.F32, out_eltype=A.F32, offsets=offs, radius=2, boundary=bcs, weights=w,
                       scatter_op=A.OP_ADD, scatter_rule=A.SCATTER_CENTER_WEIGHTS)
        ref = tdst.clone()
        A.check(A.lib().sb200_scatter(h.ptr(), tsrc.data_ptr(), ref.data_ptr(), stream()))
        torch.cuda.synchronize()
        bad = (mine.view(torch.uint8) != ref[lo:hi].view(torch.uint8)).sum()
        dist.all_reduce(bad)
        if rank == 0:
            print(f"slab_scatter Positional + {shape} bcs={bcs}: mismatching bytes = {int(bad)}", flush=True)
        bad_total += int(bad)
    return bad_total


def plan_cases(rank, world, dev):
    """The C-ABI slab plan in its rank form (sb200_plan_create_rank / _connect: one process per GPU, IPC mailboxes, flag-ordered
    exchange) against sb200_iterate on the undivided array on this GPU, bit for bit."""
    from stencils_b200._desc import build_desc
    from stencils_b200.slab import SlabPlan
    bad_total = 0
    cases = [
        ("life", (4096, 1024 * world), np.uint8, (A.WRAP, A.WRAP), 0, 0, (70, 9)),                      # G = 32: 8 generations per launch
        ("life", (2048, 512 * world + 3), np.uint8, (A.REFLECT, A.REMOVE), 2, A.PLAN_OVERLAP_ON, (9,)),  # fused mirror store, ragged
        ("diffusion", (256, 192, 64 * world), np.float32, (A.WRAP, A.WRAP, A.WRAP), 0, 0, (22, 5)),      # G = 4, two steps per launch, overlap
        ("diffusion", (256, 192, 64 * world), np.float32, (A.WRAP, A.WRAP, A.WRAP), 4, A.PLAN_SINGLE_STEP, (13,)),   # stream3d MIRROR store
        ("diffusion", (128, 100, 40 * world + 1), np.float32, (A.REMOVE, A.WRAP, A.REFLECT), 2, A.PLAN_OVERLAP_ON, (7,)),
    ]
    for name, shape, dt, bcs, ghost, pflags, chunks in cases:
        full = synth_torch(shape, dt, 0xABC, dev)
        if name == "life":
            st, red, kw, et = sb.Moore(1), A.LIFE, dict(born_mask=8, survive_mask=12), A.U8
        else:
            st, red, kw, et = sb.VonNeumann(1, 3), A.DIFFUSION, dict(alpha=0.1), A.F32
        tfull = full.permute(*reversed(range(len(shape)))).contiguous()   # split axis first == column-major memory
        plan = SlabPlan(shape, offsets=st.offsets(), radius=1, reducer=red, boundary=bcs, eltype=et, ghost=ghost, rank=rank, world=world,
                        reducer_kwargs=kw, padval=0, plan_flags=pflags)
        try:
            lo, hi, _, ptr = plan.slab(0)
            mine = tfull[lo:hi].contiguous()
            A.check(A.lib().sb200_memcpy_d2d(ptr, mine.data_ptr(), mine.numel() * mine.element_size(), None))
            A.check(A.lib().sb200_stream_sync(None))
            plan.mark_dirty()
            for n in chunks:
                plan.iterate(n)
            plan.sync()
            lo, hi, _, ptr = plan.slab(0)
            got = torch.empty_like(mine)
            A.check(A.lib().sb200_memcpy_d2d(got.data_ptr(), ptr, got.numel() * got.element_size(), None))
            A.check(A.lib().sb200_stream_sync(None))
            stats = plan.stats()
            dist.barrier()
        finally:
            plan.close()
        nsteps = sum(chunks)
        h = build_desc(size=shape, eltype=et, out_eltype=et, offsets=st.offsets(), radius=1, boundary=bcs, reducer=red, padval=0, **kw)
        a = tfull.clone()
        b = torch.empty_like(a)
        A.check(A.lib().sb200_iterate(h.ptr(), a.data_ptr(), b.data_ptr(), nsteps, torch.cuda.current_stream().cuda_stream))
        ref = a if nsteps % 2 == 0 else b
        torch.cuda.synchronize()
        bad = (got.view(torch.uint8) != ref[lo:hi].view(torch.uint8)).sum()
        dist.all_reduce(bad)
        if rank == 0:
            print(f"plan {name} {shape} bcs={bcs} steps={chunks} {stats}: mismatching bytes = {int(bad)}", flush=True)
        bad_total += int(bad)
    return bad_total


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    dev = torch.device("cuda", torch.cuda.current_device())
    bad_total = plan_cases(rank, world, dev)
    if "--plan-only" in sys.argv:
        if rank == 0:
            print("MULTIGPU CHECK " + ("PASSED" if bad_total == 0 else f"FAILED ({bad_total} bytes)"), flush=True)
        dist.destroy_process_group()
        sys.exit(0 if bad_total == 0 else 1)
    cases = [
        ("life", (4096, 1024 * world), np.uint8, (A.WRAP, A.WRAP), 16, 40),
        ("life", (2048, 512 * world + 3), np.uint8, (A.REFLECT, A.REMOVE), 2, 9),
        ("diffusion", (256, 192, 64 * world), np.float32, (A.WRAP, A.WRAP, A.WRAP), 4, 22),
        ("diffusion", (128, 100, 40 * world + 1), np.float32, (A.REMOVE, A.WRAP, A.REFLECT), 2, 7),
    ]
    cases = [c + (ex,) for ex in ("p2p-fused", "p2p", "nccl") for c in cases]
    if "--quick" in sys.argv:   # one Life and one diffusion ring over fused peer stores, then the one-shot sweeps
        cases = [cases[0], cases[2], cases[3]]
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
    bad_total += one_shot_cases(rank, world, dev)
    # two diffusion steps per launch inside the slab iterator (csrc/stream3d2.cu)
    os.environ["SB200_DIFFUSION_DOUBLE_STEP"] = "1"
    from stencils_b200._desc import build_desc
    shape, bcs, nsteps = (256, 192, 64 * world), (A.WRAP, A.WRAP, A.WRAP), 22
    st = sb.VonNeumann(1, 3)
    tfull = synth_torch(shape, np.float32, 0xABD, dev).permute(2, 1, 0)
    lo, hi = split_axis_last(shape, world, rank)
    it = SlabIterator(tfull[lo:hi].contiguous(), offsets=st.offsets(), radius=1, reducer=A.DIFFUSION, boundary=bcs, eltype=A.F32,
                      ghost=4, rank=rank, world=world, reducer_kwargs=dict(alpha=0.1), padval=0)
    A.lib().sb200_launch_count(1)
    it.step(nsteps)
    launches = A.lib().sb200_launch_count(1)
    os.environ["SB200_DIFFUSION_DOUBLE_STEP"] = "0"
    h = build_desc(size=shape, eltype=A.F32, out_eltype=A.F32, offsets=st.offsets(), radius=1, boundary=bcs, reducer=A.DIFFUSION, alpha=0.1)
    a = tfull.contiguous().clone()
    b = torch.empty_like(a)
    A.check(A.lib().sb200_iterate(h.ptr(), a.data_ptr(), b.data_ptr(), nsteps, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    bad = (it.state.view(torch.uint8) != (a if nsteps % 2 == 0 else b)[lo:hi].view(torch.uint8)).sum()
    dist.all_reduce(bad)
    if rank == 0:
        print(f"diffusion {shape} two steps per launch ({launches} launches for {nsteps} steps): mismatching bytes = {int(bad)}", flush=True)
    bad_total += int(bad)
    it.close()
    dist.destroy_process_group()
    if bad_total:
        sys.exit(1)
    if rank == 0:
        print("MULTIGPU_OK", flush=True)


if __name__ == "__main__":
    main()
