"""Slab decomposition + ghost exchange logic on CPU: world_size 2 and 3 with the gloo backend, the sweep backend
injected (the CPU oracle stands in for the CUDA kernels), result compared with the single-domain sweep."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_compute(h, src, dst):
    from oracle import oracle as orc
    s = src.numpy()
    d = dst.numpy()
    # torch C-order (split axis first) == column-major parent of the reversed shape
    orc.gather(h, s.T, d.T)


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import np_restatement as npr
        from stencils_b200 import _abi as A
        from stencils_b200.slab import SlabIterator, split_axis_last
        name, shape, bcs, ghost, nsteps = case
        rng = np.random.default_rng(5)
        if name == "life":
            full = (rng.random(shape) < 0.4).astype(np.uint8)
            offs, red, kw, et = npr.offsets("Moore", 1, 2), A.LIFE, dict(born_mask=8, survive_mask=12), A.U8
        else:
            full = rng.random(shape).astype(np.float32)
            offs, red, kw, et = npr.offsets("VonNeumann", 1, 3), A.DIFFUSION, dict(alpha=0.1), A.F32
        lo, hi = split_axis_last(shape, world, rank)
        local = torch.from_numpy(np.ascontiguousarray(full[..., lo:hi].T))  # split axis first
        it = SlabIterator(local, offsets=offs, radius=1, reducer=red, boundary=bcs, eltype=et, ghost=ghost, rank=rank,
                          world=world, compute=_oracle_compute, reducer_kwargs=kw, padval=1 if name == "life" else 0.5)
        it.step(nsteps)
        np.save(os.path.join(out_dir, f"part{rank}.npy"), it.state.numpy().T)
    finally:
        dist.destroy_process_group()


CASES = [
    ("life", (48, 40), (1, 1), 2, 5),          # Wrap/Wrap ring, 2 steps per exchange, 5 steps
    ("life", (32, 37), (1, 0), 1, 4),          # Remove on the split axis (padval 1), ragged slabs
    ("diffusion", (16, 12, 21), (1, 2, 2), 3, 7),   # Reflect on the split axis
    ("diffusion", (16, 12, 18), (0, 1, 1), 2, 4),
]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-{c[2]}-g{c[3]}")
def test_slabs_match_single_domain(tmp_path, world, case, orc):  # `orc` builds the oracle before the workers load it
    from oracle import np_restatement as npr
    from stencils_b200 import _abi as A
    from stencils_b200._desc import build_desc
    port = 29500 + (os.getpid() + world * 7 + CASES.index(case)) % 2000
    mp.spawn(_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    name, shape, bcs, ghost, nsteps = case
    rng = np.random.default_rng(5)
    if name == "life":
        full = np.asfortranarray((rng.random(shape) < 0.4).astype(np.uint8))
        h = build_desc(size=shape, eltype=A.U8, out_eltype=A.U8, offsets=npr.offsets("Moore", 1, 2), radius=1, boundary=bcs,
                       reducer=A.LIFE, born_mask=8, survive_mask=12, padval=1)
    else:
        full = np.asfortranarray(rng.random(shape).astype(np.float32))
        h = build_desc(size=shape, eltype=A.F32, out_eltype=A.F32, offsets=npr.offsets("VonNeumann", 1, 3), radius=1,
                       boundary=bcs, reducer=A.DIFFUSION, alpha=0.1, padval=0.5)
    want = orc.iterate(h, full.copy(order="F"), np.zeros_like(full, order="F"), nsteps)
    got = np.concatenate([np.load(tmp_path / f"part{r}.npy") for r in range(world)], axis=-1)
    assert got.dtype == want.dtype
    np.testing.assert_array_equal(got.view(np.uint8 if name == "life" else np.uint32), want.view(np.uint8 if name == "life" else np.uint32))


# ---------------------------------------------------------------------------------------------- one-shot sweeps
def _oracle_scatter(h, src, dst):
    from oracle import oracle as orc
    orc.scatter(h, src.numpy().T, dst.numpy().T)


def _oneshot_worker(rank, world, port, case, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import np_restatement as npr
        from stencils_b200 import _abi as A
        from stencils_b200.slab import slab_gather, slab_scatter, split_axis_last, split_columns_for_scatter
        kind, shape, bcs, R, extra = case
        rng = np.random.default_rng(17)
        full = (rng.random(shape) - 0.3).astype(np.float32)
        if kind == "gather":
            shp, red = extra
            offs = npr.offsets(shp, R, len(shape))
            kw = dict(weights=np.random.default_rng(3).random(len(offs)).astype(np.float32)) if red == A.KERNELDOT else {}
            lo, hi = split_axis_last(shape, world, rank)
            local = torch.from_numpy(np.ascontiguousarray(full[..., lo:hi].T))
            out = slab_gather(local, offsets=offs, radius=R, reducer=red, boundary=bcs, eltype=A.F32, rank=rank, world=world,
                              compute=_oracle_compute, reducer_kwargs=kw, padval=0.25)
        else:
            offs, op = extra
            w = np.random.default_rng(4).random(len(offs)).astype(np.float32)
            dfull = (np.random.default_rng(5).random(shape) - 0.5).astype(np.float32)
            lo, hi = split_columns_for_scatter(shape[1], world, rank, R)
            local = torch.from_numpy(np.ascontiguousarray(full[:, lo:hi].T))
            out = torch.from_numpy(np.ascontiguousarray(dfull[:, lo:hi].T))
            slab_scatter(local, out, ncols_global=shape[1], offsets=offs, radius=R, weights=w, boundary=bcs, eltype=A.F32,
                         rank=rank, world=world, scatter_op=op, compute=_oracle_scatter)
        np.save(os.path.join(out_dir, f"part{rank}.npy"), out.numpy().T)
    finally:
        dist.destroy_process_group()


MEAN_, MAX_, KERNELDOT_ = 1, 3, 4   # sb200_reducer values (include/stencils_b200.h)
ONESHOT = [
    ("gather", (40, 31), (0, 0), 1, ("Window", MEAN_)),            # configs[0]: Window(1) mean, Remove
    ("gather", (36, 30), (0, 0), 3, ("Window", KERNELDOT_)),       # configs[2]: Kernel(Window(3)) kernelproduct, Remove
    ("gather", (33, 40), (2, 2), 4, ("Circle", MAX_)),             # configs[3]a: Circle(4) maximum (Reflect here)
    ("gather", (20, 12, 17), (1, 1, 1), 1, ("Moore", MEAN_)),      # 3-D Moore(1) mean, Wrap ring
    ("scatter", (30, 40), (0, 0), 2, ([(-1, 1), (-2, -1), (1, 0), (-2, 2)], 0)),   # configs[3]b: Positional scatter +, Remove
    ("scatter", (24, 45), (1, 1), 2, ([(-1, 1), (-2, -1), (1, 0), (-2, 2)], 0)),   # Wrap ring, 45 = 9 passes of 5 columns
    ("scatter", (25, 37), (2, 2), 1, ([(0, -1), (-1, 0), (1, 0), (0, 1), (1, 1)], 0)),  # Reflect, ragged column count
    ("scatter", (16, 33), (0, 0), 1, ([(-1, -1), (0, 1), (1, 1)], 1)),             # max
]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ONESHOT, ids=lambda c: f"{c[0]}-{c[1]}-{c[2]}-R{c[3]}")
def test_one_shot_slab_sweeps_match_single_domain(tmp_path, world, case, orc):
    """slab_gather (one pre-exchange + one sweep) and slab_scatter (ghost source columns, pass-aligned slabs) against
    the single-domain oracle sweep, bit for bit."""
    from oracle import np_restatement as npr
    from stencils_b200 import _abi as A
    from stencils_b200._desc import build_desc
    port = 31500 + (os.getpid() + world * 11 + ONESHOT.index(case)) % 2000
    mp.spawn(_oneshot_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    kind, shape, bcs, R, extra = case
    rng = np.random.default_rng(17)
    full = np.asfortranarray((rng.random(shape) - 0.3).astype(np.float32))
    if kind == "gather":
        shp, red = extra
        offs = npr.offsets(shp, R, len(shape))
        kw = dict(weights=np.random.default_rng(3).random(len(offs)).astype(np.float32)) if red == A.KERNELDOT else {}
        h = build_desc(size=shape, eltype=A.F32, out_eltype=A.F32, offsets=offs, radius=R, boundary=bcs, reducer=red, padval=0.25, **kw)
        want = orc.gather(h, full, np.zeros_like(full, order="F"))
    else:
        offs, op = extra
        w = np.random.default_rng(4).random(len(offs)).astype(np.float32)
        dfull = np.asfortranarray((np.random.default_rng(5).random(shape) - 0.5).astype(np.float32))
        h = build_desc(size=shape, eltype=A.F32, out_eltype=A.F32, offsets=offs, radius=R, boundary=bcs, weights=w,
                       scatter_op=op, scatter_rule=A.SCATTER_CENTER_WEIGHTS)
        want = orc.scatter(h, full, dfull.copy(order="F"))
    got = np.concatenate([np.load(tmp_path / f"part{r}.npy") for r in range(world)], axis=-1)
    np.testing.assert_array_equal(got.view(np.uint32), np.ascontiguousarray(want).view(np.uint32))
