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
