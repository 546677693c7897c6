"""sb200_plan_* (the slab-partitioned SwitchingStencilArray loop behind the C ABI, include/stencils_b200.h) on the GPU,
torch-free: ctypes + NumPy only. Several slabs are placed on ONE device (devices = [0, 0, ...]) so that the whole protocol —
mailbox slots, fused mirror stores, event- and flag-ordered exchange, boundary-first overlap, Remove / Reflect ends, several
generations per launch — runs on the driver's one-GPU box; with two or more devices the same cases also run across devices.
Every result is compared bit for bit with the CPU oracle's single-domain iteration AND with sb200_iterate on the undivided
array."""
import ctypes as C

import numpy as np
import pytest

from oracle import np_restatement as npr
from stencils_b200 import _abi as A
from stencils_b200._desc import build_desc
from stencils_b200.slab import SlabPlan
from tests.util import bits_equal

pytestmark = pytest.mark.gpu


def ndev():
    n = C.c_int32()
    A.check(A.lib().sb200_device_count(C.byref(n)))
    return n.value


def single_domain(h, full, nsteps):
    """sb200_iterate on the undivided array through raw device buffers (sb200_malloc / memcpy)."""
    l = A.lib()
    nbytes = full.nbytes
    a, b = C.c_void_p(), C.c_void_p()
    A.check(l.sb200_malloc(C.byref(a), nbytes))
    A.check(l.sb200_malloc(C.byref(b), nbytes))
    try:
        src = np.asfortranarray(full)
        A.check(l.sb200_memcpy_h2d(a, src.ctypes.data, nbytes, None))
        A.check(l.sb200_memset(b, 0, nbytes, None))
        A.check(l.sb200_iterate(h.ptr(), a, b, nsteps, None))
        out = np.empty_like(src, order="F")
        A.check(l.sb200_memcpy_d2h(out.ctypes.data, a if nsteps % 2 == 0 else b, nbytes, None))
        A.check(l.sb200_stream_sync(None))
        return out
    finally:
        l.sb200_free(a)
        l.sb200_free(b)


def setup(name, shape, bcs):
    rng = np.random.default_rng(17)
    if name == "life":
        full = np.asfortranarray(((rng.random(shape) < 0.4) * rng.integers(1, 255, size=shape)).astype(np.uint8))
        kw = dict(eltype=A.U8, offsets=npr.offsets("Moore", 1, 2), radius=1, reducer=A.LIFE, reducer_kwargs=dict(born_mask=8, survive_mask=12), padval=1)
    elif name == "lifebool":
        full = np.asfortranarray(rng.random(shape) < 0.4)
        kw = dict(eltype=A.BOOL, offsets=npr.offsets("Moore", 1, 2), radius=1, reducer=A.LIFE, reducer_kwargs=dict(born_mask=8, survive_mask=12), padval=0)
    elif name == "diffusion":
        full = np.asfortranarray(rng.random(shape).astype(np.float32))
        kw = dict(eltype=A.F32, offsets=npr.offsets("VonNeumann", 1, 3), radius=1, reducer=A.DIFFUSION, reducer_kwargs=dict(alpha=0.1), padval=0.5)
    elif name == "diffusion64":
        full = np.asfortranarray(rng.random(shape))
        kw = dict(eltype=A.F64, offsets=npr.offsets("VonNeumann", 1, 3), radius=1, reducer=A.DIFFUSION, reducer_kwargs=dict(alpha=0.05), padval=0.25)
    elif name == "max":   # a reducer without multi-generation kernels, R = 2: Circle(2) running maximum
        full = np.asfortranarray(rng.random(shape).astype(np.float32))
        kw = dict(eltype=A.F32, offsets=npr.offsets("Circle", 2, 2), radius=2, reducer=A.MAX, reducer_kwargs={}, padval=0.0)
    else:
        raise AssertionError(name)
    return full, kw


CASES = [
    # name, shape, bcs, ghost (0 = library default), nslabs, plan_flags, step chunks
    ("life", (1024, 480), (A.WRAP, A.WRAP), 0, 3, 0, (70, 9, 33)),                         # G = 32, eight generations per launch
    ("life", (1024, 401), (A.WRAP, A.WRAP), 8, 2, A.PLAN_OVERLAP_ON, (40, 7)),            # ragged, overlap + mirror fallback copy
    ("life", (1024, 400), (A.WRAP, A.WRAP), 16, 2, A.PLAN_FLAGS_SYNC, (50,)),              # flag-ordered exchange on one device
    ("life", (1024, 400), (A.WRAP, A.WRAP), 16, 2, A.PLAN_FLAGS_SYNC | A.PLAN_OVERLAP_ON, (37, 12)),
    ("life", (1024, 300), (A.WRAP, A.WRAP), 8, 1, 0, (21,)),                                # one slab = its own neighbour
    ("lifebool", (512, 200), (A.WRAP, A.REMOVE), 4, 3, 0, (11, 6)),
    ("life", (528, 210), (A.REFLECT, A.REFLECT), 3, 2, A.PLAN_OVERLAP_ON, (10,)),           # single-generation kernel with fused mirror
    ("life", (1024, 400), (A.WRAP, A.WRAP), 4, 2, A.PLAN_SINGLE_STEP | A.PLAN_OVERLAP_ON, (13,)),
    ("diffusion", (64, 24, 90), (A.WRAP, A.WRAP, A.WRAP), 0, 3, A.PLAN_OVERLAP_ON, (14, 3, 8)),   # two steps per launch, overlap
    ("diffusion", (64, 24, 64), (A.WRAP, A.WRAP, A.WRAP), 4, 2, A.PLAN_OVERLAP_OFF, (9,)),
    ("diffusion", (64, 24, 64), (A.WRAP, A.WRAP, A.WRAP), 2, 2, A.PLAN_FLAGS_SYNC | A.PLAN_OVERLAP_ON, (12,)),
    ("diffusion", (64, 20, 61), (A.REMOVE, A.WRAP, A.REFLECT), 2, 3, A.PLAN_OVERLAP_ON, (7, 2)),
    ("diffusion64", (32, 20, 50), (A.WRAP, A.REFLECT, A.REMOVE), 1, 2, A.PLAN_OVERLAP_ON, (5,)),
    ("max", (512, 120), (A.WRAP, A.WRAP), 4, 3, A.PLAN_OVERLAP_ON, (5, 4)),
]


def run_case(orc, case, devices):
    name, shape, bcs, ghost, nslabs, pflags, chunks = case
    full, kw = setup(name, shape, bcs)
    rk = dict(kw)
    reducer_kwargs = rk.pop("reducer_kwargs")
    padval = rk.pop("padval")
    h = build_desc(size=shape, out_eltype=rk["eltype"], boundary=bcs, padval=padval, **rk, **reducer_kwargs)
    plan = SlabPlan(shape, boundary=bcs, ghost=ghost, devices=devices, plan_flags=pflags, reducer_kwargs=reducer_kwargs, padval=padval, **rk)
    try:
        assert plan.nslabs() == nslabs
        plan.load(full)
        total = 0
        for n in chunks:
            plan.iterate(n)
            total += n
        plan.sync()
        got = plan.store()
        st = plan.stats()
    finally:
        plan.close()
    want = orc.iterate(h, full.copy(order="F"), np.zeros_like(full, order="F"), total)
    bits_equal(got, want)
    bits_equal(got, single_domain(h, full, total))
    assert st["generations"] == total and st["exchanges"] >= total // st["generations_per_exchange"]
    return st


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-{c[2]}-g{c[3]}-s{c[4]}-f{c[5]}")
def test_plan_slabs_on_one_device_match_single_domain(orc, case):
    st = run_case(orc, case, [0] * case[4])
    name, pflags = case[0], case[5]
    assert st["sync"] == ("flags" if pflags & A.PLAN_FLAGS_SYNC else "events")
    if name == "life" and case[2] == (A.WRAP, A.WRAP) and not (pflags & A.PLAN_SINGLE_STEP) and st["ghost_planes"] >= 8:
        assert st["max_generations_per_launch"] == 8, st
    if name == "diffusion" and case[2][-1] == A.WRAP:
        assert st["max_generations_per_launch"] == 2, st


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-{c[2]}-g{c[3]}-s{c[4]}-f{c[5]}")
def test_plan_slabs_across_devices_match_single_domain(orc, case):
    n = ndev()
    if n < 2:
        pytest.skip("needs two or more CUDA devices")
    run_case(orc, case, [i % n for i in range(case[4])])


def test_plan_rank_form_world_one_and_errors(orc):
    """The rank form with world = 1 (its own ring neighbour, flag-ordered), direct device initialisation through
    sb200_plan_slab + sb200_plan_mark_dirty, iterate_timed, and the argument errors."""
    l = A.lib()
    shape, bcs = (1024, 256), (A.WRAP, A.WRAP)
    full, kw = setup("life", shape, bcs)
    rk = dict(kw)
    reducer_kwargs = rk.pop("reducer_kwargs")
    padval = rk.pop("padval")
    h = build_desc(size=shape, out_eltype=rk["eltype"], boundary=bcs, padval=padval, **rk, **reducer_kwargs)
    plan = SlabPlan(shape, boundary=bcs, ghost=8, rank=0, world=1, reducer_kwargs=reducer_kwargs, padval=padval, **rk)
    try:
        lo, hi, dev, ptr = plan.slab(0)
        assert (lo, hi) == (0, shape[1])
        A.check(l.sb200_memcpy_h2d(ptr, full.ctypes.data, full.nbytes, None))
        A.check(l.sb200_stream_sync(None))
        plan.mark_dirty()
        ms = plan.iterate_timed(27)
        assert ms > 0
        got = plan.store()
        assert plan.stats()["sync"] == "flags"
    finally:
        plan.close()
    bits_equal(got, orc.iterate(h, full.copy(order="F"), np.zeros_like(full, order="F"), 27))
    # errors: slab thinner than the ghost zone, ghost not a multiple of the radius, Halo-padded parents, bad device
    with pytest.raises(A.ArgumentError):
        SlabPlan((1024, 40), boundary=bcs, ghost=32, devices=[0, 0], reducer_kwargs=reducer_kwargs, padval=padval, **rk)
    full2, kw2 = setup("max", (512, 64), bcs)
    rk2 = dict(kw2)
    rkw2 = rk2.pop("reducer_kwargs")
    pv2 = rk2.pop("padval")
    with pytest.raises(A.ArgumentError):
        SlabPlan((512, 64), boundary=bcs, ghost=3, devices=[0], reducer_kwargs=rkw2, padval=pv2, **rk2)
    with pytest.raises(A.ArgumentError):
        SlabPlan((512, 64), boundary=bcs, ghost=2, devices=[99], reducer_kwargs=rkw2, padval=pv2, **rk2)
    with pytest.raises(A.ArgumentError):
        SlabPlan((512, 64), boundary=(A.WRAP, A.USE), ghost=2, devices=[0], reducer_kwargs=rkw2, padval=pv2, **rk2)


def test_plan_rank_form_two_processes():
    """One process per GPU (torchrun): IPC mailboxes + flag-ordered exchange, every rank compares its slab with sb200_iterate
    on the undivided array (tests/multigpu_check.py --plan-only). Skipped below two devices."""
    import os
    import subprocess
    import sys
    n = ndev()
    if n < 2:
        pytest.skip("needs two or more CUDA devices")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tests", "multigpu_check.py"), "--plan-only"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0 and "MULTIGPU CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
