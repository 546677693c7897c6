"""The cycle schedule of the C-ABI slab plans (csrc/slab_sched.h, exported as sb200_slab_schedule) on the CPU: the op list is
interpreted over 1-3 simulated slabs with the CPU oracle as the sweep — mailbox slots, exchange parity, fused mirror planes,
boundary-first overlap, Remove / Reflect ends, several generations per launch — and the result must equal the single-domain
iteration bit for bit. The GPU executor (csrc/slab_plan.cu) interprets the SAME list (tests/test_gpu_plan.py)."""
import numpy as np
import pytest

from oracle import np_restatement as npr
from stencils_b200 import _abi as A
from stencils_b200._desc import build_desc
from stencils_b200.slab import slab_schedule, split_axis_last


class SimSlab:
    def __init__(self, full, lo, hi, G, up, down):
        self.n, self.G = hi - lo, G
        self.ext = self.n + 2 * G
        rest = full.shape[:-1]
        poison = 7 if full.dtype == np.uint8 else np.nan
        self.buf = [np.full(rest + (self.ext,), poison, dtype=full.dtype, order="F") for _ in range(2)]
        self.buf[0][..., G:G + self.n] = full[..., lo:hi]
        self.cur = 0
        self.slots = [[None, None], [None, None]]   # [side][parity]; None = empty / consumed
        self.seq = 0
        self.up, self.down = up, down


def run_ops(orc, slabs, ops, kw, bcs, padval):
    nd = len(bcs)
    for o in ops:
        for s in slabs:
            G, n, ext = s.G, s.n, s.ext
            cur, nxt = s.buf[s.cur], s.buf[1 - s.cur]
            absp = lambda v: v if v >= 0 else ext + v   # noqa: E731
            if o["kind"] == A.SLAB_SWEEP:
                lo, hi = absp(o["lo"]), absp(o["hi"])
                if hi <= lo:
                    continue
                size = cur.shape
                h = build_desc(size=size, boundary=tuple(bcs[:-1]) + (A.WRAP,), padval=padval, **kw)
                t = cur
                for _ in range(o["gens"]):
                    t = orc.gather(h, np.asfortranarray(t), np.zeros(size, dtype=cur.dtype, order="F"))
                # only cells whose whole dependency cone lies inside the parent are meaningful
                assert lo >= kw["radius"] * o["gens"] and hi <= ext - kw["radius"] * o["gens"]
                nxt[..., lo:hi] = t[..., lo:hi]
                par = (s.seq + 1) & 1
                if o["mirror"] == A.SLAB_MIRROR_DOWN and s.down is not None:
                    assert lo <= G and hi >= 2 * G
                    slabs[s.down].slots[1][par] = nxt[..., G:2 * G].copy()
                if o["mirror"] == A.SLAB_MIRROR_UP and s.up is not None:
                    assert lo <= n and hi >= n + G
                    slabs[s.up].slots[0][par] = nxt[..., n:n + G].copy()
            elif o["kind"] == A.SLAB_PUSH:
                b = cur if o["buf"] == A.SLAB_CUR else nxt
                par = (s.seq + 1) & 1
                if s.up is not None:
                    slabs[s.up].slots[0][par] = b[..., n:n + G].copy()
                if s.down is not None:
                    slabs[s.down].slots[1][par] = b[..., G:2 * G].copy()
            elif o["kind"] == A.SLAB_SIGNAL:
                s.seq += 1
            elif o["kind"] == A.SLAB_PULL:
                b = cur if o["buf"] == A.SLAB_CUR else nxt
                par = s.seq & 1
                if s.down is not None:
                    assert s.slots[0][par] is not None, "pull from an empty slot"
                    b[..., :G] = s.slots[0][par]
                    s.slots[0][par] = None
                if s.up is not None:
                    assert s.slots[1][par] is not None, "pull from an empty slot"
                    b[..., G + n:] = s.slots[1][par]
                    s.slots[1][par] = None
                end_fill(s, b, bcs[-1], padval)
            elif o["kind"] == A.SLAB_ENDFILL:
                end_fill(s, cur if o["buf"] == A.SLAB_CUR else nxt, bcs[-1], padval)
            elif o["kind"] == A.SLAB_SWAP:
                s.cur = 1 - s.cur
            else:
                assert o["kind"] == A.SLAB_JOIN


def end_fill(s, b, bc, padval):
    G, n = s.G, s.n
    if bc == A.WRAP:
        return
    if bc == A.REMOVE:
        if s.down is None:
            b[..., :G] = padval
        if s.up is None:
            b[..., G + n:] = padval
    else:
        if s.down is None:
            b[..., :G] = b[..., G + 1:2 * G + 1][..., ::-1]
        if s.up is None:
            b[..., G + n:] = b[..., n - 1:G + n - 1][..., ::-1]


def case_setup(name, shape, bcs):
    rng = np.random.default_rng(11)
    if name == "life":
        full = np.asfortranarray((rng.random(shape) < 0.4).astype(np.uint8))
        kw = dict(eltype=A.U8, out_eltype=A.U8, offsets=npr.offsets("Moore", 1, 2), radius=1, reducer=A.LIFE, born_mask=8, survive_mask=12)
        pad = 1
    else:
        full = np.asfortranarray(rng.random(shape).astype(np.float32))
        kw = dict(eltype=A.F32, out_eltype=A.F32, offsets=npr.offsets("VonNeumann", 1, 3), radius=1, reducer=A.DIFFUSION, alpha=0.1)
        pad = 0.5
    return full, kw, pad


CASES = [
    # name, shape, bcs, ghost, nslabs, overlap, max_gens, min_planes_multi, step chunks
    ("life", (24, 60), (A.WRAP, A.WRAP), 8, 3, False, 8, 0, (19, 8, 1, 4)),
    ("life", (24, 61), (A.WRAP, A.WRAP), 8, 2, True, 4, 0, (16, 5, 11)),       # ragged slabs, overlap with multi-generation sweeps
    ("life", (24, 64), (A.WRAP, A.WRAP), 8, 2, True, 8, 12, (17, 15)),         # thin boundary sweeps rejected -> no overlap on those steps
    ("life", (24, 40), (A.WRAP, A.WRAP), 4, 1, True, 4, 0, (9,)),              # one slab: its own neighbour on the ring
    # negative max_gens = -(mask of launch sizes), preference 7, 8, 6, 5, 4, 3, 2 (the Life plans since r02t)
    ("life", (24, 90), (A.WRAP, A.WRAP), 14, 3, False, -0x1FC, 0, (31, 14, 5, 9)),   # cycles of two launches of seven
    ("life", (24, 96), (A.WRAP, A.WRAP), 14, 2, True, -0x1FC, 0, (28, 17)),          # overlap on a seven-generation last sweep
    ("life", (24, 70), (A.WRAP, A.WRAP), 10, 2, False, -0x0E8, 0, (23, 10)),         # sizes 7, 6, 5, 3: 7 + 3 per cycle
    ("life", (20, 37), (A.WRAP, A.REMOVE), 2, 3, False, 1, 0, (5, 2)),
    ("life", (20, 37), (A.REFLECT, A.REFLECT), 3, 2, True, 1, 0, (7,)),
    ("diffusion", (8, 6, 30), (A.WRAP, A.WRAP, A.WRAP), 4, 3, True, 2, 0, (10, 3, 4)),
    ("diffusion", (8, 6, 30), (A.WRAP, A.WRAP, A.WRAP), 4, 2, False, 2, 0, (9,)),
    ("diffusion", (8, 6, 31), (A.REMOVE, A.WRAP, A.REFLECT), 2, 3, True, 1, 0, (6, 1)),
    ("diffusion", (8, 6, 26), (A.WRAP, A.REFLECT, A.REMOVE), 1, 2, True, 1, 0, (4,)),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-{c[2]}-g{c[3]}-s{c[4]}-ov{int(c[5])}-m{c[6]}")
def test_schedule_interpreted_with_the_oracle_matches_single_domain(orc, case):
    name, shape, bcs, G, nslabs, overlap, max_gens, min_multi, chunks = case
    full, kw, pad = case_setup(name, shape, bcs)
    wrap = bcs[-1] == A.WRAP
    slabs = []
    for r in range(nslabs):
        lo, hi = split_axis_last(shape, nslabs, r)
        up = (r + 1) % nslabs if (wrap or r < nslabs - 1) else None
        down = (r - 1) % nslabs if (wrap or r > 0) else None
        slabs.append(SimSlab(full, lo, hi, G, up, down))
    n_min = min(s.n for s in slabs)
    since, first, total = None, True, 0
    kinds = set()
    for nsteps in chunks:
        ops, since = slab_schedule(1, G, n_min, split_wrap=wrap, overlap=overlap, max_gens=max_gens, min_planes_multi=min_multi,
                                   since=since, first_sweep=first, nsteps=nsteps)
        kinds |= {(o["kind"], o["gens"], o["async_"]) for o in ops}
        run_ops(orc, slabs, ops, kw, bcs, pad)
        first = False
        total += nsteps
    h = build_desc(size=shape, boundary=bcs, padval=pad, **kw)
    want = orc.iterate(h, full.copy(order="F"), np.zeros_like(full, order="F"), total)
    got = np.concatenate([s.buf[s.cur][..., s.G:s.G + s.n] for s in slabs], axis=-1)
    u = np.uint8 if name == "life" else np.uint32
    np.testing.assert_array_equal(np.ascontiguousarray(got).view(u), np.ascontiguousarray(want).view(u))
    if max_gens > 1 and min_multi == 0:
        assert any(k == A.SLAB_SWEEP and g == max_gens for k, g, _ in kinds), "the multi-generation launches were never scheduled"
    if overlap and nslabs >= 1 and min_multi == 0 and n_min >= 2 * G + 2:
        assert any(k == A.SLAB_PULL and a for k, _, a in kinds), "the overlapped exchange was never scheduled"


def test_schedule_shapes():
    """Structure of one cycle: k generations per exchange, launches shrink by R * gens planes per side, the overlapped last
    sweep is boundary / boundary / signal / async pull / interior / join."""
    ops, since = slab_schedule(1, 4, 100, overlap=True, max_gens=2, nsteps=4)
    kinds = [o["kind"] for o in ops]
    assert kinds == [A.SLAB_PUSH, A.SLAB_SIGNAL, A.SLAB_PULL, A.SLAB_SWEEP, A.SLAB_SWAP,
                     A.SLAB_SWEEP, A.SLAB_SWEEP, A.SLAB_SIGNAL, A.SLAB_PULL, A.SLAB_SWEEP, A.SLAB_JOIN, A.SLAB_SWAP]
    sw = [o for o in ops if o["kind"] == A.SLAB_SWEEP]
    assert [(o["lo"], o["hi"], o["gens"], o["mirror"]) for o in sw] == [(2, -2, 2, 0), (4, 9, 2, A.SLAB_MIRROR_DOWN), (-9, -4, 2, A.SLAB_MIRROR_UP),
                                                                       (9, -9, 2, 0)]
    assert sw[0]["first"] == 1 and sw[1]["first"] == 0 and since == 0
    # the next cycle needs no blocking exchange: the ghosts arrived under the interior sweep
    ops2, since2 = slab_schedule(1, 4, 100, overlap=True, max_gens=2, since=since, first_sweep=False, nsteps=3)
    assert [o["kind"] for o in ops2][:2] == [A.SLAB_SWEEP, A.SLAB_SWAP] and since2 == 3
    assert [o["gens"] for o in ops2 if o["kind"] == A.SLAB_SWEEP] == [2, 1]
    # a slab too thin for boundary-first sweeps never overlaps
    ops3, _ = slab_schedule(1, 4, 9, overlap=True, max_gens=1, nsteps=8)
    assert not any(o["kind"] == A.SLAB_JOIN for o in ops3)
    with pytest.raises(A.ArgumentError):
        slab_schedule(2, 3, 100, nsteps=1)
