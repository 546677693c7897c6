"""Torch-free driver for `compute-sanitizer` (memcheck / racecheck / synccheck / initcheck) runs of every kernel family.

    compute-sanitizer --tool memcheck  --error-exitcode 9 python tests/sanitize_cases.py
    compute-sanitizer --tool racecheck --error-exitcode 9 python tests/sanitize_cases.py --quick
    compute-sanitizer --tool synccheck --error-exitcode 9 python tests/sanitize_cases.py --quick

Every buffer is its own `sb200_malloc` (cudaMalloc) allocation of exactly the parent's size, so an access one byte
outside a parent is an error for memcheck (a caching allocator would hide it). Each case also checks the result
against the CPU oracle bit for bit and records which kernel the library dispatched to; kernel families listed in
WANT that were never exercised are reported. Test infrastructure: lives under tests/, may use oracle/.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import np_restatement as npr  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from stencils_b200 import _abi as A  # noqa: E402
from stencils_b200._desc import build_desc  # noqa: E402

WANT = ["life_bit_kernel<8", "life_bit_kernel<4", "life_bit_kernel<2", "life_tma2_kernel", "life_tma_kernel", "life_swar_kernel",
        "stream2d_kernel", "stream3d_kernel", "stream3d2_kernel", "gather_stream_kernel", "gather_stream3d_kernel", "box3d_kernel<window>", "box3d_kernel<moore>", "gather_generic",
        "scatter_stream_kernel", "scatter_fast_kernel", "halo_kernel"]
seen = {}
DRY = False  # --dry: no device calls, only the oracle side (checks the case table itself on a box without a GPU)


class Dev:
    """One cudaMalloc allocation holding a column-major NumPy parent."""

    def __init__(self, a: np.ndarray):
        self.shape, self.dtype = a.shape, a.dtype
        self.nbytes = max(a.nbytes, 1)
        self.p = C.c_void_p()
        if DRY:
            return
        A.check(A.lib().sb200_malloc(C.byref(self.p), self.nbytes))
        host = np.asfortranarray(a)
        A.check(A.lib().sb200_memcpy_h2d(self.p, host.ctypes.data, a.nbytes, None))
        A.check(A.lib().sb200_stream_sync(None))

    def get(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=self.dtype, order="F")
        if DRY:
            return None
        A.check(A.lib().sb200_memcpy_d2h(out.ctypes.data, self.p, out.nbytes, None))
        A.check(A.lib().sb200_stream_sync(None))
        return out

    def free(self):
        if not DRY:
            A.check(A.lib().sb200_free(self.p))


class _DryLib:
    """Stands in for the CUDA library under --dry: every call succeeds and does nothing."""

    def __getattr__(self, name):
        if name == "sb200_last_kernel":
            return lambda: b"(dry)"
        return lambda *a: 0


def _lib():
    return _DryLib() if DRY else A.lib()


def same(a, b, what):
    if a is None:
        return
    u = {1: np.uint8, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize]
    av, bv = np.ascontiguousarray(a).view(u), np.ascontiguousarray(b).view(u)
    if a.dtype.kind == "f":
        ok = (av == bv) | (np.isnan(a) & np.isnan(b))
    else:
        ok = av == bv
    if not ok.all():
        raise AssertionError(f"{what}: {(~ok).sum()} of {a.size} cells differ from the oracle")


def note(name, what):
    k = _lib().sb200_last_kernel().decode()
    seen.setdefault(k, []).append(what)
    print(f"  {what:58s} -> {k}", flush=True)


def rand(rng, shape, dt):
    dt = np.dtype(dt)
    if dt == np.uint8 or dt == np.bool_:
        return np.asfortranarray((rng.random(shape) < 0.4).astype(dt))
    if dt.kind == "i":
        return np.asfortranarray(rng.integers(-50, 50, size=shape).astype(dt))
    return np.asfortranarray((rng.random(shape) - 0.3).astype(dt))


def gather_case(what, r, offs, R, bc, red, *, pad="cond", flags=0, region=None, nsteps=0, **kw):
    """pad: 'cond' (parent == array) or 'out' (Halo ring of R cells on every axis, refreshed by sb200_update_halo)."""
    l = _lib()
    nd = r.ndim
    et = A.ELTYPE_OF_DTYPE[r.dtype]
    oet = orc.out_eltype(red, et)
    if pad == "cond":
        parent, off = r.copy(order="F"), (0,) * nd
    else:
        parent = np.full(tuple(s + 2 * R for s in r.shape), 3, dtype=r.dtype, order="F")
        parent[tuple(slice(R, R + s) for s in r.shape)] = r
        off = (R,) * nd
    bcs = bc if isinstance(bc, tuple) else (bc,) * nd
    h = build_desc(size=r.shape, eltype=et, out_eltype=oet, offsets=offs, radius=R, boundary=bcs, reducer=red,
                   src_off=off, src_ext=parent.shape, flags=flags, region=region, **kw)
    dst0 = np.full(r.shape, 5, dtype=A.DTYPE_OF_ELTYPE[oet], order="F")
    if nsteps:  # SwitchingStencilArray loop through sb200_iterate (Conditional padding)
        a, b = Dev(parent), Dev(dst0)
        A.check(l.sb200_iterate(h.ptr(), a.p, b.p, nsteps, None))
        A.check(l.sb200_stream_sync(None))
        note(red, what)
        got = (a if nsteps % 2 == 0 else b).get()
        want = orc.iterate(h, parent.copy(order="F"), dst0.copy(order="F"), nsteps)
        same(got, want, what)
        a.free(); b.free()
        return
    s, d = Dev(parent), Dev(dst0)
    p_cpu = parent.copy(order="F")
    gens = {0: 1, 1: 2, 2: 4, 4: 8}.get((flags >> 4) & 7, (flags >> 4) & 7)   # SB200_FLAG_GENS_OF
    if gens > 1:  # dest = f(f(src)) / f^4(src) on the output region, everything else of dest untouched
        A.check(l.sb200_gather(h.ptr(), s.p, d.p, None))
        A.check(l.sb200_stream_sync(None))
        note(red, what)
        h1 = build_desc(size=r.shape, eltype=et, out_eltype=oet, offsets=offs, radius=R, boundary=bcs, reducer=red, **kw)
        allg = orc.iterate(h1, parent.copy(order="F"), dst0.copy(order="F"), gens)
        want = dst0.copy(order="F")
        sl = tuple(slice(region[0][a], region[1][a]) for a in range(nd)) if region else tuple(slice(None) for _ in range(nd))
        want[sl] = allg[sl]
        same(d.get(), want, what)
        s.free(); d.free()
        return
    if pad != "cond":
        A.check(l.sb200_update_halo(h.ptr(), s.p, None))
        orc.update_halo(h, p_cpu)
        seen.setdefault(l.sb200_last_kernel().decode(), []).append(what)
    A.check(l.sb200_gather(h.ptr(), s.p, d.p, None))
    A.check(l.sb200_stream_sync(None))
    note(red, what)
    want = orc.gather(h, p_cpu, dst0.copy(order="F"))
    same(d.get(), want, what)
    same(s.get(), p_cpu, what + " (source ring)")
    s.free(); d.free()


def scatter_case(what, r, offs, R, bc, w, rule=A.SCATTER_CENTER_WEIGHTS, op=A.OP_ADD, flags=0):
    l = _lib()
    et = A.ELTYPE_OF_DTYPE[r.dtype]
    h = build_desc(size=r.shape, eltype=et, out_eltype=et, offsets=offs, radius=R, boundary=bc, weights=w,
                   scatter_op=op, scatter_rule=rule, flags=flags)
    rng = np.random.default_rng(3)
    dst0 = rand(rng, r.shape, r.dtype)
    s, d = Dev(r), Dev(dst0)
    A.check(l.sb200_scatter(h.ptr(), s.p, d.p, None))
    A.check(l.sb200_stream_sync(None))
    note("scatter", what)
    want = orc.scatter(h, r.copy(order="F"), dst0.copy(order="F"))
    same(d.get(), want, what)
    s.free(); d.free()


def plan_case(what, full_arr, offs, R, bcs, reducer, nslabs, ghost, pflags, nsteps, **kw):
    """A slab plan (sb200_plan_*: csrc/slab_plan.cu) with several slabs on device 0 against the oracle's single-domain iteration —
    puts the plan's own kernels (fused push / pull, signal, wait, end fills) and its peer copies under the sanitizer."""
    from stencils_b200.slab import SlabPlan
    et = A.ELTYPE_OF_DTYPE[full_arr.dtype]
    h = build_desc(size=full_arr.shape, eltype=et, out_eltype=et, offsets=offs, radius=R, boundary=bcs, reducer=reducer, **kw)
    want = orc.iterate(h, full_arr.copy(order="F"), np.zeros_like(full_arr, order="F"), nsteps)
    if DRY:
        seen.setdefault("(dry)", []).append(what)
        print(f"  {what:58s} -> (dry)")
        return
    rk = {k: v for k, v in kw.items() if k in ("born_mask", "survive_mask", "alpha")}
    plan = SlabPlan(full_arr.shape, offsets=offs, radius=R, reducer=reducer, boundary=bcs, eltype=et, ghost=ghost, devices=[0] * nslabs,
                    reducer_kwargs=rk, padval=kw.get("padval", 0), plan_flags=pflags)
    try:
        plan.load(full_arr)
        plan.iterate(nsteps)
        plan.sync()
        got = plan.store()
    finally:
        plan.close()
    seen.setdefault("slab plan", []).append(what)
    print(f"  {what:58s} -> slab plan ({_lib().sb200_last_kernel().decode()})")
    same(got, want, what)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dry", action="store_true", help="oracle side only (no GPU)")
    ap.add_argument("--quick", action="store_true", help="one case per kernel family (racecheck / synccheck are slow)")
    args = ap.parse_args()
    global DRY
    DRY = args.dry
    orc.lib()
    rng = np.random.default_rng(11)
    RE, WR, RF = A.REMOVE, A.WRAP, A.REFLECT
    moore = npr.offsets("Moore", 1, 2)
    full = not args.quick

    print("Life (Moore(1), UInt8)")
    g = rand(rng, (1024, 96), np.uint8)
    g16 = rand(rng, (1040, 70), np.uint8)   # 65 x 16 cells wide: not a multiple of 32, so two generations run the SWAR kernel
    gather_case("life 1024x96 wrap, one generation", g, moore, 1, WR, A.LIFE)
    gather_case("life 1040x70 wrap, two generations (SWAR)", g16, moore, 1, WR, A.LIFE, flags=A.FLAG_DOUBLE_STEP)
    gather_case("life 1024x96 wrap, two generations (bit-sliced)", g, moore, 1, WR, A.LIFE, flags=A.FLAG_DOUBLE_STEP)
    gather_case("life 1024x96 wrap, four generations (bit-sliced)", g, moore, 1, WR, A.LIFE, flags=A.FLAG_QUAD_STEP)
    gather_case("life 1024x96 wrap, eight generations (bit-sliced, one halo lane)", g, moore, 1, WR, A.LIFE, flags=A.FLAG_OCT_STEP)
    gather_case("life 1024x96 wrap, five generations (bit-sliced)", g, moore, 1, WR, A.LIFE, flags=A.flag_gens(5))
    if full:
        for n in (3, 6, 7):
            gather_case(f"life 1024x96 wrap, {n} generations (bit-sliced)", g, moore, 1, WR, A.LIFE, flags=A.flag_gens(n))
    gather_case("life 1024x96 wrap, iterate x21", g, moore, 1, WR, A.LIFE, nsteps=21)
    os.environ["SB200_LIFE_PACKED"] = "1"   # packed runs on a small grid: byte -> bits, bits -> bits, bits -> byte launches
    gather_case("life 1024x96 wrap, iterate x41 with the state packed between the launches", g, moore, 1, WR, A.LIFE, nsteps=41)
    gather_case("life 1024x96 wrap, iterate x58 with the state packed between the launches", g, moore, 1, WR, A.LIFE, nsteps=58)
    del os.environ["SB200_LIFE_PACKED"]
    gather_case("life 1024x96 wrap, no TMA", g, moore, 1, WR, A.LIFE, flags=A.FLAG_NO_TMA)
    if full:
        gather_case("life 8192x40 wrap/reflect, four generations", rand(rng, (8192, 40), np.uint8), moore, 1, (WR, RF), A.LIFE, flags=A.FLAG_QUAD_STEP)
        gather_case("life 4128x33 wrap, one generation (ragged strip)", rand(rng, (4128, 33), np.uint8), moore, 1, WR, A.LIFE)
        gather_case("life 250x77 remove (odd width)", rand(rng, (250, 77), np.uint8), moore, 1, RE, A.LIFE, padval=1)
        gather_case("life 512x64 reflect, B36/S23", rand(rng, (512, 64), np.uint8), moore, 1, RF, A.LIFE, born_mask=0b1001000, survive_mask=0b1100)
        gather_case("life 1024x96 region rows 8..80, two generations", g, moore, 1, WR, A.LIFE, flags=A.FLAG_DOUBLE_STEP, region=((0, 8, 0), (1024, 80, 0)))
        gather_case("life 1040x70 wrap, B36/S23 two generations (SWAR, table)", g16, moore, 1, WR, A.LIFE, flags=A.FLAG_DOUBLE_STEP, born_mask=0b1001000, survive_mask=0b1100)

    print("stream2d (named shapes, Float32 / Float64)")
    f64 = rand(rng, (200, 90), np.float64)
    f32 = rand(rng, (260, 70), np.float32)
    gather_case("Window(1) mean F64 remove", f64, npr.offsets("Window", 1, 2), 1, RE, A.MEAN, padval=0.5)
    if full:
        gather_case("Window(1) mean F64 wrap", f64, npr.offsets("Window", 1, 2), 1, WR, A.MEAN)
        gather_case("Window(1) sum F64 reflect", f64, npr.offsets("Window", 1, 2), 1, RF, A.SUM)
        gather_case("Window(1) mean F64 remove, Halo ring (cp.async producer)", f64, npr.offsets("Window", 1, 2), 1, RE, A.MEAN, pad="out", padval=0.5)
        gather_case("Window(1) mean F64 202x51 (unaligned rows)", rand(rng, (201, 51), np.float64), npr.offsets("Window", 1, 2), 1, WR, A.MEAN)
        w7 = rng.random(49).astype(np.float32)
        gather_case("Kernel(Window(3)) F32 remove", f32, npr.offsets("Window", 3, 2), 3, RE, A.KERNELDOT, weights=w7)
        gather_case("Kernel(Window(3)) F32 wrap", f32, npr.offsets("Window", 3, 2), 3, WR, A.KERNELDOT, weights=w7)
        gather_case("Circle(4) max F32 remove", f32, npr.offsets("Circle", 4, 2), 4, RE, A.MAX, padval=-1.0)
        gather_case("Circle(4) min F32 reflect", f32, npr.offsets("Circle", 4, 2), 4, RF, A.MIN)
        gather_case("VonNeumann(2) diffusion F32 wrap", f32, npr.offsets("VonNeumann", 2, 2), 2, WR, A.DIFFUSION, alpha=0.1)
        gather_case("Cross(2) sum F32 remove, 4100x19", rand(rng, (4100, 19), np.float32), npr.offsets("Cross", 2, 2), 2, RE, A.SUM)
    else:
        gather_case("Circle(4) max F32 remove", f32, npr.offsets("Circle", 4, 2), 4, RE, A.MAX, padval=-1.0)

    print("gather_stream (run-time tables)")
    pos = [(-1, 1), (-2, -1), (1, 0), (-2, 2)]
    gather_case("Positional 4 taps F32 wrap", f32, pos, 2, WR, A.SUM)
    if full:
        gather_case("Positional 4 taps F32 remove, 263x41 (cp.async)", rand(rng, (263, 41), np.float32), pos, 2, RE, A.MEAN, padval=2.0)
        gather_case("Window(4) sum Int32 reflect", rand(rng, (256, 60), np.int32), npr.offsets("Window", 4, 2), 4, RF, A.SUM)
        gather_case("Moore(3) max Int64 wrap", rand(rng, (128, 50), np.int64), npr.offsets("Moore", 3, 2), 3, WR, A.MAX)
        gather_case("Positional F64 remove, Halo ring", f64, pos, 2, RE, A.SUM, pad="out", padval=1.0)

    print("3-D")
    v = rand(rng, (64, 40, 24), np.float32)
    vn3 = npr.offsets("VonNeumann", 1, 3)
    gather_case("VonNeumann(1,3) diffusion F32 wrap", v, vn3, 1, WR, A.DIFFUSION, alpha=0.1)
    gather_case("VonNeumann(1,3) diffusion F32 wrap, two steps per launch", v, vn3, 1, WR, A.DIFFUSION, alpha=0.1, flags=A.FLAG_DOUBLE_STEP)
    gather_case("Window(1,3) mean F32 wrap (box3d)", v, npr.offsets("Window", 1, 3), 1, WR, A.MEAN)
    gather_case("Moore(1,3) max F32 remove/reflect/wrap (box3d)", v, npr.offsets("Moore", 1, 3), 1, (RE, RF, WR), A.MAX, padval=0.75)
    gather_case("Circle(2,3) sum F32 wrap (run-time table)", v, npr.offsets("Circle", 2, 3), 2, WR, A.SUM)
    if full:
        gather_case("VonNeumann(1,3) diffusion F32 remove/reflect/wrap", v, vn3, 1, (RE, RF, WR), A.DIFFUSION, alpha=0.1, padval=0.25)
        gather_case("VonNeumann(1,3) diffusion F64 300x17x12 two steps, region z 2..10", rand(rng, (300, 17, 12), np.float64), vn3, 1, (WR, WR, RE), A.DIFFUSION,
                    alpha=0.1, flags=A.FLAG_DOUBLE_STEP, region=((0, 0, 2), (300, 17, 10)))
        gather_case("VonNeumann(1,3) diffusion F32 iterate x5", v, vn3, 1, WR, A.DIFFUSION, alpha=0.1, nsteps=5)
        gather_case("VonNeumann(1,3) sum F64 reflect, region z 3..20", rand(rng, (36, 21, 24), np.float64), vn3, 1, RF, A.SUM, region=((0, 0, 3), (36, 21, 20)))
        gather_case("Moore(1,3) max F64 remove", rand(rng, (36, 21, 13), np.float64), npr.offsets("Moore", 1, 3), 1, RE, A.MAX, padval=-3.0)
        gather_case("VonNeumann(2,3) sum Int32 reflect", rand(rng, (40, 20, 12), np.int32), npr.offsets("VonNeumann", 2, 3), 2, RF, A.SUM)
        gather_case("Window(1,3) sum F32 37x22x19 (generic)", rand(rng, (37, 22, 19), np.float32), npr.offsets("Window", 1, 3), 1, WR, A.SUM)

    print("generic / 1-D")
    gather_case("1-D Window(2) mean F64 reflect", rand(rng, (1000,), np.float64), npr.offsets("Window", 2, 1), 2, RF, A.MEAN)
    if full:
        gather_case("Window(1) sum UInt8 wrap (generic)", rand(rng, (130, 50), np.uint8), npr.offsets("Window", 1, 2), 1, WR, A.SUM)
        gather_case("Window(1) mean F64 forced generic, Halo ring", f64, npr.offsets("Window", 1, 2), 1, WR, A.MEAN, pad="out", flags=A.FLAG_FORCE_GENERIC)

    print("scatter")
    w4 = np.array([0.5, 0.25, 2.0, 1.5], dtype=np.float32)
    sf = rand(rng, (256, 80), np.float32)
    scatter_case("Positional + F32 remove (stream)", sf, pos, 2, RE, w4)
    scatter_case("Positional + F32 remove (no TMA)", sf, pos, 2, RE, w4, flags=A.FLAG_NO_TMA)
    if full:
        scatter_case("Positional + F32 wrap", sf, pos, 2, WR, w4)
        scatter_case("Positional + F32 reflect, zero dest", sf, pos, 2, RF, w4, flags=A.FLAG_ZERO_DEST)
        scatter_case("Moore(1) max F64 remove, 203x45", rand(rng, (203, 45), np.float64), moore, 1, RE, rng.random(8), rule=A.SCATTER_WEIGHTS, op=A.OP_MAX)
        scatter_case("VonNeumann(2) + Int64 wrap", rand(rng, (128, 40), np.int64), npr.offsets("VonNeumann", 2, 2), 2, WR, rng.integers(-3, 4, 12))

    print("slab plans (several slabs on one device)")
    plan_case("life 1024x200, 2 slabs, flags, G=16, 40 generations", rand(rng, (1024, 200), np.uint8), moore, 1, (WR, WR), A.LIFE, 2, 16,
              A.PLAN_FLAGS_SYNC, 40)
    plan_case("diffusion 64x24x60, 3 slabs, events, overlap, 11 steps", rand(rng, (64, 24, 60), np.float32), npr.offsets("VonNeumann", 1, 3), 1,
              (WR, WR, WR), A.DIFFUSION, 3, 4, A.PLAN_OVERLAP_ON, 11, alpha=0.1)
    if full:
        plan_case("diffusion 64x20x50 remove/wrap/reflect, 2 slabs, flags, overlap", rand(rng, (64, 20, 50), np.float32), npr.offsets("VonNeumann", 1, 3),
                  1, (RE, WR, RF), A.DIFFUSION, 2, 2, A.PLAN_FLAGS_SYNC | A.PLAN_OVERLAP_ON, 7, alpha=0.1, padval=0.5)
        plan_case("life 512x150 wrap/remove, 3 slabs, events", rand(rng, (512, 150), np.uint8), moore, 1, (WR, RE), A.LIFE, 3, 4, 0, 9, padval=1)

    print("\nkernels exercised:")
    for k in sorted(seen):
        print(f"  {k}: {len(seen[k])} case(s)")
    missing = [w for w in WANT if not any(k.startswith(w) for k in seen)]
    if missing and full and not DRY:
        print("NOT exercised (pick other shapes for these):", missing)
    print("all cases match the oracle")
    return 0


if __name__ == "__main__":
    sys.exit(main())
