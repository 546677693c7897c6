"""GPU parity: the C ABI (libstencils_b200.so, sm_100a kernels) against the CPU oracle on the same seeded
inputs and the same sb200_desc. Bit-exact for every eltype and reducer (integer work, max/min and the float
folds all follow the reference's operation order, so the tolerance is 0 ulp; BASELINE.json allows 2)."""
import zlib

import numpy as np
import pytest

from oracle import np_restatement as npr
from stencils_b200 import _abi as A
from stencils_b200._desc import build_desc
from tests.util import bits_equal, dst_like, gpu_gather, gpu_scatter

pytestmark = pytest.mark.gpu

BC = {"remove": A.REMOVE, "wrap": A.WRAP, "reflect": A.REFLECT, "use": A.USE}
RED = {"sum": A.SUM, "mean": A.MEAN, "min": A.MIN, "max": A.MAX, "kerneldot": A.KERNELDOT, "life": A.LIFE,
       "diffusion": A.DIFFUSION}
DTYPES = [np.bool_, np.uint8, np.int32, np.int64, np.float32, np.float64]


def rand_array(rng, shape, dt):
    dt = np.dtype(dt)
    if dt == np.bool_:
        return np.asfortranarray(rng.random(shape) < 0.4)
    if dt.kind in "iu":
        return np.asfortranarray(rng.integers(0, 200 if dt == np.uint8 else 100000, size=shape).astype(dt))
    return np.asfortranarray((rng.random(shape) - 0.3).astype(dt))


def both(orc, r, offs, R, bc, pad, red, flags=0, switching=False, **kw):
    """Build parent + descriptor like StencilArray would, run oracle and GPU, compare dest and source ring."""
    r = np.asfortranarray(r)
    nd = r.ndim
    et = A.ELTYPE_OF_DTYPE[r.dtype]
    if pad == "cond":
        parent, size, off = r.copy(order="F"), r.shape, (0,) * nd
    elif pad == "out":
        parent = np.full(tuple(s + 2 * R for s in r.shape), 77, dtype=r.dtype, order="F")
        parent[tuple(slice(R, R + s) for s in r.shape)] = r
        size, off = r.shape, (R,) * nd
    else:
        parent, size, off = r.copy(order="F"), tuple(s - 2 * R for s in r.shape), (R,) * nd
    oet = orc.out_eltype(RED[red], et)
    dst_off = off if switching else (0,) * nd
    h = build_desc(size=size, eltype=et, out_eltype=oet, offsets=offs, radius=R, boundary=BC[bc], reducer=RED[red],
                   src_off=off, dst_off=dst_off, src_ext=parent.shape, flags=flags, **kw)
    halo = pad != "cond" and bc != "use"
    p_cpu = parent.copy(order="F")
    if halo:
        orc.update_halo(h, p_cpu)
    want = orc.gather(h, p_cpu, dst_like(h, 5))
    got, p_gpu = gpu_gather(h, parent, dst_like(h, 5), halo=halo)
    bits_equal(got, want)
    bits_equal(p_gpu, p_cpu)  # the ring refresh is a visible side effect
    return got


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("bc", ["remove", "wrap", "reflect"])
@pytest.mark.parametrize("pad", ["cond", "out", "in"])
def test_all_reducers_2d(orc, dt, bc, pad):
    rng = np.random.default_rng(zlib.crc32(repr((str(dt), bc, pad)).encode()))
    padval = {np.bool_: 1, np.uint8: 7}.get(dt, 3)
    for shape_, (shape, R) in zip([(67, 45), (130, 96), (33, 70), (64, 64), (51, 38)],
                                  [("Window", 1), ("Moore", 1), ("VonNeumann", 2), ("Circle", 3), ("Cross", 2)]):
        r = rand_array(rng, shape_, dt)
        offs = npr.offsets(shape, R, 2)
        reds = ["sum", "mean", "min", "max"] + (["life"] if len(offs) <= 31 else [])
        if np.dtype(dt).kind == "f":
            reds += ["kerneldot", "diffusion"]
        elif np.dtype(dt) in (np.int32, np.int64):
            reds += ["kerneldot"]
        for red in reds:
            w = rng.integers(-3, 4, size=len(offs)) if np.dtype(dt).kind != "f" else rng.random(len(offs))
            for flags in (0, A.FLAG_FORCE_GENERIC):
                both(orc, r, offs, R, bc, pad, red, flags=flags, padval=padval, weights=w, alpha=0.1,
                     born_mask=0b1001000, survive_mask=0b1100)


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32, np.uint8])
@pytest.mark.parametrize("nd", [1, 3])
def test_1d_3d(orc, dt, nd):
    rng = np.random.default_rng(nd)
    r = rand_array(rng, (1000,) if nd == 1 else (37, 22, 19), dt)
    for bc in ("remove", "wrap", "reflect"):
        for pad in ("cond", "out", "in"):
            for sh, R in [("Window", 1), ("VonNeumann", 1), ("Moore", 1), ("Window", 2)]:
                offs = npr.offsets(sh, R, nd)
                for red in ("sum", "mean", "max"):
                    both(orc, r, offs, R, bc, pad, red, padval=2)
                if np.dtype(dt).kind == "f":
                    both(orc, r, offs, R, bc, pad, "diffusion", alpha=0.05)
                    both(orc, r, offs, R, bc, pad, "diffusion", alpha=0.05, flags=A.FLAG_FORCE_GENERIC)


def test_use_boundary_and_switching_dest(orc):
    rng = np.random.default_rng(8)
    r = rand_array(rng, (40, 30), np.float64)
    offs = npr.offsets("Window", 2, 2)
    both(orc, r, offs, 2, "use", "in", "sum")
    both(orc, r, offs, 2, "use", "in", "mean", switching=True)
    both(orc, r, offs, 2, "wrap", "out", "mean", switching=True)
    both(orc, r, offs, 2, "reflect", "in", "max", switching=True)


def test_stencil_dims_below_array_dims(orc):
    rng = np.random.default_rng(9)
    r3 = rand_array(rng, (20, 17, 9), np.float32)
    for pad in ("cond", "out"):
        for bc in ("remove", "wrap"):
            both(orc, r3, npr.offsets("Window", 1, 2), 1, bc, pad, "sum")
            both(orc, r3, npr.offsets("Window", 1, 1), 1, bc, pad, "mean")
    r2 = rand_array(rng, (31, 29), np.int64)
    both(orc, r2, npr.offsets("Window", 1, 1), 1, "reflect", "cond", "sum")


def test_positional_named_rectangle_tables(orc):
    rng = np.random.default_rng(10)
    r = rand_array(rng, (45, 52), np.float32)
    tables = [
        ([(-1, 1), (-2, -1), (1, 0), (-2, 2)], 2),                       # README.md:102
        ([(-1, 0), (0, -1), (1, 0), (0, 1)], 1),                         # NamedStencil n,e,w,s (test/stencils.jl:193)
        ([(i, j) for j in range(-2, 2) for i in range(-1, 1)], 2),       # Rectangle((-1,0),(-2,1))
        ([(3, -3)], 3),
    ]
    for offs, R in tables:
        for bc in ("remove", "wrap", "reflect"):
            for red in ("sum", "max", "min", "mean"):
                both(orc, r, offs, R, bc, "cond", red, padval=-1.5)


def test_float_specials(orc):
    """NaN propagation and signed zeros of Julia's max/min (SURVEY §7 hard parts)."""
    rng = np.random.default_rng(12)
    for dt in (np.float32, np.float64):
        r = rand_array(rng, (48, 40), dt)
        r[rng.random(r.shape) < 0.05] = np.nan
        r[rng.random(r.shape) < 0.2] = 0.0
        r[rng.random(r.shape) < 0.2] = -0.0
        r[rng.random(r.shape) < 0.02] = np.inf
        r[rng.random(r.shape) < 0.02] = -np.inf
        for sh, R in (("Window", 1), ("Circle", 2), ("Circle", 4)):
            offs = npr.offsets(sh, R, 2)
            for red in ("max", "min", "sum", "mean"):
                for flags in (0, A.FLAG_FORCE_GENERIC):
                    both(orc, r, offs, R, "wrap", "cond", red, flags=flags)
                    both(orc, r, offs, R, "remove", "cond", red, flags=flags, padval=-0.0)


def test_region(orc):
    rng = np.random.default_rng(13)
    r = rand_array(rng, (70, 60), np.float64)
    offs = npr.offsets("Window", 1, 2)
    h = build_desc(size=r.shape, eltype=A.F64, out_eltype=A.F64, offsets=offs, radius=1, boundary=A.WRAP, reducer=A.MEAN,
                   region=((5, 10, 0), (64, 33, 0)))
    want = orc.gather(h, r, dst_like(h, -5.0))
    got, _ = gpu_gather(h, r, dst_like(h, -5.0))
    bits_equal(got, want)


@pytest.mark.parametrize("bc", ["remove", "wrap", "reflect"])
@pytest.mark.parametrize("op", ["add", "max", "min"])
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32])
def test_scatter(orc, bc, op, dt):
    rng = np.random.default_rng(14)
    cases = [([(-1, 1), (-2, -1), (1, 0), (-2, 2)], 2), (npr.offsets("Moore", 1, 2), 1), (npr.offsets("VonNeumann", 2, 2), 2)]
    for offs, R in cases:
        S = 2 * R + 1
        for (ny, nx) in [(40, 7 * S), (33, 6 * S + 1), (2 * R + 1, 2 * R + 2)]:
            if bc == "wrap" and nx % S:
                continue  # reference race: two columns of one pass hit the same dest column
            src = rand_array(rng, (ny, nx), dt)
            w = (rng.random(len(offs)) if np.dtype(dt).kind == "f" else rng.integers(1, 5, len(offs))).astype(dt)
            dest0 = rand_array(rng, (ny, nx), dt)
            for rule in (A.SCATTER_WEIGHTS, A.SCATTER_CENTER_WEIGHTS):
                for flags in (0, A.FLAG_ZERO_DEST, A.FLAG_FORCE_GENERIC):
                    et = A.ELTYPE_OF_DTYPE[np.dtype(dt)]
                    h = build_desc(size=(ny, nx), eltype=et, out_eltype=et, offsets=offs, radius=R, boundary=BC[bc],
                                   weights=w, scatter_op={"add": A.OP_ADD, "max": A.OP_MAX, "min": A.OP_MIN}[op],
                                   scatter_rule=rule, flags=flags)
                    want = orc.scatter(h, src, dest0.copy(order="F"))
                    bits_equal(gpu_scatter(h, src, dest0.copy(order="F")), want)


@pytest.mark.parametrize("bc", ["remove", "wrap", "reflect"])
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int64])
def test_scatter_stream_strips_and_runs(orc, bc, dt):
    """Several strips (ragged last one), several runs per strip, taps crossing strip edges: the TMA-fed kernel."""
    rng = np.random.default_rng(24)
    for offs, R, (ny, nx) in [([(-1, 1), (-2, -1), (1, 0), (-2, 2)], 2, (2600, 75)),
                              (npr.offsets("Window", 1, 2), 1, (1028, 300)),
                              (npr.offsets("Circle", 3, 2), 3, (1100, 63))]:
        src = rand_array(rng, (ny, nx), dt)
        w = (rng.random(len(offs)) if np.dtype(dt).kind == "f" else rng.integers(1, 5, len(offs))).astype(dt)
        dest0 = rand_array(rng, (ny, nx), dt)
        et = A.ELTYPE_OF_DTYPE[np.dtype(dt)]
        for op, rule, flags in [(A.OP_ADD, A.SCATTER_CENTER_WEIGHTS, 0), (A.OP_MAX, A.SCATTER_CENTER_WEIGHTS, A.FLAG_ZERO_DEST),
                                (A.OP_MIN, A.SCATTER_WEIGHTS, 0)]:
            h = build_desc(size=(ny, nx), eltype=et, out_eltype=et, offsets=offs, radius=R, boundary=BC[bc], weights=w,
                           scatter_op=op, scatter_rule=rule, flags=flags)
            want = orc.scatter(h, src, dest0.copy(order="F"))
            bits_equal(gpu_scatter(h, src, dest0.copy(order="F")), want)
            assert A.lib().sb200_last_kernel().decode() == "scatter_stream_kernel"


def test_iterate_life_and_diffusion(orc):
    from tests.util import stream, sync, to_dev, to_host
    rng = np.random.default_rng(15)
    l = A.lib()
    a = np.asfortranarray((rng.random((256, 192)) < 0.35).astype(np.uint8))
    h = build_desc(size=a.shape, eltype=A.U8, out_eltype=A.U8, offsets=npr.offsets("Moore", 1, 2), radius=1,
                   boundary=A.WRAP, reducer=A.LIFE)
    want = orc.iterate(h, a.copy(order="F"), np.zeros_like(a, order="F"), 25)
    ta, tb = to_dev(a), to_dev(np.zeros_like(a, order="F"))
    A.check(l.sb200_iterate(h.ptr(), ta.data_ptr(), tb.data_ptr(), 25, stream()))
    sync()
    bits_equal(to_host(tb, a.shape, a.dtype), want)
    # diffusion 3-D with a Halo ring refreshed every step
    g = rand_array(rng, (34, 30, 26), np.float32)
    offs = npr.offsets("VonNeumann", 1, 3)
    h = build_desc(size=(32, 28, 24), eltype=A.F32, out_eltype=A.F32, offsets=offs, radius=1, boundary=A.WRAP,
                   reducer=A.DIFFUSION, alpha=0.1, src_off=(1, 1, 1), dst_off=(1, 1, 1))
    b0 = g.copy(order="F")
    want = orc.iterate(h, g.copy(order="F"), b0.copy(order="F"), 6)
    ta, tb = to_dev(g), to_dev(b0)
    A.check(l.sb200_iterate(h.ptr(), ta.data_ptr(), tb.data_ptr(), 6, stream()))
    sync()
    bits_equal(to_host(ta, g.shape, g.dtype), want)


def test_errors_are_status_codes():
    l = A.lib()
    r = np.zeros((8, 8), order="F")
    offs = npr.offsets("Window", 1, 2)
    from tests.util import to_dev
    t = to_dev(r)
    h = build_desc(size=r.shape, eltype=A.F64, out_eltype=A.F64, offsets=offs, radius=1, boundary=A.USE, reducer=A.SUM)
    assert l.sb200_gather(h.ptr(), t.data_ptr(), to_dev(r).data_ptr(), None) == A.EUNSUPPORTED
    assert b"Use" in l.sb200_last_error()
    h = build_desc(size=r.shape, eltype=A.F64, out_eltype=A.F64, offsets=offs, radius=1, boundary=A.REMOVE, reducer=99)
    assert l.sb200_gather(h.ptr(), t.data_ptr(), to_dev(r).data_ptr(), None) == A.EUNSUPPORTED
    h = build_desc(size=r.shape, eltype=A.F64, out_eltype=A.F64, offsets=npr.offsets("Window", 8, 2), radius=8,
                   boundary=A.REMOVE, reducer=A.SUM)
    assert l.sb200_gather(h.ptr(), t.data_ptr(), to_dev(r).data_ptr(), None) == A.ESIZE
    assert l.sb200_gather(h.ptr(), t.data_ptr(), t.data_ptr(), None) != A.OK


@pytest.mark.parametrize("dt", [np.uint8, np.bool_])
@pytest.mark.parametrize("bc", ["remove", "wrap", "reflect"])
def test_life_swar_paths(orc, dt, bc):
    """The packed-byte Life kernel (csrc/life.cu) against the oracle: widths that are multiples of 16 with
    1..many 512-byte column groups, every boundary, Conway and a general born/survive table, non-0/1 bytes."""
    rng = np.random.default_rng(21)
    l = A.lib()
    moore = npr.offsets("Moore", 1, 2)
    for (W, H) in [(16, 5), (32, 33), (512, 70), (528, 64), (1040, 129), (4096, 40), (4112, 17), (8192 + 512, 50)]:
        g = (rng.random((W, H)) < 0.4)
        g = g.astype(dt) if dt == np.bool_ else (g * rng.integers(1, 255, size=(W, H))).astype(np.uint8)
        g = np.asfortranarray(g)
        for born, surv in ((1 << 3, 0b1100), (0b101001000, 0b100111110), (1, 0b111111111)):
            for pv in (0, 1):
                both(orc, g, moore, 1, bc, "cond", "life", born_mask=born, survive_mask=surv, padval=pv)
                if W >= 512 and H >= 16:
                    assert b"life_tma" in l.sb200_last_kernel()    # bulk-copy fed variant
                both(orc, g, moore, 1, bc, "cond", "life", born_mask=born, survive_mask=surv, padval=pv,
                     flags=A.FLAG_NO_TMA)
                assert b"life_swar" in l.sb200_last_kernel()       # register-prefetch variant


def test_life_swar_ghost_rows_and_regions(orc):
    """Slab layout: ghost rows on axis 1 (USE), wrap on axis 0, output restricted to a row region."""
    from tests.util import dst_like, gpu_gather
    rng = np.random.default_rng(22)
    moore = npr.offsets("Moore", 1, 2)
    W, H, G = 1024, 96, 4
    parent = np.asfortranarray((rng.random((W, H + 2 * G)) < 0.35).astype(np.uint8))
    for region in (None, ((0, 0, 0), (W, 7, 0)), ((0, 7, 0), (W, H - 5, 0)), ((0, H - 5, 0), (W, H, 0))):
        for flags in (0, A.FLAG_NO_TMA, A.FLAG_FORCE_GENERIC):
            h = build_desc(size=(W, H), eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1,
                           boundary=(A.WRAP, A.USE), reducer=A.LIFE, src_off=(0, G), dst_off=(0, G), src_ext=parent.shape,
                           dst_ext=parent.shape, region=region, flags=flags)
            want = orc.gather(h, parent, dst_like(h, 9))
            got, _ = gpu_gather(h, parent, dst_like(h, 9))
            bits_equal(got, want)


S2_SHAPES = [("Window", 1), ("Window", 2), ("Window", 3), ("Moore", 1), ("Moore", 2), ("VonNeumann", 1), ("VonNeumann", 2),
             ("Circle", 2), ("Circle", 3), ("Circle", 4), ("Cross", 1), ("Cross", 2), ("Diamond", 1), ("Diamond", 2)]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("bc", ["remove", "wrap", "reflect"])
def test_stream2d_kernels(orc, dt, bc):
    """The TMA-fed streaming kernels (csrc/stream2d.cuh): every compiled (shape, R) x reducer, 1..3 strips, ragged
    last strip, runs shorter and longer than a stage, against the oracle, bit-exact."""
    rng = np.random.default_rng(31)
    l = A.lib()
    es = np.dtype(dt).itemsize
    sizes = [(4096 // es + 64, 37), (64 // es * 4, 100), (2 * 4096 // es + 4096 // es // 2, 23)]
    for si, (shape, R) in enumerate(S2_SHAPES):
        offs = npr.offsets(shape, R, 2)
        W, H = sizes[si % len(sizes)]
        r = rand_array(rng, (W, H), dt)
        w = rng.random(len(offs))
        for red in ("sum", "mean", "min", "max", "kerneldot", "diffusion"):
            if red == "kerneldot" and len(offs) > 81:
                continue
            both(orc, r, offs, R, bc, "cond", red, padval=1.25, weights=w, alpha=0.07)
            assert b"stream2d" in l.sb200_last_kernel(), (shape, R, red, l.sb200_last_kernel())


def test_stream2d_ring_rows_regions_specials(orc):
    rng = np.random.default_rng(32)
    l = A.lib()
    W, H, G = 1536, 60, 4
    for dt in (np.float32, np.float64):
        parent = rand_array(rng, (W, H + 2 * G), dt)
        parent[rng.random(parent.shape) < 0.03] = np.nan
        parent[rng.random(parent.shape) < 0.1] = -0.0
        parent[rng.random(parent.shape) < 0.1] = 0.0
        et = A.ELTYPE_OF_DTYPE[np.dtype(dt)]
        for shape, R, red in (("VonNeumann", 1, A.DIFFUSION), ("Window", 1, A.MEAN), ("Circle", 4, A.MAX), ("Circle", 3, A.MIN),
                              ("Window", 3, A.KERNELDOT)):
            offs = npr.offsets(shape, R, 2)
            w = rng.random(len(offs))
            for region in (None, ((0, 0, 0), (W, 9, 0)), ((0, 9, 0), (W, H - 3, 0)), ((0, H - 3, 0), (W, H, 0))):
                for bc0 in (A.WRAP, A.REMOVE, A.REFLECT):
                    h = build_desc(size=(W, H), eltype=et, out_eltype=et, offsets=offs, radius=R, boundary=(bc0, A.USE),
                                   reducer=red, src_off=(0, G), dst_off=(0, G), src_ext=parent.shape, dst_ext=parent.shape,
                                   region=region, weights=w, alpha=0.1, padval=-2.5)
                    want = orc.gather(h, parent, dst_like(h, 9))
                    got, _ = gpu_gather(h, parent, dst_like(h, 9))
                    bits_equal(got, want)
                    assert b"stream2d" in l.sb200_last_kernel()


GS_TABLES = [
    ([(-1, 1), (-2, -1), (1, 0), (-2, 2)], 2),                                  # Positional, README.md:102
    ([(-1, 0), (0, -1), (1, 0), (0, 1)], 1),                                    # NamedStencil n,e,w,s (test/stencils.jl:193)
    ([(i, j) for j in range(-2, 2) for i in range(-1, 1)], 2),                  # Rectangle((-1,0),(-2,1))
    ([(3, -3)], 3),
    ("Annulus", 3, 1), ("Cardinal", 2, 0), ("Ordinal", 2, 0), ("AngledCross", 2, 0), ("BackSlash", 3, 0),
    ("ForwardSlash", 2, 0), ("Vertical", 3, 0), ("Horizontal", 4, 0), ("Window", 4, 0), ("Moore", 3, 0),
]


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("bc", ["remove", "wrap", "reflect"])
def test_gather_stream_any_table(orc, dt, bc):
    """The run-time-table streaming gather (csrc/gather_stream.cu): Positional / Named / Rectangle tables and the named
    shapes without a compile-time instantiation, 1..3 strips with a ragged last one, every reducer, bit-exact."""
    rng = np.random.default_rng(41)
    l = A.lib()
    es = np.dtype(dt).itemsize
    sizes = [(4096 // es + 64, 37), (64 // es * 4, 100), (2 * 4096 // es + 4096 // es // 2, 23)]
    isf = np.dtype(dt).kind == "f"
    for ti, tab in enumerate(GS_TABLES):
        if isinstance(tab[0], str):
            name, R, RI = tab
            offs = npr.offsets(name, R, 2, RI)
        else:
            offs, R = tab
        W, H = sizes[ti % len(sizes)]
        r = rand_array(rng, (W, H), dt)
        w = rng.random(len(offs)) if isf else rng.integers(1, 5, len(offs))
        for red in (("sum", "mean", "min", "max", "kerneldot", "diffusion") if isf else ("sum", "min", "max", "kerneldot")):
            both(orc, r, offs, R, bc, "cond", red, padval=1.25 if isf else 3, weights=w, alpha=0.07)
            assert l.sb200_last_kernel() == b"gather_stream_kernel", (tab, red, l.sb200_last_kernel())


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32])
@pytest.mark.parametrize("pad", ["out", "in", "cond-odd"])
def test_gather_stream_halo_padding_and_odd_widths(orc, dt, pad):
    """Halo padding (ring on every axis, parent index = logical + R) and rows that are not 16-byte multiples: the
    element-granular cp.async producer of csrc/gather_stream.cu. Named shapes that stream2d declines come here too."""
    rng = np.random.default_rng(43)
    l = A.lib()
    isf = np.dtype(dt).kind == "f"
    tables = [([(-1, 1), (-2, -1), (1, 0), (-2, 2)], 2), ("Window", 1, 0), ("Moore", 2, 0), ("Circle", 3, 0), ("Annulus", 4, 2),
              ("VonNeumann", 1, 0)]
    sizes = [(1030, 41), (2051, 37), (517, 64)] if pad == "cond-odd" else [(1040, 41), (2100, 37), (520, 64)]
    for ti, tab in enumerate(tables):
        if isinstance(tab[0], str):
            name, R, RI = tab
            offs = npr.offsets(name, R, 2, RI)
        else:
            offs, R = tab
        W, H = sizes[ti % len(sizes)]
        r = rand_array(rng, (W, H), dt)
        w = rng.random(len(offs)) if isf else rng.integers(1, 5, len(offs))
        for bc in ("remove", "wrap", "reflect"):
            for red in (("sum", "mean", "max", "kerneldot", "diffusion") if isf else ("sum", "min", "kerneldot")):
                both(orc, r, offs, R, bc, "cond" if pad == "cond-odd" else pad, red, padval=1.25 if isf else 3, weights=w, alpha=0.07,
                     switching=(pad == "out" and red == "sum"))
                # streaming either way: the compile-time kernels take the named shapes they instantiate (Halo rings and
                # unaligned source rows through their cp.async producer), the run-time-table kernel everything else
                assert l.sb200_last_kernel() in (b"gather_stream_kernel", b"stream2d_kernel"), (tab, bc, red, l.sb200_last_kernel())
                if isf and pad == "out" and isinstance(tab[0], str) and tab[0] in ("Window", "Moore", "Circle", "VonNeumann") \
                        and not (pad == "out" and red == "sum"):
                    assert l.sb200_last_kernel() == b"stream2d_kernel", (tab, bc, red, pad)


def test_gather_stream_ring_rows_regions_specials(orc):
    rng = np.random.default_rng(42)
    l = A.lib()
    W, H, G = 1536, 60, 4
    for dt in (np.float32, np.float64):
        parent = rand_array(rng, (W, H + 2 * G), dt)
        parent[rng.random(parent.shape) < 0.03] = np.nan
        parent[rng.random(parent.shape) < 0.1] = -0.0
        parent[rng.random(parent.shape) < 0.1] = 0.0
        et = A.ELTYPE_OF_DTYPE[np.dtype(dt)]
        for offs, R, red in (([(-1, 1), (-2, -1), (1, 0), (-2, 2)], 2, A.MAX), (npr.offsets("Annulus", 3, 2, 1), 3, A.MIN),
                             (npr.offsets("Cardinal", 2, 2), 2, A.DIFFUSION), (npr.offsets("Window", 4, 2), 4, A.KERNELDOT)):
            w = rng.random(len(offs))
            for region in (None, ((0, 0, 0), (W, 9, 0)), ((0, 9, 0), (W, H - 3, 0)), ((0, H - 3, 0), (W, H, 0))):
                for bc0 in (A.WRAP, A.REMOVE, A.REFLECT):
                    h = build_desc(size=(W, H), eltype=et, out_eltype=et, offsets=offs, radius=R, boundary=(bc0, A.USE),
                                   reducer=red, src_off=(0, G), dst_off=(0, G), src_ext=parent.shape, dst_ext=parent.shape,
                                   region=region, weights=w, alpha=0.1, padval=-2.5)
                    want = orc.gather(h, parent, dst_like(h, 9))
                    got, _ = gpu_gather(h, parent, dst_like(h, 9))
                    bits_equal(got, want)
                    assert l.sb200_last_kernel() == b"gather_stream_kernel"


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("bc", ["remove", "wrap", "reflect"])
def test_stream3d_vonneumann(orc, dt, bc):
    """The 2.5-D streaming kernel (csrc/stream3d.cu) for VonNeumann{1,3}: ragged tiles, several x tiles, short z."""
    rng = np.random.default_rng(41)
    l = A.lib()
    es = np.dtype(dt).itemsize
    offs = npr.offsets("VonNeumann", 1, 3)
    for (X, Y, Z) in [(1024 // es + 32 // es * 2, 21, 9), (32 // es, 40, 12), (2048 // es, 16, 3), (64, 5, 70)]:
        r = rand_array(rng, (X, Y, Z), dt)
        for red in ("diffusion", "sum", "mean", "max", "min"):
            both(orc, r, offs, 1, bc, "cond", red, padval=0.75, alpha=0.1)
            assert b"stream3d" in l.sb200_last_kernel(), (X, Y, Z, red, l.sb200_last_kernel())


def test_stream3d_ghost_planes_and_regions(orc):
    rng = np.random.default_rng(42)
    l = A.lib()
    X, Y, Z, G = 320, 37, 20, 3
    offs = npr.offsets("VonNeumann", 1, 3)
    parent = rand_array(rng, (X, Y, Z + 2 * G), np.float32)
    for region in (None, ((0, 0, 0), (X, Y, 2)), ((0, 0, 2), (X, Y, Z - 1)), ((0, 0, Z - 1), (X, Y, Z))):
        for bcs in ((A.WRAP, A.WRAP, A.USE), (A.REMOVE, A.REFLECT, A.USE)):
            h = build_desc(size=(X, Y, Z), eltype=A.F32, out_eltype=A.F32, offsets=offs, radius=1, boundary=bcs,
                           reducer=A.DIFFUSION, src_off=(0, 0, G), dst_off=(0, 0, G), src_ext=parent.shape, dst_ext=parent.shape,
                           region=region, alpha=0.1, padval=1.5)
            want = orc.gather(h, parent, dst_like(h, 9))
            got, _ = gpu_gather(h, parent, dst_like(h, 9))
            bits_equal(got, want)
            assert b"stream3d" in l.sb200_last_kernel()


G3_TABLES = [("Window", 1), ("Moore", 1), ("VonNeumann", 2), ("Cross", 2), ("Circle", 2), ("Window", 2),
             [(0, 0, -1), (1, -1, 0), (-2, 0, 2), (0, 2, 1), (1, 1, 1)]]


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32])
@pytest.mark.parametrize("bc", ["remove", "wrap", "reflect"])
def test_gather_stream3d_any_table(orc, dt, bc):
    """The run-time-table 3-D streaming gather (csrc/gather_stream3d.cu): named 3-D shapes other than VonNeumann(1,3)
    and a Positional table, several tiles in x and y with ragged last ones, z runs, every reducer, bit-exact."""
    rng = np.random.default_rng(61)
    l = A.lib()
    es = np.dtype(dt).itemsize
    isf = np.dtype(dt).kind == "f"
    sizes = [(512 // es + 32, 37, 21), (64 // es * 3, 20, 40), (2 * 512 // es, 16, 19)]
    for ti, tab in enumerate(G3_TABLES):
        if isinstance(tab, tuple):
            offs, R = npr.offsets(tab[0], tab[1], 3), tab[1]
        else:
            offs, R = tab, 2
        shape = sizes[ti % len(sizes)]
        r = rand_array(rng, shape, dt)
        w = rng.random(len(offs)) if isf else rng.integers(1, 5, len(offs))
        for red in (("sum", "mean", "max", "kerneldot", "diffusion") if isf else ("sum", "min", "kerneldot")):
            both(orc, r, offs, R, bc, "cond", red, padval=1.25 if isf else 3, weights=w, alpha=0.07)
            # Window(1,3) / Moore(1,3) float folds have compile-time kernels (csrc/box3d.cu); everything else is the table kernel
            box = isf and tab in (("Window", 1), ("Moore", 1)) and red in ("sum", "mean", "max", "min")
            want_kernel = (b"box3d_kernel<window>" if tab[0] == "Window" else b"box3d_kernel<moore>") if box else b"gather_stream3d_kernel"
            assert l.sb200_last_kernel() == want_kernel, (tab, red, l.sb200_last_kernel())


def test_gather_stream3d_ghost_planes_regions_specials(orc):
    rng = np.random.default_rng(62)
    l = A.lib()
    X, Y, Z, G = 192, 24, 30, 3
    parent = rand_array(rng, (X, Y, Z + 2 * G), np.float32)
    parent[rng.random(parent.shape) < 0.03] = np.nan
    parent[rng.random(parent.shape) < 0.1] = -0.0
    for offs, R, red in ((npr.offsets("Window", 1, 3), 1, A.MAX), (npr.offsets("Moore", 1, 3), 1, A.SUM),
                         (npr.offsets("VonNeumann", 2, 3), 2, A.DIFFUSION)):
        for region in (None, ((0, 0, 0), (X, Y, 7)), ((0, 0, 7), (X, Y, Z - 2)), ((0, 0, Z - 2), (X, Y, Z))):
            for bc0, bc1 in ((A.WRAP, A.REFLECT), (A.REMOVE, A.WRAP), (A.REFLECT, A.REMOVE)):
                h = build_desc(size=(X, Y, Z), eltype=A.F32, out_eltype=A.F32, offsets=offs, radius=R, boundary=(bc0, bc1, A.USE),
                               reducer=red, src_off=(0, 0, G), dst_off=(0, 0, G), src_ext=parent.shape, dst_ext=parent.shape,
                               region=region, alpha=0.1, padval=-2.5)
                want = orc.gather(h, parent, dst_like(h, 9))
                got, _ = gpu_gather(h, parent, dst_like(h, 9))
                bits_equal(got, want)
                assert l.sb200_last_kernel() == (b"gather_stream3d_kernel" if R == 2 else b"box3d_kernel<window>" if red == A.MAX else b"box3d_kernel<moore>")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("shape_name", ["Window", "Moore"])
def test_box3d_compile_time_box_stencils(orc, dt, shape_name):
    """csrc/box3d.cu: Window(1,3) / Moore(1,3) x sum / mean / minimum / maximum with the folds in registers (two running
    chains per cell for the sums, separable extrema): several tiles in x and y with ragged last ones, z runs, every boundary
    per axis, NaN / -0.0 / Inf cells, output regions along z, against the oracle bit for bit; and whole-grid equality with
    the one-thread-per-cell generic kernel on a grid that fills every CTA several times."""
    rng = np.random.default_rng(63)
    l = A.lib()
    es = np.dtype(dt).itemsize
    et = A.ELTYPE_OF_DTYPE[np.dtype(dt)]
    offs = npr.offsets(shape_name, 1, 3)
    name = b"box3d_kernel<window>" if shape_name == "Window" else b"box3d_kernel<moore>"
    for shape in [(1024 // es + 64 // es, 37, 21), (32 // es * 3, 20, 40), (2 * 1024 // es, 16, 9), (512 // es, 3, 5)]:
        r = rand_array(rng, shape, dt)
        r[rng.random(shape) < 0.02] = np.nan
        r[rng.random(shape) < 0.05] = -0.0
        r[rng.random(shape) < 0.01] = np.inf
        for bcs in [(A.WRAP, A.WRAP, A.WRAP), (A.REMOVE, A.REFLECT, A.WRAP), (A.REFLECT, A.REMOVE, A.REMOVE), (A.WRAP, A.WRAP, A.REFLECT)]:
            for red in (A.SUM, A.MEAN, A.MAX, A.MIN):
                h = build_desc(size=shape, eltype=et, out_eltype=et, offsets=offs, radius=1, boundary=bcs, reducer=red, padval=-1.5)
                want = orc.gather(h, r, dst_like(h))
                got, _ = gpu_gather(h, r, dst_like(h))
                assert l.sb200_last_kernel() == name, l.sb200_last_kernel()
                bits_equal(got, want)
        Z = shape[2]
        if Z >= 9:
            for region in (((0, 0, 0), shape[:2] + (4,)), ((0, 0, 3), shape[:2] + (Z - 2,)), ((0, 0, Z - 1), shape)):
                h = build_desc(size=shape, eltype=et, out_eltype=et, offsets=offs, radius=1, boundary=(A.WRAP, A.REFLECT, A.REMOVE),
                               reducer=A.SUM, padval=0.5, region=region)
                want = orc.gather(h, r, dst_like(h, 9))
                got, _ = gpu_gather(h, r, dst_like(h, 9))
                assert l.sb200_last_kernel() == name
                bits_equal(got, want)
    big = rand_array(rng, (1536 // es * 2, 150, 70), dt)
    for red in (A.MEAN, A.MAX):
        h = build_desc(size=big.shape, eltype=et, out_eltype=et, offsets=offs, radius=1, boundary=A.WRAP, reducer=red)
        got, _ = gpu_gather(h, big, dst_like(h))
        assert l.sb200_last_kernel() == name
        ref, _ = gpu_gather(build_desc(size=big.shape, eltype=et, out_eltype=et, offsets=offs, radius=1, boundary=A.WRAP, reducer=red,
                                       flags=A.FLAG_FORCE_GENERIC), big, dst_like(h))
        assert l.sb200_last_kernel() == b"gather_generic"
        bits_equal(got, ref)


@pytest.mark.parametrize("dt", [np.uint8, np.bool_])
def test_life_two_generations_per_launch(orc, dt):
    """SB200_FLAG_DOUBLE_STEP (csrc/life.cu: life_tma2_kernel): dest = step(step(src)) against two oracle sweeps, and
    sb200_iterate (which uses it by itself) against the oracle for step counts of every residue mod 4."""
    from tests.util import stream, sync, to_dev, to_host
    rng = np.random.default_rng(51)
    l = A.lib()
    moore = npr.offsets("Moore", 1, 2)
    for (W, H), bc1 in [((512, 40), A.WRAP), ((3840 + 512, 67), A.WRAP), ((4096, 33), A.REFLECT), ((8192 + 1024, 130), A.WRAP),
                        ((1040, 50), A.REFLECT)]:
        g = (rng.random((W, H)) < 0.4)
        g = g.astype(dt) if dt == np.bool_ else (g * rng.integers(1, 255, size=(W, H))).astype(np.uint8)
        g = np.asfortranarray(g)
        et = A.ELTYPE_OF_DTYPE[np.dtype(dt)]
        for born, surv in ((1 << 3, 0b1100), (0b101001000, 0b100111110)):
            kw = dict(size=(W, H), eltype=et, out_eltype=et, offsets=moore, radius=1, boundary=(A.WRAP, bc1), reducer=A.LIFE,
                      born_mask=born, survive_mask=surv)
            h1 = build_desc(**kw)
            mid = orc.gather(h1, g, dst_like(h1))
            want = orc.gather(h1, mid, dst_like(h1))
            h2 = build_desc(flags=A.FLAG_DOUBLE_STEP, **kw)
            got, _ = gpu_gather(h2, g, dst_like(h2))
            bits_equal(got, want)
            assert l.sb200_last_kernel().startswith((b"life_tma2_kernel", b"life_bit_kernel"))
            # a region that stays inside the parent (what the slab iterator asks for)
            hr = build_desc(flags=A.FLAG_DOUBLE_STEP, region=((0, 2, 0), (W, H - 2, 0)), **kw)
            got, _ = gpu_gather(hr, g, dst_like(hr, 7))
            want_r = dst_like(hr, 7)
            want_r[:, 2:H - 2] = want[:, 2:H - 2]
            bits_equal(got, want_r)
        # four generations per launch (bit-sliced kernel): B3/S23 and widths that are multiples of 32 only
        hc = build_desc(**dict(kw, born_mask=1 << 3, survive_mask=0b1100))
        h4 = build_desc(flags=A.FLAG_QUAD_STEP, **dict(kw, born_mask=1 << 3, survive_mask=0b1100))
        if W % 32 == 0:
            want4 = g
            for _ in range(4):
                want4 = orc.gather(hc, want4, dst_like(hc))
            got, _ = gpu_gather(h4, g, dst_like(h4))
            bits_equal(got, want4)
            assert l.sb200_last_kernel().startswith(b"life_bit_kernel<4")
            hr4 = build_desc(flags=A.FLAG_QUAD_STEP, region=((0, 4, 0), (W, H - 4, 0)), **dict(kw, born_mask=1 << 3, survive_mask=0b1100))
            got, _ = gpu_gather(hr4, g, dst_like(hr4, 7))
            want_r = dst_like(hr4, 7)
            want_r[:, 4:H - 4] = want4[:, 4:H - 4]
            bits_equal(got, want_r)
        else:
            t_ = to_dev(g)
            assert l.sb200_gather(h4.ptr(), t_.data_ptr(), to_dev(g).data_ptr(), None) == A.EUNSUPPORTED
        h1 = hc
        for n in (4, 5, 6, 7, 9, 16, 17, 18, 19, 23, 40):
            want = orc.iterate(h1, g.copy(order="F"), np.zeros_like(g, order="F"), n)
            ta, tb = to_dev(g), to_dev(np.zeros_like(g, order="F"))
            A.check(l.sb200_iterate(h1.ptr(), ta.data_ptr(), tb.data_ptr(), n, stream()))
            sync()
            bits_equal(to_host(ta if n % 2 == 0 else tb, g.shape, g.dtype), want)
    # not supported -> status code, no silent single step
    hbad = build_desc(size=(512, 40), eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1, boundary=A.REMOVE, reducer=A.LIFE,
                      flags=A.FLAG_DOUBLE_STEP)
    t = to_dev(np.zeros((512, 40), dtype=np.uint8, order="F"))
    assert l.sb200_gather(hbad.ptr(), t.data_ptr(), to_dev(np.zeros((512, 40), dtype=np.uint8, order="F")).data_ptr(), None) == A.EUNSUPPORTED


def test_iterate_small_grids_replay_a_cuda_graph(orc):
    """sb200_iterate on launch-bound grids captures 32 launches as a CUDA graph and replays it; results and the launch
    count are those of the plain loop (Life with two generations per launch, Life single, diffusion with a ring)."""
    import torch
    from tests.util import to_dev, to_host
    rng = np.random.default_rng(71)
    l = A.lib()
    st = torch.cuda.Stream()
    cases = []
    g = np.asfortranarray((rng.random((1024, 96)) < 0.4).astype(np.uint8))
    moore = npr.offsets("Moore", 1, 2)
    cases.append((g, build_desc(size=g.shape, eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1, boundary=A.WRAP, reducer=A.LIFE), 331))
    cases.append((g, build_desc(size=g.shape, eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1, boundary=A.REFLECT, reducer=A.LIFE), 150))
    f = rand_array(rng, (256, 64), np.float32)
    cases.append((f, build_desc(size=f.shape, eltype=A.F32, out_eltype=A.F32, offsets=npr.offsets("VonNeumann", 1, 2), radius=1,
                                boundary=A.WRAP, reducer=A.DIFFUSION, alpha=0.1), 200))
    for a0, h, n in cases:
        want = orc.iterate(h, a0.copy(order="F"), np.zeros_like(a0, order="F"), n)
        ta, tb = to_dev(a0), to_dev(np.zeros_like(a0, order="F"))
        l.sb200_launch_count(1)
        with torch.cuda.stream(st):
            A.check(l.sb200_iterate(h.ptr(), ta.data_ptr(), tb.data_ptr(), n, st.cuda_stream))
        st.synchronize()
        launches = l.sb200_launch_count(1)
        bits_equal(to_host(ta if n % 2 == 0 else tb, a0.shape, a0.dtype), want)
        assert n // 8 <= launches <= n   # up to eight Life generations per launch


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_diffusion_two_steps_per_launch(orc, dt, monkeypatch):
    """SB200_FLAG_DOUBLE_STEP with the Diffusion reducer (csrc/stream3d2.cu: stream3d2_kernel): dest = step(step(src))
    against two oracle sweeps bit for bit (tile edges, ragged tiles, z-runs, Wrap seams, interior regions), and
    sb200_iterate with SB200_DIFFUSION_DOUBLE_STEP=1 against the oracle for odd and even step counts."""
    from tests.util import stream, sync, to_dev, to_host
    rng = np.random.default_rng(91)
    l = A.lib()
    offs = npr.offsets("VonNeumann", 1, 3)
    es = np.dtype(dt).itemsize
    shapes = [(64, 20, 9), (160, 17, 8), (300, 16, 8), (1024 // es + 8, 33, 21), (512, 40, 70), (16, 4, 4)]
    for shape in shapes:
        g = rand_array(rng, shape, dt)
        et = A.ELTYPE_OF_DTYPE[np.dtype(dt)]
        kw = dict(size=shape, eltype=et, out_eltype=et, offsets=offs, radius=1, reducer=A.DIFFUSION, alpha=0.1)
        h1 = build_desc(boundary=A.WRAP, **kw)
        want = orc.gather(h1, orc.gather(h1, g, dst_like(h1)), dst_like(h1))
        h2 = build_desc(boundary=A.WRAP, flags=A.FLAG_DOUBLE_STEP, **kw)
        got, _ = gpu_gather(h2, g, dst_like(h2))
        assert l.sb200_last_kernel() == b"stream3d2_kernel"
        bits_equal(got, want)
        # an output region that stays two planes inside the parent (what the slab iterator asks for): the boundary rule
        # named for axis 2 is never exercised, everything outside the region keeps its old dest value
        Z = shape[2]
        if Z >= 8:
            for bc2, lo, hi in ((A.REMOVE, 2, Z - 2), (A.REFLECT, 3, Z - 3), (A.WRAP, 0, 3), (A.WRAP, Z - 3, Z)):
                hr = build_desc(boundary=(A.WRAP, A.WRAP, bc2), flags=A.FLAG_DOUBLE_STEP, region=((0, 0, lo), (shape[0], shape[1], hi)), **kw)
                got, _ = gpu_gather(hr, g, dst_like(hr, 7))
                assert l.sb200_last_kernel() == b"stream3d2_kernel"
                want_r = dst_like(hr, 7)
                want_r[:, :, lo:hi] = want[:, :, lo:hi]
                bits_equal(got, want_r)
    # NaN / Inf / signed zeros travel through both levels like through two sweeps
    g = rand_array(rng, (128, 18, 10), dt)
    g[rng.random(g.shape) < 0.02] = np.nan
    g[rng.random(g.shape) < 0.02] = np.inf
    g[rng.random(g.shape) < 0.1] = -0.0
    kw = dict(size=g.shape, eltype=A.ELTYPE_OF_DTYPE[np.dtype(dt)], out_eltype=A.ELTYPE_OF_DTYPE[np.dtype(dt)], offsets=offs, radius=1,
              reducer=A.DIFFUSION, alpha=0.25, boundary=A.WRAP)
    h1 = build_desc(**kw)
    with np.errstate(all="ignore"):
        want = orc.gather(h1, orc.gather(h1, g, dst_like(h1)), dst_like(h1))
    got, _ = gpu_gather(build_desc(flags=A.FLAG_DOUBLE_STEP, **kw), g, dst_like(h1))
    bits_equal(got, want)
    # sb200_iterate schedules the pairs by itself when asked to
    monkeypatch.setenv("SB200_DIFFUSION_DOUBLE_STEP", "1")
    g = rand_array(rng, (96, 30, 22), dt)
    kw["size"] = g.shape
    h1 = build_desc(**kw)
    for n in (4, 5, 6, 7, 11):
        want = orc.iterate(h1, g.copy(order="F"), np.zeros_like(g, order="F"), n)
        ta, tb = to_dev(g), to_dev(np.zeros_like(g, order="F"))
        l.sb200_launch_count(1)
        A.check(l.sb200_iterate(h1.ptr(), ta.data_ptr(), tb.data_ptr(), n, stream()))
        sync()
        assert l.sb200_launch_count(1) < n, "no pair of steps was fused"
        bits_equal(to_host(ta if n % 2 == 0 else tb, g.shape, g.dtype), want)
    monkeypatch.delenv("SB200_DIFFUSION_DOUBLE_STEP")
    # layouts the kernel does not take -> status code, never a silent single step
    for bad in (dict(boundary=A.REFLECT), dict(boundary=(A.WRAP, A.REFLECT, A.WRAP))):
        hb = build_desc(flags=A.FLAG_DOUBLE_STEP, **dict(kw, **bad))
        t = to_dev(g)
        assert l.sb200_gather(hb.ptr(), t.data_ptr(), to_dev(g).data_ptr(), None) == A.EUNSUPPORTED


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_diffusion_two_steps_per_launch_remove_axes(orc, dt):
    """Two diffusion steps per launch with Remove(padval) on any subset of the axes (PAD variant of stream3d2_kernel; default
    since r02a) — out-of-bounds cells read padval at BOTH time levels — against two oracle sweeps, bit for bit."""
    rng = np.random.default_rng(92)
    l = A.lib()
    offs = npr.offsets("VonNeumann", 1, 3)
    et = A.ELTYPE_OF_DTYPE[np.dtype(dt)]
    for shape in [(64, 20, 9), (160, 17, 8), (300, 30, 8), (512, 15, 12), (1024 // np.dtype(dt).itemsize + 8, 33, 21)]:
        g = rand_array(rng, shape, dt)
        for bcs in [(A.REMOVE,) * 3, (A.REMOVE, A.WRAP, A.REMOVE), (A.WRAP, A.REMOVE, A.WRAP), (A.REMOVE, A.REMOVE, A.WRAP)]:
            kw = dict(size=shape, eltype=et, out_eltype=et, offsets=offs, radius=1, reducer=A.DIFFUSION, alpha=0.1, boundary=bcs,
                      padval=0.25)
            h1 = build_desc(**kw)
            want = orc.gather(h1, orc.gather(h1, g, dst_like(h1)), dst_like(h1))
            got, _ = gpu_gather(build_desc(flags=A.FLAG_DOUBLE_STEP, **kw), g, dst_like(h1))
            assert l.sb200_last_kernel() == b"stream3d2_kernel"
            bits_equal(got, want)


def test_life_eight_generations_per_launch(orc, monkeypatch):
    """SB200_FLAG_OCT_STEP (one-halo-lane layout of life_bit_kernel, the default build since r02a): dest = step^8(src) against
    eight oracle sweeps, an interior region, and sb200_iterate for step counts of every residue mod 16."""
    from tests.util import stream, sync, to_dev, to_host
    rng = np.random.default_rng(53)
    l = A.lib()
    moore = npr.offsets("Moore", 1, 2)
    for (W, H), bc1 in [((1024, 96), A.WRAP), ((3840 + 512, 67), A.WRAP), ((4096, 40), A.REFLECT), ((8192 + 1024, 130), A.WRAP)]:
        g = np.asfortranarray(((rng.random((W, H)) < 0.4) * rng.integers(1, 255, size=(W, H))).astype(np.uint8))
        kw = dict(size=(W, H), eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1, boundary=(A.WRAP, bc1), reducer=A.LIFE)
        h1 = build_desc(**kw)
        want = g
        for _ in range(8):
            want = orc.gather(h1, want, dst_like(h1))
        got, _ = gpu_gather(build_desc(flags=A.FLAG_OCT_STEP, **kw), g, dst_like(h1))
        assert l.sb200_last_kernel().startswith(b"life_bit_kernel<8")
        bits_equal(got, want)
        hr = build_desc(flags=A.FLAG_OCT_STEP, region=((0, 8, 0), (W, H - 8, 0)), **kw)
        got, _ = gpu_gather(hr, g, dst_like(hr, 7))
        want_r = dst_like(hr, 7)
        want_r[:, 8:H - 8] = want[:, 8:H - 8]
        bits_equal(got, want_r)
        monkeypatch.setenv("SB200_OCT_STEP", "1")
        for n in (48, 49, 50, 55, 63, 64, 65, 79, 80, 97):
            want_n = orc.iterate(h1, g.copy(order="F"), np.zeros_like(g, order="F"), n)
            ta, tb = to_dev(g), to_dev(np.zeros_like(g, order="F"))
            A.check(l.sb200_iterate(h1.ptr(), ta.data_ptr(), tb.data_ptr(), n, stream()))
            sync()
            bits_equal(to_host(ta if n % 2 == 0 else tb, g.shape, g.dtype), want_n)
        monkeypatch.delenv("SB200_OCT_STEP")


@pytest.mark.parametrize("gens", [3, 5, 6, 7])
def test_life_any_generations_per_launch(orc, gens):
    """SB200_FLAG_GENS(n) for the sizes between the power-of-two flags (include/stencils_b200.h): dest = step^n(src) in one launch
    of life_bit_kernel<n> against n oracle sweeps — Wrap and Reflect on axis 1, an interior region, UInt8 cells that are not 0/1,
    Bool cells — and refusals: a rule other than B3/S23, a diffusion sweep."""
    rng = np.random.default_rng(530 + gens)
    l = A.lib()
    moore = npr.offsets("Moore", 1, 2)
    fl = A.flag_gens(gens)
    for (W, H), bc1, et in [((1024, 96), A.WRAP, np.uint8), ((3840 + 512, 67), A.WRAP, np.uint8), ((4096, 40), A.REFLECT, np.uint8),
                            ((2048, 50), A.WRAP, np.bool_)]:
        alive = rng.random((W, H)) < 0.4
        g = np.asfortranarray(alive if et is np.bool_ else (alive * rng.integers(1, 255, size=(W, H))).astype(np.uint8))
        e = A.ELTYPE_OF_DTYPE[np.dtype(et)]
        kw = dict(size=(W, H), eltype=e, out_eltype=e, offsets=moore, radius=1, boundary=(A.WRAP, bc1), reducer=A.LIFE)
        h1 = build_desc(**kw)
        want = g
        for _ in range(gens):
            want = orc.gather(h1, want, dst_like(h1))
        got, _ = gpu_gather(build_desc(flags=fl, **kw), g, dst_like(h1))
        assert l.sb200_last_kernel().startswith(b"life_bit_kernel<%d" % gens)
        bits_equal(got, want)
        hr = build_desc(flags=fl, region=((0, gens, 0), (W, H - gens, 0)), **kw)
        got, _ = gpu_gather(hr, g, dst_like(hr, 7))
        want_r = dst_like(hr, 7)
        want_r[:, gens:H - gens] = want[:, gens:H - gens]
        bits_equal(got, want_r)
    from tests.util import stream, to_dev
    W, H = 1024, 64
    g = np.asfortranarray((rng.random((W, H)) < 0.4).astype(np.uint8))
    ta, tb = to_dev(g), to_dev(np.zeros_like(g, order="F"))
    other = build_desc(size=(W, H), eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1, boundary=A.WRAP, reducer=A.LIFE,
                       born_mask=(1 << 3) | (1 << 6), survive_mask=0b1100, flags=fl)
    assert l.sb200_gather(other.ptr(), ta.data_ptr(), tb.data_ptr(), stream()) == A.EUNSUPPORTED
    v = np.asfortranarray(rng.random((64, 16, 12)).astype(np.float32))
    h3 = build_desc(size=v.shape, eltype=A.F32, out_eltype=A.F32, offsets=npr.offsets("VonNeumann", 1, 3), radius=1, boundary=A.WRAP,
                    reducer=A.DIFFUSION, alpha=0.1, flags=fl)
    fa, fb = to_dev(v), to_dev(np.zeros_like(v, order="F"))
    assert l.sb200_gather(h3.ptr(), fa.data_ptr(), fb.data_ptr(), stream()) == A.EUNSUPPORTED


@pytest.mark.parametrize("gens", [1, 2, 3, 5, 8])
def test_life_packed_state(orc, gens):
    """SB200_FLAG_SRC_BITS / SB200_FLAG_DST_BITS (include/stencils_b200.h): the Life state one bit per cell — row r of a packed parent
    starts at byte r * ext[0] / 8, cell c is bit c % 8 of byte c / 8 (np.packbits(..., bitorder="little") along axis 0). Byte -> packed,
    packed -> packed and packed -> byte launches of `gens` generations against the oracle's single sweeps on the byte grid: Wrap and
    Reflect on axis 1, an interior region, cells that are not 0 / 1 on the byte side; and the refusals."""
    import torch
    from tests.util import stream, sync, to_dev
    rng = np.random.default_rng(640 + gens)
    l = A.lib()
    moore = npr.offsets("Moore", 1, 2)
    fl = A.flag_gens(gens)

    def pack(a):      # (W, H) 0/1 -> (W / 8, H) bytes, the packed parent
        return np.asfortranarray(np.packbits(np.asfortranarray(a != 0), axis=0, bitorder="little"))

    def run(h, src_np, dst_np):
        ts, td = to_dev(src_np), to_dev(dst_np)
        A.check(l.sb200_gather(h.ptr(), ts.data_ptr(), td.data_ptr(), stream()))
        sync()
        return np.asfortranarray(td.cpu().numpy().T)

    for (W, H), bc1 in [((1024, 96), A.WRAP), ((3840 + 512, 67), A.WRAP), ((4096, 40), A.REFLECT), ((16384, 48), A.WRAP)]:
        g = np.asfortranarray(((rng.random((W, H)) < 0.4) * rng.integers(1, 255, size=(W, H))).astype(np.uint8))
        kw = dict(size=(W, H), eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1, boundary=(A.WRAP, bc1), reducer=A.LIFE)
        h1 = build_desc(**kw)
        want = g
        for _ in range(gens):
            want = orc.gather(h1, want, dst_like(h1))
        zeros_b = np.zeros((W // 8, H), dtype=np.uint8, order="F")
        # byte -> packed (two generations and more; a packed source also runs single generations)
        if gens >= 2:
            got = run(build_desc(flags=fl | A.FLAG_DST_BITS, **kw), g, zeros_b)
            assert l.sb200_last_kernel() == b"life_bit_kernel<%d,u8->bits>" % gens
            bits_equal(got, pack(want))
        # packed -> packed, packed -> byte
        got = run(build_desc(flags=fl | A.FLAG_SRC_BITS | A.FLAG_DST_BITS, **kw), pack(g), zeros_b)
        assert l.sb200_last_kernel() == b"life_bit_kernel<%d,bits->bits>" % gens
        bits_equal(got, pack(want))
        got = run(build_desc(flags=fl | A.FLAG_SRC_BITS, **kw), pack(g), dst_like(h1))
        assert l.sb200_last_kernel() == b"life_bit_kernel<%d,bits->u8>" % gens
        bits_equal(got, want)
        # an interior region (the sweeps of the slab plans): everything outside stays as it was
        reg = dict(region=((0, gens, 0), (W, H - gens, 0)))
        got = run(build_desc(flags=fl | A.FLAG_SRC_BITS | A.FLAG_DST_BITS, **reg, **kw), pack(g), np.full((W // 8, H), 0x5A, dtype=np.uint8, order="F"))
        want_r = np.full((W // 8, H), 0x5A, dtype=np.uint8, order="F")
        want_r[:, gens:H - gens] = pack(want)[:, gens:H - gens]
        bits_equal(got, want_r)
    # refusals: a width that is not a multiple of 128 cells, one generation, another reducer, another rule
    W, H = 1024 + 32, 64
    ta, tb = torch.zeros(W * H, dtype=torch.uint8, device="cuda"), torch.zeros(W * H, dtype=torch.uint8, device="cuda")
    kw = dict(size=(W, H), eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1, boundary=A.WRAP, reducer=A.LIFE)
    assert l.sb200_gather(build_desc(flags=fl | A.FLAG_SRC_BITS | A.FLAG_DST_BITS, **kw).ptr(), ta.data_ptr(), tb.data_ptr(), stream()) == A.EUNSUPPORTED
    kw = dict(kw, size=(1024, 64))
    assert l.sb200_gather(build_desc(flags=A.FLAG_DST_BITS, **kw).ptr(), ta.data_ptr(), tb.data_ptr(), stream()) == A.EUNSUPPORTED
    assert l.sb200_gather(build_desc(flags=fl | A.FLAG_SRC_BITS, **dict(kw, born_mask=(1 << 3) | (1 << 6))).ptr(), ta.data_ptr(), tb.data_ptr(),
                          stream()) == A.EUNSUPPORTED
    assert l.sb200_gather(build_desc(flags=fl | A.FLAG_SRC_BITS, **dict(kw, reducer=A.MAX)).ptr(), ta.data_ptr(), tb.data_ptr(),
                          stream()) == A.EUNSUPPORTED


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_small_grid_kernel(orc, dt, monkeypatch):
    """small2d_kernel (csrc/small2d.cu): radius-1 shapes on grids that fit the L2 — no ring, every thread reads its three rows
    directly. Window / Moore / VonNeumann x sum / mean / minimum / maximum x Remove(padval) / Wrap / Reflect x Conditional / Halo{:out}
    (ring read straight through, refreshed ring compared too) on sizes that are no multiples of the 16-byte groups, against the oracle
    AND against the streaming kernel on the same input, bit for bit; larger grids keep the streaming kernel. The kernel is an
    experiment that is OFF by default (it measured no faster than the streaming kernel at 1000 x 1000): SB200_SMALL2D_MAX_CELLS enables it."""
    rng = np.random.default_rng(23)
    l = A.lib()
    for (W, H) in ((1000, 37), (131, 64), (66, 5), (258, 130)):
        r = np.asfortranarray((rng.random((W, H)) - 0.3).astype(dt))
        r[rng.random((W, H)) < 0.01] = -0.0
        for name in ("Window", "Moore", "VonNeumann"):
            offs = npr.offsets(name, 1, 2)
            for bc in ("remove", "wrap", "reflect"):
                for pad in ("cond", "out"):
                    for red in ("sum", "mean", "min", "max"):
                        monkeypatch.setenv("SB200_SMALL2D_MAX_CELLS", "1500000")
                        got = both(orc, r, offs, 1, bc, pad, red, padval=1.25)
                        assert l.sb200_last_kernel() == b"small2d_kernel", (name, bc, pad, red, l.sb200_last_kernel())
                        monkeypatch.delenv("SB200_SMALL2D_MAX_CELLS")
                        ref = both(orc, r, offs, 1, bc, pad, red, padval=1.25)
                        assert l.sb200_last_kernel() != b"small2d_kernel"
                        if got is not None and ref is not None:
                            bits_equal(got, ref)
    monkeypatch.setenv("SB200_SMALL2D_MAX_CELLS", "1500000")
    big = np.asfortranarray(rng.random((2048, 1024)).astype(dt))
    both(orc, big, npr.offsets("Window", 1, 2), 1, "remove", "cond", "mean", padval=0.0)
    assert l.sb200_last_kernel() == b"stream2d_kernel"


def test_iterate_packed_runs(orc, monkeypatch):
    """sb200_iterate keeps Life runs of >= 12 generations packed between the first and the last launch (bytes -> bits ... bits ->
    bytes, the packed grids inside the two buffers themselves). SB200_LIFE_PACKED=1 forces the path for a small grid: step counts of
    every residue mod 8 and both parities land in the buffer the contract names, bit-identical to the oracle; UInt8 cells that
    are not 0 / 1 and Bool cells; Reflect on axis 1; a width the packed kernels refuse falls back to byte launches."""
    from tests.util import stream, sync, to_dev, to_host
    rng = np.random.default_rng(61)
    l = A.lib()
    moore = npr.offsets("Moore", 1, 2)
    monkeypatch.setenv("SB200_LIFE_PACKED", "1")
    for (W, H), bc1, et in [((1024, 80), A.WRAP, np.uint8), ((2048, 50), A.REFLECT, np.uint8), ((1280, 64), A.WRAP, np.bool_), ((1024 + 32, 64), A.WRAP, np.uint8)]:
        alive = rng.random((W, H)) < 0.4
        g = np.asfortranarray(alive if et is np.bool_ else (alive * rng.integers(1, 255, size=(W, H))).astype(np.uint8))
        e = A.ELTYPE_OF_DTYPE[np.dtype(et)]
        h1 = build_desc(size=(W, H), eltype=e, out_eltype=e, offsets=moore, radius=1, boundary=(A.WRAP, bc1), reducer=A.LIFE)
        want, done = g.copy(order="F"), 0
        for n in (12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 31, 32, 33, 47, 64, 101):
            while done < n:
                want = orc.gather(h1, want, dst_like(h1))
                done += 1
            ta, tb = to_dev(g), to_dev(np.full_like(g, 1, order="F"))
            A.check(l.sb200_iterate(h1.ptr(), ta.data_ptr(), tb.data_ptr(), n, stream()))
            sync()
            assert (b"bits->u8" in l.sb200_last_kernel()) == (W % 128 == 0), l.sb200_last_kernel()
            bits_equal(to_host(ta if n % 2 == 0 else tb, g.shape, g.dtype), want)


def test_iterate_every_step_count(orc):
    """sb200_iterate splits a run into launches of 1 .. 8 generations (split_steps in csrc/api.cu): every step count 0 .. 40 plus
    a few long ones lands in the buffer the contract names, bit-identical to the oracle's single steps, with the launch count
    sb200_debug_split_steps predicts."""
    import ctypes as C
    from tests.util import stream, sync, to_dev, to_host
    rng = np.random.default_rng(59)
    l = A.lib()
    moore = npr.offsets("Moore", 1, 2)
    W, H = 1024, 80
    g = np.asfortranarray(((rng.random((W, H)) < 0.4) * rng.integers(1, 255, size=(W, H))).astype(np.uint8))
    h1 = build_desc(size=(W, H), eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1, boundary=A.WRAP, reducer=A.LIFE)
    want = g.copy(order="F")
    done = 0
    for n in list(range(0, 41)) + [57, 100, 131]:
        while done < n:
            want = orc.gather(h1, want, dst_like(h1))
            done += 1
        ta, tb = to_dev(g), to_dev(np.zeros_like(g, order="F"))
        l.sb200_launch_count(1)
        A.check(l.sb200_iterate(h1.ptr(), ta.data_ptr(), tb.data_ptr(), n, stream()))
        sync()
        launches = l.sb200_launch_count(1)
        bits_equal(to_host(ta if n % 2 == 0 else tb, g.shape, g.dtype), want if n else g)
        out = (C.c_int32 * 9)()
        A.check(l.sb200_debug_split_steps(n, 0x1FC if n >= 4 else 0, 1, out))
        assert sum(k * out[k] for k in range(9)) == n and sum(out) % 2 == n % 2
        assert launches == sum(out), (n, launches, list(out))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_kernelproduct_allow_fma(orc, dt):
    """SB200_FLAG_ALLOW_FMA: acc = fma(v_k, w_k, acc) instead of the reference's separately rounded multiply and add
    (src/stencils/kernel.jl:37-43). Not bit-exact by construction, and NOT inside the 2-ulp tolerance of BASELINE.json's north
    star either (r02e: 4 ulp on a 3 x 3 Float32 kernel, two differently rounded 9-term chains), which is why it is opt-in and the
    default stays the bit-exact fold. Measured here against the oracle: (a) the BASELINE config's kind of data (uniform [0, 1)
    field, weights normalised to sum 1): a few ulp, asserted <= 16 and printed; (b) adversarial cancelling weights: the error is bounded by the usual
    dot-product bound L * eps * sum |v_k w_k| (ulps of a cancelled result are meaningless); (c) shapes without a contracted
    kernel ignore the flag and stay bit-exact; (d) without the flag the same kernel family is bit-exact."""
    from tests.util import max_ulp
    rng = np.random.default_rng(97)
    l = A.lib()
    et = A.ELTYPE_OF_DTYPE[np.dtype(dt)]
    eps = np.finfo(dt).eps
    es = np.dtype(dt).itemsize
    shape = (8192 // es + 256, 300)
    g = np.asfortranarray(rng.random(shape).astype(dt))
    worst = {}
    for R in (1, 2, 3):
        offs = npr.offsets("Window", R, 2)
        w = rng.random(len(offs)).astype(dt)
        w = (w / w.sum(dtype=dt)).astype(dt)
        kw = dict(size=shape, eltype=et, out_eltype=et, offsets=offs, radius=R, boundary=A.REMOVE, reducer=A.KERNELDOT, padval=0.0)
        want = orc.gather(build_desc(weights=w, **kw), g, dst_like(build_desc(weights=w, **kw)))
        got, _ = gpu_gather(build_desc(weights=w, flags=A.FLAG_ALLOW_FMA, **kw), g, dst_like(build_desc(weights=w, **kw)))
        assert l.sb200_last_kernel() == b"stream2d_kernel<fma>"
        worst[R] = max_ulp(got, want)
        assert worst[R] <= 16, worst
        exact, _ = gpu_gather(build_desc(weights=w, **kw), g, dst_like(build_desc(weights=w, **kw)))
        assert l.sb200_last_kernel() == b"stream2d_kernel"
        bits_equal(exact, want)
        # adversarial: alternating signs, sums cancel to ~0
        wa = (w * np.where(np.arange(len(offs)) % 2 == 0, 1, -1)).astype(dt)
        want_a = orc.gather(build_desc(weights=wa, **kw), g, dst_like(build_desc(weights=wa, **kw)))
        got_a, _ = gpu_gather(build_desc(weights=wa, flags=A.FLAG_ALLOW_FMA, **kw), g, dst_like(build_desc(weights=wa, **kw)))
        bound = len(offs) * eps * float(np.abs(wa).sum())   # |v| < 1
        assert float(np.abs(got_a.astype(np.float64) - want_a.astype(np.float64)).max()) <= bound
    print(f"kernelproduct with FMA, max ulp vs oracle on uniform data ({np.dtype(dt).name}): {worst}")
    # a shape without a contracted kernel: the flag is a permission, the result stays bit-exact
    offs = npr.offsets("Moore", 1, 2)
    w = rng.random(len(offs)).astype(dt)
    kw = dict(size=shape, eltype=et, out_eltype=et, offsets=offs, radius=1, boundary=A.WRAP, reducer=A.KERNELDOT, weights=w)
    want = orc.gather(build_desc(**kw), g, dst_like(build_desc(**kw)))
    got, _ = gpu_gather(build_desc(flags=A.FLAG_ALLOW_FMA, **kw), g, dst_like(build_desc(**kw)))
    bits_equal(got, want)


def test_multi_generation_flags_need_aligned_parents(orc):
    """the multi-generation kernels need 16-byte aligned parents. A *_STEP flag on a parent that starts one element off
    is SB200_EUNSUPPORTED (never a silent single sweep), and sb200_iterate on such buffers still returns the right state —
    one generation per launch."""
    import torch
    from tests.util import stream, sync
    rng = np.random.default_rng(77)
    l = A.lib()
    moore = npr.offsets("Moore", 1, 2)
    W, H = 1024, 64
    g = np.asfortranarray((rng.random((W, H)) < 0.4).astype(np.uint8))
    h1 = build_desc(size=(W, H), eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1, boundary=A.WRAP, reducer=A.LIFE)
    flat_a = torch.zeros(W * H + 16, dtype=torch.uint8, device="cuda")
    flat_b = torch.zeros(W * H + 16, dtype=torch.uint8, device="cuda")
    flat_a[1:1 + W * H] = torch.from_numpy(np.ascontiguousarray(g.T).reshape(-1)).cuda()
    pa, pb = flat_a.data_ptr() + 1, flat_b.data_ptr() + 1
    for fl in (A.FLAG_DOUBLE_STEP, A.FLAG_QUAD_STEP, A.FLAG_OCT_STEP):
        hf = build_desc(size=(W, H), eltype=A.U8, out_eltype=A.U8, offsets=moore, radius=1, boundary=A.WRAP, reducer=A.LIFE, flags=fl)
        assert l.sb200_gather(hf.ptr(), pa, pb, stream()) == A.EUNSUPPORTED
    for n in (5, 20):
        flat_a[1:1 + W * H] = torch.from_numpy(np.ascontiguousarray(g.T).reshape(-1)).cuda()
        l.sb200_launch_count(1)
        A.check(l.sb200_iterate(h1.ptr(), pa, pb, n, stream()))
        sync()
        assert l.sb200_launch_count(1) >= n      # one generation per launch (at least one kernel each)
        res = (flat_a if n % 2 == 0 else flat_b)[1:1 + W * H].cpu().numpy().reshape(H, W).T
        want = orc.iterate(h1, g.copy(order="F"), np.zeros_like(g, order="F"), n)
        bits_equal(np.asfortranarray(res), want)
    v = np.asfortranarray(rng.random((64, 16, 12)).astype(np.float32))
    offs3 = npr.offsets("VonNeumann", 1, 3)
    h3 = build_desc(size=v.shape, eltype=A.F32, out_eltype=A.F32, offsets=offs3, radius=1, boundary=A.WRAP, reducer=A.DIFFUSION, alpha=0.1,
                    flags=A.FLAG_DOUBLE_STEP)
    fa = torch.zeros(v.size + 8, dtype=torch.float32, device="cuda")
    fb = torch.zeros(v.size + 8, dtype=torch.float32, device="cuda")
    assert l.sb200_gather(h3.ptr(), fa.data_ptr() + 4, fb.data_ptr() + 4, stream()) == A.EUNSUPPORTED
