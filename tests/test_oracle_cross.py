"""Cross-checks the two independent CPU restatements (C oracle vs NumPy pad-and-shift) on seeded random
inputs over every boundary x padding x eltype x reducer, so the oracle is not only pinned on the reference's
tiny goldens. Bit-exact comparisons throughout (both follow the same operation order). CPU only."""
import numpy as np
import pytest

from oracle import np_restatement as npr
from stencils_b200 import _abi as A
from stencils_b200._desc import build_desc

BC = {"remove": A.REMOVE, "wrap": A.WRAP, "reflect": A.REFLECT, "use": A.USE}
RED = {"sum": A.SUM, "mean": A.MEAN, "min": A.MIN, "max": A.MAX, "kerneldot": A.KERNELDOT, "life": A.LIFE,
       "diffusion": A.DIFFUSION}
DTYPES = [np.bool_, np.uint8, np.int32, np.int64, np.float32, np.float64]


def rand_array(rng, shape, dt):
    dt = np.dtype(dt)
    if dt == np.bool_:
        return np.asfortranarray(rng.random(shape) < 0.4)
    if dt.kind in "iu":
        hi = 200 if dt == np.uint8 else 1000
        return np.asfortranarray(rng.integers(0, hi, size=shape).astype(dt))
    a = (rng.random(shape) - 0.3).astype(dt)
    return np.asfortranarray(a)


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape, (a.dtype, b.dtype, a.shape, b.shape)
    if a.dtype.kind == "f":
        assert np.array_equal(a.view(f"u{a.itemsize}"), b.view(f"u{a.itemsize}")) or \
            np.array_equal(a, b, equal_nan=True) and np.array_equal(np.signbit(a), np.signbit(b))
    else:
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("bc", ["remove", "wrap", "reflect"])
@pytest.mark.parametrize("pad", ["cond", "out", "in"])
def test_reducers_2d(orc, dt, bc, pad):
    import zlib
    rng = np.random.default_rng(zlib.crc32(repr((str(dt), bc, pad)).encode()))
    r = rand_array(rng, (13, 9), dt)
    padval = {np.bool_: 1, np.uint8: 7}.get(dt, 3)
    for shape, R in [("Window", 1), ("Moore", 2), ("VonNeumann", 2), ("Circle", 2), ("Cross", 1)]:
        offs = npr.offsets(shape, R, 2)
        reds = ["sum", "mean", "min", "max", "life"] if len(offs) <= 31 else ["sum", "mean", "min", "max"]
        if np.dtype(dt).kind == "f":
            reds += ["kerneldot", "diffusion"]
        elif np.dtype(dt) in (np.int32, np.int64):
            reds += ["kerneldot"]
        for red in reds:
            w = rng.integers(-3, 4, size=len(offs)) if np.dtype(dt).kind != "f" else rng.random(len(offs))
            got = orc.stencil_array_sweep(r, offs, R, BC[bc], pad, RED[red], padval=padval, weights=w, alpha=0.1,
                                          born_mask=0b1001000, survive_mask=0b1100)
            want = npr.gather(r, offs, R, bc, pad, red, padval=padval, weights=w, alpha=0.1,
                              born_mask=0b1001000, survive_mask=0b1100)
            eq(got, want)


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32])
@pytest.mark.parametrize("nd", [1, 3])
def test_reducers_1d_3d(orc, dt, nd):
    rng = np.random.default_rng(nd)
    shape = (17,) if nd == 1 else (7, 6, 5)
    r = rand_array(rng, shape, dt)
    for bc in ("remove", "wrap", "reflect"):
        for pad in ("cond", "out", "in"):
            for sh, R in [("Window", 1), ("VonNeumann", 1), ("Moore", 1)]:
                offs = npr.offsets(sh, R, nd)
                for red in ("sum", "mean", "max"):
                    got = orc.stencil_array_sweep(r, offs, R, BC[bc], pad, RED[red], padval=2)
                    eq(got, npr.gather(r, offs, R, bc, pad, red, padval=2))


def test_use_boundary_reads_ring(orc):
    rng = np.random.default_rng(5)
    r = rand_array(rng, (9, 8), np.float64)
    offs = npr.offsets("Window", 2, 2)
    eq(orc.stencil_array_sweep(r, offs, 2, A.USE, "in", A.SUM), npr.gather(r, offs, 2, "use", "in", "sum"))


def test_stencil_with_fewer_dims_than_array(orc):
    """test/array.jl:81-86, 109-118: 1-D Window on a 2-D array == Vertical line; 2-D Window on 3-D == Positional."""
    rng = np.random.default_rng(6)
    r = rand_array(rng, (11, 7), np.float64)
    a = orc.stencil_array_sweep(r, npr.offsets("Window", 1, 1), 1, A.REMOVE, "cond", A.SUM)
    b = orc.stencil_array_sweep(r, npr.offsets("Vertical", 1, 2), 1, A.REMOVE, "cond", A.SUM)
    eq(a, b)
    r3 = rand_array(rng, (6, 5, 4), np.float64)
    pos3 = [(-1, -1, 0), (0, -1, 0), (1, -1, 0), (-1, 0, 0), (0, 0, 0), (1, 0, 0), (-1, 1, 0), (0, 1, 0), (1, 1, 0)]
    eq(orc.stencil_array_sweep(r3, npr.offsets("Window", 1, 2), 1, A.REMOVE, "cond", A.SUM),
       orc.stencil_array_sweep(r3, pos3, 1, A.REMOVE, "cond", A.SUM))
    # Halo pads every array axis by R even though the stencil is 2-D (src/padding.jl:104-110)
    eq(orc.stencil_array_sweep(r3, npr.offsets("Window", 1, 2), 1, A.WRAP, "out", A.SUM),
       orc.stencil_array_sweep(r3, pos3, 1, A.WRAP, "cond", A.SUM))


def test_float_specials(orc):
    """Julia max/min: NaN-propagating and -0.0 < +0.0; sum keeps IEEE order."""
    r = np.asfortranarray(np.array([[0.0, -0.0, 1.0], [np.nan, -0.0, 0.0], [-1.0, np.inf, -0.0]], dtype=np.float32))
    offs = npr.offsets("Window", 1, 2)
    for red in ("max", "min", "sum", "mean"):
        eq(orc.stencil_array_sweep(r, offs, 1, A.WRAP, "cond", RED[red]), npr.gather(r, offs, 1, "wrap", "cond", red))
    z = np.asfortranarray(np.array([[-0.0, 0.0], [0.0, -0.0]], dtype=np.float64))
    h2 = [(0, 0), (1, 0)]
    mx = orc.stencil_array_sweep(z, h2, 1, A.WRAP, "cond", A.MAX)
    mn = orc.stencil_array_sweep(z, h2, 1, A.WRAP, "cond", A.MIN)
    assert not np.signbit(mx).any() and np.signbit(mn).all()


@pytest.mark.parametrize("bc", ["remove", "wrap", "reflect"])
@pytest.mark.parametrize("op", ["add", "max", "min"])
def test_scatter_pass_order_equals_sorted_fold(orc, bc, op):
    """The literal pass loops (src/scatterstencil.jl:56-71) equal the per-destination sorted fold
    (pass, column, row, k) — bit-exact in Float32 — which is the formulation the GPU kernel uses."""
    rng = np.random.default_rng(11)
    offs = [(-1, 1), (-2, -1), (1, 0), (-2, 2)]  # README Positional shape, README.md:102
    for (ny, nx) in [(10, 15), (7, 10), (12, 11)]:
        if bc == "wrap" and nx % 5:
            continue  # reference race
        src = rand_array(rng, (ny, nx), np.float32)
        w = rng.random(4).astype(np.float32)
        dest0 = rand_array(rng, (ny, nx), np.float32)
        for rule, rname in ((A.SCATTER_WEIGHTS, "weights"), (A.SCATTER_CENTER_WEIGHTS, "center_weights")):
            h = build_desc(size=(ny, nx), eltype=A.F32, out_eltype=A.F32, offsets=offs, radius=2, boundary=BC[bc], weights=w,
                           scatter_op={"add": A.OP_ADD, "max": A.OP_MAX, "min": A.OP_MIN}[op], scatter_rule=rule)
            got = orc.scatter(h, src, dest0.copy(order="F"))
            eq(got, npr.scatter(src, dest0, offs, 2, bc, op, rname, w))


def test_iterate_equals_repeated_gather(orc):
    rng = np.random.default_rng(3)
    a = np.asfortranarray((rng.random((20, 16)) < 0.35).astype(np.uint8))
    offs = npr.offsets("Moore", 1, 2)
    h = build_desc(size=a.shape, eltype=A.U8, out_eltype=A.U8, offsets=offs, radius=1, boundary=A.WRAP, reducer=A.LIFE)
    cur = a.copy(order="F")
    for _ in range(5):
        cur = orc.gather(h, cur)
    b = np.zeros_like(a, order="F")
    eq(orc.iterate(h, a.copy(order="F"), b, 5), cur)
    cur2 = a
    for _ in range(5):
        cur2 = npr.gather(cur2, offs, 1, "wrap", "cond", "life")
    eq(cur, cur2)


def test_region_restricts_output(orc):
    rng = np.random.default_rng(4)
    r = rand_array(rng, (12, 10), np.float64)
    offs = npr.offsets("Window", 1, 2)
    full = orc.stencil_array_sweep(r, offs, 1, A.WRAP, "cond", A.MEAN)
    h = build_desc(size=r.shape, eltype=A.F64, out_eltype=A.F64, offsets=offs, radius=1, boundary=A.WRAP, reducer=A.MEAN,
                   region=((2, 3, 0), (9, 7, 0)))
    part = orc.gather(h, r, np.full(r.shape, -5.0, order="F"))
    want = np.full(r.shape, -5.0)
    want[2:9, 3:7] = full[2:9, 3:7]
    eq(part, want)


def test_errors(orc):
    r = np.zeros((4, 4), order="F")
    with pytest.raises(orc.OracleError):  # Use + Conditional: no getneighbor method (src/array.jl:133-138)
        orc.stencil_array_sweep(r, npr.offsets("Window", 1, 2), 1, A.USE, "cond", A.SUM)
    with pytest.raises(orc.OracleError):  # radius larger than the axis (src/array.jl:451-453)
        orc.stencil_array_sweep(r, npr.offsets("Window", 4, 2), 4, A.REMOVE, "cond", A.SUM)
