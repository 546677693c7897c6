"""Pins the CPU oracle (and the independent NumPy restatement) to every golden vector the reference's
own tests hold for the hot path (test/array.jl, test/stencils.jl; lifted by tests/golden/extract_goldens.py).
CPU only."""
import numpy as np
import pytest

from oracle import np_restatement as npr
from stencils_b200 import _abi as A
from stencils_b200._desc import build_desc

SHAPE_ENUM = {n: i for i, n in enumerate(npr.SHAPE_NAMES)}
BC = {"remove": A.REMOVE, "wrap": A.WRAP, "reflect": A.REFLECT, "use": A.USE}
RED = {"sum": A.SUM, "mean": A.MEAN, "min": A.MIN, "max": A.MAX, "kerneldot": A.KERNELDOT, "life": A.LIFE,
       "diffusion": A.DIFFUSION}


def r2d():
    return np.asfortranarray(np.outer(np.arange(1.0, 6.0), np.arange(100.0, 106.0)))  # (1.0:5.0) * (100.0:105.0)'


def window(R, N=2):
    return npr.offsets("Window", R, N)


def test_shape_enum_matches_header():
    assert [SHAPE_ENUM[n] for n in ("Window", "Moore", "VonNeumann", "Circle", "Annulus", "Ordinal")] == \
        [A.WINDOW, A.MOORE, A.VONNEUMANN, A.CIRCLE, A.ANNULUS, A.ORDINAL]


def test_offset_goldens(orc, goldens):
    for name, g in goldens["offsets"].items():
        want = [tuple(t) for t in g["v"]]
        got = orc.offsets(SHAPE_ENUM[g["shape"]], g["R"], g["N"], g.get("RI", 0))
        assert got == want, (name, g["ref"])
        assert npr.offsets(g["shape"], g["R"], g["N"], g.get("RI", 0)) == want, name


@pytest.mark.parametrize("shape", npr.SHAPE_NAMES)
@pytest.mark.parametrize("N", [1, 2, 3])
@pytest.mark.parametrize("R", [1, 2, 3])
def test_offsets_two_restatements_agree(orc, shape, N, R):
    RI = R - 1
    assert orc.offsets(SHAPE_ENUM[shape], R, N, RI) == npr.offsets(shape, R, N, RI)


def test_circle4_has_69_offsets(orc):
    offs = orc.offsets(A.CIRCLE, 4, 2)
    assert len(offs) == 69  # row half-widths by |o2|: 0,1,2 -> 4; 3 -> 3; 4 -> 2
    for o2, hw in {0: 4, 1: 4, 2: 4, 3: 3, 4: 2}.items():
        assert max(o[0] for o in offs if abs(o[1]) == o2) == hw


def test_indices_goldens(orc, goldens):
    """indices(A, (1,1)) under Remove/Wrap/Reflect (test/array.jl:5-13): probe each neighbour with a
    one-offset stencil over an array whose values encode their own index."""
    moore = npr.offsets("Moore", 1, 2)
    ids = np.asfortranarray(np.fromfunction(lambda i, j: (i + 1) * 10 + (j + 1), (4, 4)))
    for key in ("remove_4x4_at_1_1", "wrap_4x4_at_1_1", "reflect_4x4_at_1_1"):
        g = goldens["indices"][key]
        for k, o in enumerate(moore):
            out = orc.stencil_array_sweep(ids, [o], 1, BC[g["boundary"]], "cond", A.SUM, padval=-1.0)
            i, j = g["v"][k]
            want = -1.0 if not (1 <= i <= 4 and 1 <= j <= 4) else i * 10 + j
            assert out[0, 0] == want, (key, k, g["ref"])
    # plain stencil indices (test/stencils.jl:29-30): offsets + centre
    g = goldens["indices"]["moore_at_1_1"]
    assert [[1 + o[0], 1 + o[1]] for o in moore] == g["v"]
    g = goldens["indices"]["kernel_window_at_2_2"]
    assert [[2 + o[0], 2 + o[1]] for o in window(1)] == g["v"]


def _check(got, g):
    want = np.array(g["v"])
    if g["approx"]:
        np.testing.assert_allclose(got, want, rtol=1e-8)  # Julia `≈`: rtol = sqrt(eps)
    else:
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("impl", ["oracle", "numpy"])
def test_mapstencil_goldens_2d(orc, goldens, impl):
    """test/array.jl:121-259: A (Conditional), B (Halo{:out}), C (Halo{:in}), SA/SB/SC switching twins."""
    M = goldens["mapstencil"]
    r = r2d()

    def sweep(bc, pad, red, switching=False):
        if impl == "oracle":
            return orc.stencil_array_sweep(r, window(1), 1, BC[bc], pad, RED[red], switching=switching)
        return npr.gather(r, window(1), 1, bc, pad, red)

    # Remove / Use
    A1, B1 = sweep("remove", "cond", "mean"), sweep("remove", "out", "mean")
    SA1, SB1 = sweep("remove", "cond", "mean", True), sweep("remove", "out", "mean", True)
    for x in (B1, SA1, SB1):
        np.testing.assert_array_equal(A1, x)  # `A1 == B1 == SA1 == SB1`
    _check(A1, M["remove_mean_2d"])
    C1, SC1 = sweep("use", "in", "mean"), sweep("use", "in", "mean", True)
    np.testing.assert_array_equal(C1, SC1)
    np.testing.assert_array_equal(A1[1:-1, 1:-1], C1)
    np.testing.assert_array_equal(C1, r[1:-1, 1:-1])  # test/array.jl:147
    A1, B1 = sweep("remove", "cond", "sum"), sweep("remove", "out", "sum")
    np.testing.assert_array_equal(A1, B1)
    C1 = sweep("use", "in", "sum")
    np.testing.assert_array_equal(A1[1:-1, 1:-1], C1)
    _check(C1, M["remove_sum_2d_interior"])
    # Wrap
    A1, B1, SB1 = sweep("wrap", "cond", "mean"), sweep("wrap", "out", "mean"), sweep("wrap", "out", "mean", True)
    np.testing.assert_array_equal(A1, B1)
    np.testing.assert_array_equal(A1, SB1)
    _check(A1, M["wrap_mean_2d"])
    np.testing.assert_array_equal(sweep("wrap", "in", "mean"), sweep("wrap", "in", "mean", True))
    # Reflect
    A1, B1 = sweep("reflect", "cond", "mean"), sweep("reflect", "out", "mean")
    np.testing.assert_array_equal(A1, B1)
    _check(A1, M["reflect_mean_2d"])
    C1 = sweep("reflect", "in", "mean")
    np.testing.assert_array_equal(A1[2:-2, 2:-2], C1[1:-1, 1:-1])  # test/array.jl:256
    np.testing.assert_array_equal(A1[1:-1, 1:-1], r[1:-1, 1:-1])   # test/array.jl:258
    assert not np.array_equal(A1[1:-1, 1:-1], C1)                  # test/array.jl:259


@pytest.mark.parametrize("impl", ["oracle", "numpy"])
def test_mapstencil_goldens_1d(orc, goldens, impl):
    M = goldens["mapstencil"]
    x = np.arange(1.0, 6.0)
    w1 = npr.offsets("Window", 1, 1)

    def sweep(bc, pad):
        if impl == "oracle":
            return orc.stencil_array_sweep(x, w1, 1, BC[bc], pad, A.MEAN)
        return npr.gather(x, w1, 1, bc, pad, "mean")

    _check(sweep("wrap", "cond"), M["wrap_mean_1d"])
    _check(sweep("wrap", "out"), M["wrap_mean_1d"])
    _check(sweep("wrap", "in"), M["wrap_mean_1d_halo_in"])
    _check(sweep("reflect", "cond"), M["reflect_mean_1d"])
    _check(sweep("reflect", "out"), M["reflect_mean_1d"])
    _check(sweep("reflect", "in"), M["reflect_mean_1d_halo_in"])


def test_halo_in_overwrites_ring_of_user_array(orc):
    """Halo{:in} + Wrap rewrites the outer ring of the user's own array each sweep
    (src/array.jl:202-239); that is why test/array.jl:182 expects [3,3,3]."""
    x = np.arange(1.0, 6.0)
    h = build_desc(size=(3,), eltype=A.F64, out_eltype=A.F64, offsets=[(-1,), (0,), (1,)], radius=1,
                   boundary=A.WRAP, src_off=(1,), src_ext=(5,), reducer=A.MEAN)
    orc.update_halo(h, x)
    np.testing.assert_array_equal(x, [4.0, 2.0, 3.0, 4.0, 2.0])


def test_fill_goldens(orc, goldens):
    """stencil(A, I) neighbour vectors (test/stencils.jl): read each neighbour through a one-offset sweep."""
    F = goldens["fills"]
    win = np.asfortranarray(np.array(F["win_5x5"]["v"], dtype=np.int64))
    init = np.asfortranarray(np.array(F["init_6x6"]["v"], dtype=np.int64))

    def neighbors(arr, offs, R, at):
        return [int(orc.stencil_array_sweep(arr, [o], R, A.REMOVE, "cond", A.SUM)[at[0] - 1, at[1] - 1]) for o in offs]

    g = F["positional_h1_at_3_3"]
    assert neighbors(win, g["offsets"], 2, (3, 3)) == g["neighbors"]
    assert int(orc.stencil_array_sweep(win, g["offsets"], 2, A.REMOVE, "cond", A.SUM)[2, 2]) == g["sum"]
    g = F["positional_h2_at_3_3"]
    assert neighbors(win, g["offsets"], 1, (3, 3)) == g["neighbors"]
    g = F["rectangle_h1_at_3_3"]  # offsets(Rectangle): CartesianIndices(map(splat(:), O)), first axis fastest
    (a0, a1), (b0, b1) = g["axis_ranges"]
    rect = [(i, j) for j in range(b0, b1 + 1) for i in range(a0, a1 + 1)]
    assert len(rect) == g["length"]
    assert neighbors(np.asfortranarray(np.array(g["A"], dtype=np.int64)), rect, g["radius"], (3, 3)) == g["neighbors"]
    g = F["vonneumann_init_at_2_2"]
    assert neighbors(init, orc.offsets(A.VONNEUMANN, 1, 2), 1, (2, 2)) == g["neighbors"]
    g = F["named_h1_at_3_3"]
    assert neighbors(win, g["offsets"], 1, (3, 3)) == g["neighbors"]


def test_named_map_goldens(orc, goldens):
    """Full-matrix mapstencil goldens over NamedStencil offsets (test/stencils.jl:205-239). The user
    closures `s.n + s.w + center(s)` etc. are sums over the named offsets."""
    win = np.asfortranarray(np.array(goldens["fills"]["win_5x5"]["v"], dtype=np.int64))
    N = goldens["named_maps"]
    for impl in ("oracle", "numpy"):
        def sweep(offs):
            if impl == "oracle":
                return orc.stencil_array_sweep(win, offs, 1, A.REMOVE, "cond", A.SUM)
            return npr.gather(win, [tuple(o) for o in offs], 1, "remove", "cond", "sum")
        g = N["n_plus_w_plus_center"]
        np.testing.assert_array_equal(sweep(g["offsets"]), np.array(g["golden_minus_input"]) + win)
        for key in ("cardinal_W_plus_S", "ordinal_NE_plus_NW"):
            np.testing.assert_array_equal(sweep(N[key]["offsets"]), np.array(N[key]["v"]), err_msg=key)


def test_kernelproduct_goldens(orc, goldens):
    K = goldens["kernelproduct"]
    g = K["window_1to9"]  # Window{1,2}(SVector(1:9), 5) with kernel reshape(1:9,3,3): column-major linear weights
    hood = np.asfortranarray(np.arange(1, 10, dtype=np.int64).reshape(3, 3, order="F"))
    out = orc.stencil_array_sweep(hood, window(1), 1, A.REMOVE, "cond", A.KERNELDOT,
                                  weights=np.arange(1, 10).reshape(3, 3, order="F"))
    assert out[1, 1] == g["v"] == 285
    g = K["moore_vals"]
    out = orc.stencil_array_sweep(hood, npr.offsets("Moore", 1, 2), 1, A.REMOVE, "cond", A.KERNELDOT, weights=g["kernel"])
    assert out[1, 1] == g["v"]
    g = K["positional_60"]
    out = orc.stencil_array_sweep(hood, g["offsets"], 1, A.REMOVE, "cond", A.KERNELDOT, weights=g["kernel"])
    assert out[1, 1] == g["v"] == 60
    outf = orc.stencil_array_sweep(hood.astype(np.float32), g["offsets"], 1, A.REMOVE, "cond", A.KERNELDOT,
                                   weights=g["kernel"])
    assert outf.dtype == np.float32 and outf[1, 1] == 60.0


def _scatter_case(orc, g, impl):
    ny, nx = g["size"]
    src = np.full((ny, nx), g.get("src_fill", 0.0), order="F")
    if g.get("src") == "i+j":
        src = np.asfortranarray(np.fromfunction(lambda i, j: i + j + 2.0, (ny, nx)))
    offs = npr.offsets(g["shape"], g["R"], 2)
    w = np.full(len(offs), g["w"])
    rule = A.SCATTER_WEIGHTS if g["rule"] == "weights" else A.SCATTER_CENTER_WEIGHTS
    op = {"add": A.OP_ADD, "max": A.OP_MAX, "min": A.OP_MIN}[g["op"]]
    dest = np.zeros((ny, nx), order="F") if not g.get("zero_dest") else np.full((ny, nx), 123.0, order="F")
    if impl == "oracle":
        h = build_desc(size=(ny, nx), eltype=A.F64, out_eltype=A.F64, offsets=offs, radius=g["R"], boundary=A.REMOVE,
                       weights=w, scatter_op=op, scatter_rule=rule, flags=A.FLAG_ZERO_DEST if g.get("zero_dest") else 0)
        return orc.scatter(h, src, dest)
    if g.get("zero_dest"):
        dest[:] = 0
    return npr.scatter(src, dest, offs, g["R"], "remove", g["op"], g["rule"], w)


@pytest.mark.parametrize("impl", ["oracle", "numpy"])
def test_scatter_goldens(orc, goldens, impl):
    for name, g in goldens["scatter"].items():
        dest = _scatter_case(orc, g, impl)
        for cell, want in g["cells"].items():
            i, j = (int(v) for v in cell.split(","))
            if g["approx"]:
                assert dest[i - 1, j - 1] == pytest.approx(want, rel=1e-8), (name, cell, g["ref"])
            else:
                assert dest[i - 1, j - 1] == want, (name, cell, g["ref"])


def test_out_eltype_table(orc):
    assert orc.out_eltype(A.SUM, A.BOOL) == A.I64          # reduce_first(+, ::Bool) = Int
    assert orc.out_eltype(A.SUM, A.U8) == A.U8
    assert orc.out_eltype(A.MEAN, A.I32) == A.F64
    assert orc.out_eltype(A.MEAN, A.F32) == A.F32
    assert orc.out_eltype(A.MAX, A.BOOL) == A.BOOL
    assert orc.out_eltype(A.LIFE, A.U8) == A.U8
    with pytest.raises(orc.OracleError):
        orc.out_eltype(A.DIFFUSION, A.I32)
