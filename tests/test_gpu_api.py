"""GPU tests through the public host API (the Python mirror of the Julia API), written to read like the
reference's own test/array.jl and test/stencils.jl, plus the host-buffer C-ABI entry points."""
import builtins
import operator

import numpy as np
import pytest

import stencils_b200 as sb
from oracle import np_restatement as npr
from stencils_b200 import _abi as A
from stencils_b200.synth import synth_np, synth_torch
from tests.util import bits_equal

pytestmark = pytest.mark.gpu


def dev(a):
    import torch
    a = np.asfortranarray(a)
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().permute(*reversed(range(a.ndim)))


def host(t):
    if hasattr(t, "cpu"):
        return t.cpu().numpy()
    return np.asarray(t)  # NumPy arrays and (Switching)StencilArrays (logical cells, copied to the host)


def r2d():
    return np.outer(np.arange(1.0, 6.0), np.arange(100.0, 106.0))  # (1.0:5.0) * (100.0:105.0)'


def _check(got, g):
    want = np.array(g["v"])
    if g["approx"]:
        np.testing.assert_allclose(host(got), want, rtol=1e-8)
    else:
        np.testing.assert_array_equal(host(got), want)


@pytest.mark.parametrize("where", ["device", "host"])
def test_mapstencil_remove_use(goldens, where):
    """test/array.jl:121-162"""
    put = dev if where == "device" else np.asfortranarray
    M = goldens["mapstencil"]
    r = r2d()
    W = sb.Window(1, 2)
    A_ = sb.StencilArray(put(r), W, padding=sb.Conditional(), boundary=sb.Remove(0.0))
    B = sb.StencilArray(put(r), W, padding=sb.Halo("out"), boundary=sb.Remove(0.0))
    C = sb.StencilArray(put(r.copy()), W, padding=sb.Halo("in"), boundary=sb.Use())
    SA = sb.SwitchingStencilArray(put(r.copy()), W, padding=sb.Conditional(), boundary=sb.Remove(0.0))
    SB = sb.SwitchingStencilArray(put(r.copy()), W, padding=sb.Halo("out"), boundary=sb.Remove(0.0))
    SC = sb.SwitchingStencilArray(put(r.copy()), W, padding=sb.Halo("in"), boundary=sb.Use())
    A1, B1, C1 = sb.mapstencil(sb.mean, A_), sb.mapstencil(sb.mean, B), sb.mapstencil(sb.mean, C)
    SA1, SB1, SC1 = sb.mapstencil_(sb.mean, SA), sb.mapstencil_(sb.mean, SB), sb.mapstencil_(sb.mean, SC)
    assert SA1 is not SA and SA1.source is SA.dest  # switch(A) swaps the buffers, src/array.jl:610-611
    for x in (B1, SA1, SB1):
        np.testing.assert_array_equal(host(A1), host(x))
    _check(A1, M["remove_mean_2d"])
    np.testing.assert_array_equal(host(C1), np.asarray(SC1))
    np.testing.assert_array_equal(host(A1)[1:-1, 1:-1], host(C1))
    np.testing.assert_array_equal(host(C1), r[1:-1, 1:-1])
    A1, B1, C1 = sb.mapstencil(sb.sum, A_), sb.mapstencil(sb.sum, B), sb.mapstencil(sb.sum, C)
    np.testing.assert_array_equal(host(A1), host(B1))
    _check(C1, M["remove_sum_2d_interior"])
    np.testing.assert_array_equal(host(A1)[1:-1, 1:-1], host(C1))


@pytest.mark.parametrize("bc,key", [(sb.Wrap, "wrap"), (sb.Reflect, "reflect")])
def test_mapstencil_wrap_reflect(goldens, bc, key):
    """test/array.jl:164-259"""
    M = goldens["mapstencil"]
    x = np.arange(1.0, 6.0)
    s1 = sb.Window(1, 1)
    A1 = sb.mapstencil(sb.mean, sb.StencilArray(dev(x), s1, padding=sb.Conditional(), boundary=bc()))
    B1 = sb.mapstencil(sb.mean, sb.StencilArray(dev(x), s1, padding=sb.Halo("out"), boundary=bc()))
    C1 = sb.mapstencil(sb.mean, sb.StencilArray(dev(x.copy()), s1, padding=sb.Halo("in"), boundary=bc()))
    SC1 = sb.mapstencil_(sb.mean, sb.SwitchingStencilArray(dev(x.copy()), s1, padding=sb.Halo("in"), boundary=bc()))
    np.testing.assert_array_equal(host(A1), host(B1))
    _check(A1, M[f"{key}_mean_1d"])
    _check(C1, M[f"{key}_mean_1d_halo_in"])
    np.testing.assert_array_equal(host(C1), np.asarray(SC1))
    r = r2d()
    s2 = sb.Window(1, 2)
    A1 = sb.mapstencil(sb.mean, sb.StencilArray(dev(r), s2, padding=sb.Conditional(), boundary=bc()))
    B1 = sb.mapstencil(sb.mean, sb.StencilArray(dev(r), s2, padding=sb.Halo("out"), boundary=bc()))
    SB1 = sb.mapstencil_(sb.mean, sb.SwitchingStencilArray(dev(r.copy()), s2, padding=sb.Halo("out"), boundary=bc()))
    np.testing.assert_array_equal(host(A1), host(B1))
    np.testing.assert_array_equal(host(A1), np.asarray(SB1))
    _check(A1, M[f"{key}_mean_2d"])
    # mapstencil(f, hood, A; kw...) form incl. the shrunken Halo{:in} dest (src/gatherstencil.jl:22-39)
    res_in = sb.mapstencil(sb.mean, s2, dev(r.copy()), boundary=bc(), padding=sb.Halo("in"))
    assert tuple(res_in.shape) == (3, 4)
    np.testing.assert_array_equal(host(sb.mapstencil(sb.mean, s2, dev(r), boundary=bc())), host(A1))


def test_lower_dim_stencils_and_named_maps(goldens):
    """test/array.jl:81-86,109-118; test/stencils.jl:205-239"""
    rng = np.random.default_rng(0)
    r = rng.random((100, 100))
    a = sb.mapstencil(sb.sum, sb.StencilArray(dev(r), sb.Vertical(1, 2), boundary=sb.Remove(0.0)))
    b = sb.mapstencil(sb.sum, sb.StencilArray(dev(r), sb.Window(1, 1), boundary=sb.Remove(0.0)))
    np.testing.assert_array_equal(host(a), host(b))
    r3 = rng.random((40, 30, 20))
    pos = sb.Positional((-1, -1, 0), (0, -1, 0), (1, -1, 0), (-1, 0, 0), (0, 0, 0), (1, 0, 0), (-1, 1, 0), (0, 1, 0), (1, 1, 0))
    np.testing.assert_array_equal(host(sb.mapstencil(sb.sum, sb.StencilArray(dev(r3), sb.Window(1, 2), boundary=sb.Remove(0.0)))),
                                  host(sb.mapstencil(sb.sum, sb.StencilArray(dev(r3), pos, boundary=sb.Remove(0.0)))))
    win = np.array(goldens["fills"]["win_5x5"]["v"], dtype=np.int64)
    N = goldens["named_maps"]
    g = N["n_plus_w_plus_center"]  # s.n + s.w + center(s)
    out = sb.mapstencil(sb.sum, sb.StencilArray(dev(win), sb.NamedStencil(n=(-1, 0), w=(1, 0), c=(0, 0))))
    np.testing.assert_array_equal(host(out), np.array(g["golden_minus_input"]) + win)
    ns = sb.NamedStencil(sb.Cardinal(1))  # s.W + s.S
    sel = sb.NamedStencil(W=ns.offsets()[ns.names.index("W")], S=ns.offsets()[ns.names.index("S")])
    np.testing.assert_array_equal(host(sb.mapstencil(sb.sum, sb.StencilArray(dev(win), sel))), np.array(N["cardinal_W_plus_S"]["v"]))
    ns = sb.NamedStencil(sb.Ordinal(1))  # s.NE + s.NW
    sel = sb.NamedStencil(NE=ns.offsets()[ns.names.index("NE")], NW=ns.offsets()[ns.names.index("NW")])
    np.testing.assert_array_equal(host(sb.mapstencil(sb.sum, sb.StencilArray(dev(win), sel))), np.array(N["ordinal_NE_plus_NW"]["v"]))


def test_kernelproduct(goldens):
    """test/stencils.jl:268-304 and the README sharpen example (README.md:205-228)."""
    hood = np.arange(1, 10, dtype=np.int64).reshape(3, 3, order="F")
    k = sb.Kernel(sb.Window(1, 2), np.arange(1, 10).reshape(3, 3, order="F"))
    out = sb.mapstencil(sb.kernelproduct, sb.StencilArray(dev(hood), k))
    assert host(out)[1, 1] == 285
    k = sb.Kernel(sb.Positional((0, -1), (-1, 0), (1, 0), (0, 1)), np.arange(1, 5))
    assert host(sb.mapstencil(sb.kernelproduct, sb.StencilArray(dev(hood), k)))[1, 1] == 60
    rng = np.random.default_rng(1)
    r = rng.random((200, 150))
    sharpen = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], dtype=np.float64)
    out = host(sb.mapstencil(sb.kernelproduct, sb.StencilArray(dev(r), sb.Kernel(sb.Window(1), sharpen))))
    want = npr.gather(r, npr.offsets("Window", 1, 2), 1, "remove", "cond", "kerneldot", weights=sharpen)
    bits_equal(np.asfortranarray(out), np.asfortranarray(want))


def test_scatterstencil(goldens):
    """test/array.jl:385-493"""
    src = np.ones((5, 5))
    d = sb.scatterstencil_(sb.ScatterWeights(0.1), operator.add, dev(np.zeros((5, 5))),
                           sb.StencilArray(dev(src), sb.Moore(1), boundary=sb.Remove(0.0)))
    d = host(d)
    assert d[2, 2] == pytest.approx(0.8) and d[0, 2] == pytest.approx(0.5) and d[2, 0] == pytest.approx(0.5)
    assert d[0, 0] == pytest.approx(0.3) and d[4, 4] == pytest.approx(0.3)
    src = np.fromfunction(lambda i, j: i + j + 2.0, (5, 5))
    d = host(sb.scatterstencil_(sb.ScatterCenterWeights(1.0), max, dev(np.zeros((5, 5))),
                                sb.StencilArray(dev(src), sb.Moore(1), boundary=sb.Remove(0.0))))
    assert d[2, 2] == 8.0 and d[0, 0] == 4.0
    d = host(sb.scatterstencil_(sb.ScatterWeights(0.25), operator.add, dev(np.zeros((5, 5))),
                                sb.StencilArray(dev(np.ones((5, 5))), sb.VonNeumann(1), boundary=sb.Remove(0.0))))
    assert d[2, 2] == pytest.approx(1.0) and d[0, 2] == pytest.approx(0.75) and d[0, 0] == pytest.approx(0.5)
    d = host(sb.scatterstencil_(sb.ScatterWeights(0.01), operator.add, dev(np.zeros((7, 7))),
                                sb.StencilArray(dev(np.ones((7, 7))), sb.Moore(2), boundary=sb.Remove(0.0))))
    assert d[3, 3] == pytest.approx(0.24)
    ssa = sb.SwitchingStencilArray(dev(np.ones((5, 5))), sb.Moore(1), boundary=sb.Remove(0.0))
    res = sb.scatterstencil_(sb.ScatterWeights(0.1), operator.add, ssa)
    assert res is not ssa and np.asarray(res)[2, 2] == pytest.approx(0.8)


@pytest.mark.parametrize("dt", [np.bool_, np.uint8, np.int32, np.int64, np.float32, np.float64])
def test_update_boundary_every_eltype_and_boundary(orc, dt):
    """update_boundary_(A) (src/array.jl:195-239) through the host mirror for every element type (Bool parents used to trip the
    dest-eltype check: ADVICE r1) x Remove / Wrap / Reflect, device and host parents, against the oracle's ring refresh; and
    scatterstencil_ refreshing the source ring first, as the reference does (src/scatterstencil.jl:39)."""
    from stencils_b200._desc import build_desc
    rng = np.random.default_rng(31)
    et = A.ELTYPE_OF_DTYPE[np.dtype(dt)]
    inner = (rng.random((40, 24)) < 0.5) if dt == np.bool_ else (rng.random((40, 24)) * 50).astype(dt)
    for bc, enum in ((sb.Remove(np.asarray(1, dtype=dt)[()]), A.REMOVE), (sb.Wrap(), A.WRAP), (sb.Reflect(), A.REFLECT)):
        for R in (1, 2):
            st = sb.Window(R)
            for to_dev in (True, False):
                sa = sb.StencilArray(dev(inner) if to_dev else np.asfortranarray(inner), st, boundary=bc, padding=sb.Halo("out"))
                par0 = np.asfortranarray(host(sa.parent)).copy(order="F")
                par0[:R, :] = 7 if dt != np.bool_ else True     # garbage in the ring: the refresh must overwrite all of it
                par0[:, -R:] = 3 if dt != np.bool_ else False
                if to_dev:
                    sa.parent.copy_(dev(par0))
                else:
                    sa.parent[...] = par0
                sb.update_boundary_(sa)
                h = build_desc(size=inner.shape, eltype=et, out_eltype=et, offsets=st.offsets(), radius=R, boundary=enum,
                               src_off=(R, R), dst_off=(R, R), padval=1, reducer=A.MIN)
                want = orc.update_halo(h, par0.copy(order="F"))
                bits_equal(np.asfortranarray(host(sa.parent)), want)
    if np.dtype(dt).kind == "f":   # scatter over a Halo-padded source whose ring holds garbage: refreshed before the sweep
        src = (rng.random((30, 20)) - 0.2).astype(dt)
        sa = sb.StencilArray(dev(src), sb.Moore(1), boundary=sb.Wrap(), padding=sb.Halo("out"))
        sa.parent[0, :] = 99.0
        d = sb.scatterstencil_(sb.ScatterCenterWeights(np.asarray(0.5, dtype=dt)[()]), operator.add, dev(np.zeros((30, 20), dtype=dt)), sa)
        ref = sb.scatterstencil_(sb.ScatterCenterWeights(np.asarray(0.5, dtype=dt)[()]), operator.add, dev(np.zeros((30, 20), dtype=dt)),
                                 sb.StencilArray(dev(src), sb.Moore(1), boundary=sb.Wrap()))
        bits_equal(host(d), host(ref))
        assert float(host(sa.parent)[0, 5]) != 99.0


def test_unsupported_function_raises_on_gpu_box_too():
    a = sb.StencilArray(dev(np.zeros((8, 8))), sb.Window(1))
    with pytest.raises(sb.ArgumentError, match="no fallback"):
        sb.mapstencil(lambda h: 0.0, a)


def test_synth_generators_agree():
    for dt in (np.uint8, np.float32, np.float64):
        a = synth_np((300, 70), dt, 0x5EED0002, lo=999)
        b = synth_torch((300, 70), dt, 0x5EED0002, "cuda", lo=999)
        bits_equal(np.asfortranarray(host(b)), a)


@pytest.mark.parametrize("bc", [A.REMOVE, A.WRAP, A.REFLECT])
def test_gather_host_chunked_pipeline(orc, bc):
    """sb200_gather_host (H2D -> sweep -> D2H pipelined over 16 chunks of the slowest axis) == oracle."""
    import torch
    from stencils_b200._desc import build_desc
    W, H = 4096, 2304  # 9 MiB of uint8: above the 8 MiB chunking threshold
    g = synth_np((W, H), np.uint8, 7)
    h = build_desc(size=(W, H), eltype=A.U8, out_eltype=A.U8, offsets=npr.offsets("Moore", 1, 2), radius=1,
                   boundary=bc, reducer=A.LIFE)
    want = orc.gather(h, g)
    src = torch.from_numpy(np.ascontiguousarray(g.T)).pin_memory()
    dst = torch.zeros_like(src).pin_memory()
    A.check(A.lib().sb200_gather_host(h.ptr(), src.data_ptr(), dst.data_ptr()))
    bits_equal(np.asfortranarray(dst.numpy().T), want)
    f = synth_np((1024, 1200), np.float64, 8)
    h = build_desc(size=f.shape, eltype=A.F64, out_eltype=A.F64, offsets=npr.offsets("Window", 1, 2), radius=1,
                   boundary=bc, reducer=A.MEAN)
    want = orc.gather(h, f)
    out = np.zeros_like(f, order="F")
    A.check(A.lib().sb200_gather_host(h.ptr(), f.ctypes.data, out.ctypes.data))  # pageable host memory works too
    bits_equal(out, want)


def test_iterate_host(orc):
    from stencils_b200._desc import build_desc
    g = synth_np((512, 300), np.uint8, 9)
    S = sb.SwitchingStencilArray(g.copy(order="F"), sb.Moore(1), boundary=sb.Wrap())
    S = sb.iterate_(sb.Life(), S, 7)
    h = build_desc(size=g.shape, eltype=A.U8, out_eltype=A.U8, offsets=npr.offsets("Moore", 1, 2), radius=1,
                   boundary=A.WRAP, reducer=A.LIFE)
    want = orc.iterate(h, g.copy(order="F"), np.zeros_like(g, order="F"), 7)
    bits_equal(np.asfortranarray(np.asarray(S)), want)


def test_slab_iterator_single_gpu_matches_iterate():
    """world = 1: the slab iterator (ghost planes, wide halo, end refresh) == sb200_iterate, bit for bit."""
    import torch
    from stencils_b200._desc import build_desc
    from stencils_b200.slab import SlabIterator
    dev = torch.device("cuda")
    for name, shape, dt, bcs, ghost, nsteps in [("life", (1024, 300), np.uint8, (A.WRAP, A.WRAP), 8, 19),
                                                ("life", (512, 200), np.uint8, (A.REMOVE, A.REFLECT), 2, 5),
                                                ("diffusion", (128, 60, 50), np.float32, (A.WRAP, A.REFLECT, A.WRAP), 2, 7),
                                                ("diffusion", (64, 40, 30), np.float32, (A.REFLECT, A.WRAP, A.REMOVE), 3, 8)]:
        full = synth_torch(shape, dt, 77, dev)
        tfull = full.permute(*reversed(range(len(shape)))).contiguous()
        if name == "life":
            st, red, kw, et = sb.Moore(1), A.LIFE, dict(born_mask=8, survive_mask=12), A.U8
        else:
            st, red, kw, et = sb.VonNeumann(1, 3), A.DIFFUSION, dict(alpha=0.1), A.F32
        it = SlabIterator(tfull.clone(), offsets=st.offsets(), radius=1, reducer=red, boundary=bcs, eltype=et, ghost=ghost,
                          reducer_kwargs=kw, padval=0)
        it.step(nsteps)
        h = build_desc(size=shape, eltype=et, out_eltype=et, offsets=st.offsets(), radius=1, boundary=bcs, reducer=red,
                       padval=0, **kw)
        a, b = tfull.clone(), torch.empty_like(tfull)
        A.check(A.lib().sb200_iterate(h.ptr(), a.data_ptr(), b.data_ptr(), nsteps, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        ref = a if nsteps % 2 == 0 else b
        assert torch.equal(it.state.view(torch.uint8), ref.view(torch.uint8)), (name, shape, bcs)


def test_multi_stencilarray_mapstencil(orc):
    """test/array.jl:312-383 (multi-StencilArray mapstencil), with the user functions written as LinearCombination."""
    from stencils_b200._desc import build_desc
    Am = np.array([[1., 2, 3], [4, 5, 6], [7, 8, 9]])
    Bm = np.array([[10., 20, 30], [40, 50, 60], [70, 80, 90]])
    sa_A, sa_B = sb.StencilArray(dev(Am), sb.Moore(1)), sb.StencilArray(dev(Bm), sb.Moore(1))
    # "two StencilArrays": center(hood_a) + center(hood_b) == A .+ B
    res = sb.mapstencil(sb.LinearCombination(sb.center, sb.center), sa_A, sa_B)
    np.testing.assert_array_equal(host(res), Am + Bm)
    # sum(neighbors(hood_a)) + sum(neighbors(hood_b))
    res2 = sb.mapstencil(sb.LinearCombination(sb.sum, sb.sum), sa_A, sa_B)
    assert host(res2).shape == Am.shape
    offs = npr.offsets("Moore", 1, 2)

    def nsum(x, bc=A.REMOVE):
        return orc.stencil_array_sweep(np.asfortranarray(x), offs, 1, bc, "cond", A.SUM, padval=0.0)
    bits_equal(host(res2), nsum(Am) + nsum(Bm))
    # "SwitchingStencilArray with StencilArray": c + 0.1 * sum(neighbors(hood_r)), three sweeps
    mutable_arr = np.array([[0., 0, 0], [0, 1, 0], [0, 0, 0]])
    readonly_arr = np.array([[1., 1, 1], [1, 0, 1], [1, 1, 1]])
    ssa = sb.SwitchingStencilArray(dev(mutable_arr), sb.Moore(1))
    sa = sb.StencilArray(dev(readonly_arr), sb.Moore(1))
    f = sb.LinearCombination(sb.center, (0.1, sb.sum))
    want = mutable_arr.copy()
    for _ in range(3):
        ssa = sb.mapstencil_(f, ssa, sa)
        want = want + 0.1 * nsum(readonly_arr)
    result = np.asarray(ssa)
    assert result.shape == mutable_arr.shape and result[1, 1] > mutable_arr[1, 1]
    bits_equal(result, want)
    # "three StencilArrays"
    sas = [sb.StencilArray(dev(k * np.ones((5, 5))), sb.Moore(1)) for k in (1.0, 2.0, 3.0)]
    res3 = sb.mapstencil(sb.LinearCombination(sb.center, sb.center, sb.center), *sas)
    assert np.all(host(res3)[1:4, 1:4] == 6.0)
    # "mixed StencilArray and regular array": center(hood_a) + b_val == A .+ B
    res4 = sb.mapstencil(sb.LinearCombination(sb.center, sb.center), sa_A, dev(Bm))
    np.testing.assert_array_equal(host(res4), Am + Bm)
    with pytest.raises(sb.ArgumentError):
        sb.mapstencil(sb.LinearCombination(sb.center, sb.sum), sa_A, dev(Bm))  # plain arrays are indexed, not stencilled
    # random fields, Float32, different stencils / boundaries / paddings per argument, coefficients on every term
    rng = np.random.default_rng(5)
    X, Y, Z = (np.asfortranarray(rng.random((200, 96)).astype(np.float32) - 0.4) for _ in range(3))
    w = rng.random((3, 3)).astype(np.float32)
    sx = sb.StencilArray(dev(X), sb.Window(1), boundary=sb.Wrap(), padding=sb.Halo("out"))
    sy = sb.StencilArray(dev(Y), sb.Kernel(sb.Window(1), w), boundary=sb.Reflect())
    sz = sb.StencilArray(dev(Z), sb.Positional((-1, 1), (-2, -1), (1, 0), (-2, 2)), boundary=sb.Remove(np.float32(0.5)))
    got = sb.mapstencil(sb.LinearCombination((0.25, sb.mean), sb.kernelproduct, (-1.5, sb.maximum)), sx, sy, sz)
    t1 = orc.stencil_array_sweep(X, npr.offsets("Window", 1, 2), 1, A.WRAP, "cond", A.MEAN)
    t2 = orc.stencil_array_sweep(Y, npr.offsets("Window", 1, 2), 1, A.REFLECT, "cond", A.KERNELDOT, weights=w)
    t3 = orc.stencil_array_sweep(Z, [(-1, 1), (-2, -1), (1, 0), (-2, 2)], 2, A.REMOVE, "cond", A.MAX, padval=0.5)
    want = (np.float32(0.25) * t1 + t2) + np.float32(-1.5) * t3
    assert want.dtype == np.float32
    bits_equal(host(got), want)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_multi_array_single_pass_kernel(orc, dt, monkeypatch):
    """multi_tile2d_kernel (csrc/multi_tile.cu): the arguments of a multi-array gather in ONE pass — every source read once, the
    dest written once (opt-in, SB200_MULTI_SINGLE_PASS=1) — against the oracle's per-argument sweeps combined left to right, and against
    the default sweep-per-argument path, bit for bit: sizes that are no multiples of the 128 x 8 tile, a radius-4 shape, Wrap +
    Halo{:out}, Reflect, Remove with a non-zero padval, Halo{:in}, coefficients, a plain-array argument; a 3-D combination keeps
    the old path."""
    rng = np.random.default_rng(17)
    l = A.lib()
    shape = (333, 77)
    X, Y, Z, P = (np.asfortranarray((rng.random(shape) - 0.4).astype(dt)) for _ in range(4))
    w = rng.random((5, 5)).astype(dt)
    sx = sb.StencilArray(dev(X), sb.Circle(4), boundary=sb.Wrap(), padding=sb.Halo("out"))
    sy = sb.StencilArray(dev(Y), sb.Kernel(sb.Window(2), w), boundary=sb.Reflect())
    sz = sb.StencilArray(dev(Z), sb.VonNeumann(2), boundary=sb.Remove(dt(0.5)))
    f = sb.LinearCombination((0.25, sb.maximum), sb.kernelproduct, (-1.5, sb.mean), (2.0, sb.center))

    def run():
        return host(sb.mapstencil(f, sx, sy, sz, dev(P)))
    monkeypatch.setenv("SB200_MULTI_SINGLE_PASS", "1")   # opt-in: the sweep-per-argument path measured faster (csrc/multi_tile.cu)
    got = run()
    assert l.sb200_last_kernel() == b"multi_tile2d_kernel"
    t1 = orc.stencil_array_sweep(X, npr.offsets("Circle", 4, 2), 4, A.WRAP, "cond", A.MAX)
    t2 = orc.stencil_array_sweep(Y, npr.offsets("Window", 2, 2), 2, A.REFLECT, "cond", A.KERNELDOT, weights=w)
    t3 = orc.stencil_array_sweep(Z, npr.offsets("VonNeumann", 2, 2), 2, A.REMOVE, "cond", A.MEAN, padval=0.5)
    want = ((dt(0.25) * t1 + t2) + dt(-1.5) * t3) + dt(2.0) * P
    assert want.dtype == np.dtype(dt)
    bits_equal(got, want)
    monkeypatch.delenv("SB200_MULTI_SINGLE_PASS")
    old = run()
    assert l.sb200_last_kernel() != b"multi_tile2d_kernel"
    bits_equal(old, got)
    monkeypatch.setenv("SB200_MULTI_SINGLE_PASS", "1")
    # Halo{:in} sources (the logical array is the inside of the parent) and a single argument with a coefficient
    big = np.asfortranarray(rng.random((140, 40)).astype(dt))
    si = sb.StencilArray(dev(big), sb.Moore(1), boundary=sb.Wrap(), padding=sb.Halo("in"))
    g1 = host(sb.mapstencil(sb.LinearCombination((3.0, sb.sum)), si))
    assert l.sb200_last_kernel() == b"multi_tile2d_kernel"
    inner = np.asfortranarray(big[1:-1, 1:-1])
    bits_equal(g1, dt(3.0) * orc.stencil_array_sweep(inner, npr.offsets("Moore", 1, 2), 1, A.WRAP, "cond", A.SUM))
    # 3-D: the sweep-per-argument path
    V, U = (np.asfortranarray(rng.random((40, 12, 9)).astype(dt)) for _ in range(2))
    sv, su = sb.StencilArray(dev(V), sb.VonNeumann(1, 3), boundary=sb.Wrap()), sb.StencilArray(dev(U), sb.Window(1, 3), boundary=sb.Wrap())
    g3 = host(sb.mapstencil(sb.LinearCombination(sb.sum, (0.5, sb.mean)), sv, su))
    assert l.sb200_last_kernel() != b"multi_tile2d_kernel"
    w3 = orc.stencil_array_sweep(V, npr.offsets("VonNeumann", 1, 3), 1, A.WRAP, "cond", A.SUM) + \
        dt(0.5) * orc.stencil_array_sweep(U, npr.offsets("Window", 1, 3), 1, A.WRAP, "cond", A.MEAN)
    bits_equal(g3, w3)


def test_layered_stencils(orc):
    """Layered (src/stencils/layered.jl:13-57; reference test test/stencils.jl:242-264): `sum(l[1]) - sum(l[2])` and
    `sum(l.l1.b) - sum(l.l2.a)` as multi-table gathers over ONE parent, on the reference's 5 x 5 array and on random Float32
    fields with Halo padding (ring sized by the largest layer radius), against per-layer oracle sweeps combined left to right."""
    p1, p2 = sb.Positional(((-1, -1), (1, 1))), sb.Positional(((-2, -2), (2, 2)))
    arr = np.asfortranarray(np.arange(1.0, 26.0).reshape(5, 5, order="F"))
    layered = sb.Layered(p1, p2)
    a = sb.StencilArray(dev(arr), layered)
    got = sb.mapstencil(sb.LinearCombination(sb.layer(0, sb.sum), (-1.0, sb.layer(1, sb.sum))), a)

    def lsum(x, st, bc=A.REMOVE, pad="cond", padval=0.0):
        return orc.stencil_array_sweep(np.asfortranarray(x), st.offsets(), st.radius, bc, pad, A.SUM, padval=padval)
    want = lsum(arr, p1) + (-1.0) * lsum(arr, p2)
    bits_equal(host(got), want)
    assert host(got)[2, 2] == (7 + 19) - (1 + 25)
    ml = sb.Layered(l1=sb.Layered(a=p1, b=p2), l2=sb.Layered(a=p1, b=p2))
    a2 = sb.StencilArray(dev(arr), ml)
    got2 = sb.mapstencil(sb.LinearCombination(sb.layer(("l1", "b"), sb.sum), (-1.0, sb.layer(("l2", "a"), sb.sum))), a2)
    bits_equal(host(got2), lsum(arr, p2) + (-1.0) * lsum(arr, p1))
    # random field, Halo{:out} ring of the LARGEST layer radius, Wrap: mean over Window(1) + 0.5 * maximum over Circle(3) - centre
    rng = np.random.default_rng(9)
    X = np.asfortranarray(rng.random((256, 96)).astype(np.float32) - 0.3)
    lay = sb.Layered(near=sb.Window(1), far=sb.Circle(3))
    for pad, padname in ((sb.Halo("out"), "cond"), (None, "cond")):
        sx = sb.StencilArray(dev(X), lay, boundary=sb.Wrap(), padding=pad)
        assert sx.halo == (3 if pad is not None else 0)
        f = sb.LinearCombination(sb.layer("near", sb.mean), (0.5, sb.layer("far", sb.maximum)), (-1.0, sb.layer("near", sb.center)))
        got3 = sb.mapstencil(f, sx)
        t1 = orc.stencil_array_sweep(X, npr.offsets("Window", 1, 2), 1, A.WRAP, padname, A.MEAN)
        t2 = orc.stencil_array_sweep(X, npr.offsets("Circle", 3, 2), 3, A.WRAP, padname, A.MAX)
        want3 = (t1 + np.float32(0.5) * t2) + np.float32(-1.0) * X
        bits_equal(host(got3), want3)
    with pytest.raises(sb.ArgumentError):
        sb.mapstencil(sb.LinearCombination(sb.sum, sb.sum), a)                      # terms must name their layer
    with pytest.raises(sb.ArgumentError):
        sb.mapstencil(sb.LinearCombination(sb.layer(5, sb.sum)), a)                 # no such layer
    with pytest.raises(sb.ArgumentError):
        sb.mapstencil(sb.LinearCombination(sb.layer("l1", sb.sum)), a2)             # a nested Layered is not a leaf


def test_kernel_from_distance_function(orc):
    """Kernel(f, hood) (src/stencils/kernel.jl:98-106): weights = f.(distances(hood)); kernelproduct on the GPU against the
    oracle with the same weights, for a Window, a Circle and a Positional hood, Float32 and Float64."""
    rng = np.random.default_rng(44)
    for dt in (np.float32, np.float64):
        r = np.asfortranarray((rng.random((256, 80)) - 0.25).astype(dt))
        for hood in (sb.Window(2), sb.Circle(3), sb.Positional((-1, 1), (-2, -1), (1, 0), (-2, 2))):
            k = sb.Kernel(lambda d: dt(1.0) / dt(1.0 + d), hood)
            assert len(k.kernel) == len(hood) and k.offsets() == hood.offsets()
            want_w = np.array([dt(1.0) / dt(1.0 + np.sqrt(float(builtins.sum(v * v for v in o)))) for o in hood.offsets()], dtype=dt)
            np.testing.assert_array_equal(np.asarray(k.kernel, dtype=dt), want_w)
            a = sb.StencilArray(dev(r), sb.Kernel(hood, np.asarray(k.kernel, dtype=dt)), boundary=sb.Wrap())
            got = sb.mapstencil(sb.kernelproduct, a)
            want = orc.stencil_array_sweep(r, hood.offsets(), hood.radius, A.WRAP, "cond", A.KERNELDOT, weights=want_w)
            bits_equal(host(got), want)


def test_shutdown_releases_and_the_library_keeps_working(orc):
    """sb200_shutdown frees the cached plans (device tables), the host-buffer scratch and the multi-gather scratch; the next call
    rebuilds what it needs (the one-entry plan memo must not hand out a freed plan)."""
    l = A.lib()
    r = np.asfortranarray(np.random.default_rng(3).random((256, 64)))
    a = sb.StencilArray(dev(r), sb.Window(1))
    want = orc.stencil_array_sweep(r, npr.offsets("Window", 1, 2), 1, A.REMOVE, "cond", A.MEAN, padval=0.0)
    for _ in range(3):
        out = sb.mapstencil(sb.mean, a)
        bits_equal(host(out), want)
        hostout = np.zeros_like(r, order="F")
        sb.mapstencil_(sb.mean, hostout, sb.StencilArray(r, sb.Window(1)))      # host path: allocates the scratch
        bits_equal(hostout, want)
        assert l.sb200_shutdown() == 0


def test_torch_free_case_table_matches_oracle():
    """tests/sanitize_cases.py (the compute-sanitizer driver: every kernel family on exact cudaMalloc allocations, no torch)
    as a plain parity run: every case bit-identical to the oracle."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "sanitize_cases.py")], cwd=root, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "all cases match the oracle" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
