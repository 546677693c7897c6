"""CPU tests of the host-side mirror (size bookkeeping, stencil algebra, error behaviour) and of the C-ABI
library's link surface. No compute call is made here: the library loads and resolves on a machine without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import stencils_b200 as sb
from stencils_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "stencils_b200.h")).read()
    declared = sorted(set(re.findall(r"^(?:int32_t|int64_t|size_t|const char\*)\s+(sb200_[a-z0-9_]+)\(", hdr, re.M)))
    assert declared == sorted(A.EXPORTED_SYMBOLS)
    lib = C.CDLL(A.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert A.lib().sb200_version() == 100


def test_desc_layout_matches_header():
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "stencils_b200.h"\nint main(){printf("%zu %zu %zu %zu",sizeof(sb200_desc),' \
          'offsetof(sb200_desc,offsets_host),offsetof(sb200_desc,alpha),offsetof(sb200_desc,flags));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    assert [int(v) for v in out] == [C.sizeof(A.Desc), A.Desc.offsets_host.offset, A.Desc.alpha.offset, A.Desc.flags.offset]


def test_offsets_match_goldens_through_the_abi(goldens):
    name2cls = {"Moore": sb.Moore, "Window": sb.Window, "VonNeumann": sb.VonNeumann, "Ordinal": sb.Ordinal,
                "Cardinal": sb.Cardinal}
    for key, g in goldens["offsets"].items():
        want = tuple(tuple(t) for t in g["v"])
        if g["shape"] == "Annulus":
            assert sb.Annulus(g["R"], g["RI"], g["N"]).offsets() == want, key
        else:
            assert name2cls[g["shape"]](g["R"], g["N"]).offsets() == want, key


def test_abi_offsets_equal_oracle_for_all_shapes(orc):
    classes = [sb.Window, sb.Moore, sb.VonNeumann, sb.Cross, sb.AngledCross, sb.ForwardSlash, sb.BackSlash, sb.Circle,
               sb.Vertical, sb.Horizontal, sb.Diamond, None, sb.Cardinal, sb.Ordinal]
    for enum, cls in enumerate(classes):
        for N in (1, 2, 3):
            for R in (1, 2, 3, 4):
                if cls is None:
                    got = sb.Annulus(R, R - 1, N).offsets()
                else:
                    got = cls(R, N).offsets()
                assert list(got) == orc.offsets(enum, R, N, R - 1), (enum, N, R)


def test_stencil_algebra():
    m = sb.Moore(1)
    assert sb.radius(m) == 1 and sb.diameter(m) == 3 and len(m) == 8                   # test/stencils.jl:18-23
    sq2 = 2 ** 0.5
    assert sb.distances(m) == (sq2, 1.0, sq2, 1.0, 1.0, sq2, 1.0, sq2)                 # test/stencils.jl:32
    assert sb.indices(m, (1, 1)) == ((0, 0), (1, 0), (2, 0), (0, 1), (2, 1), (0, 2), (1, 2), (2, 2))
    assert sb.indices(sb.Window(1, 2), (10, 10, 10)) == sb.indices(
        sb.Positional((-1, -1, 0), (0, -1, 0), (1, -1, 0), (-1, 0, 0), (0, 0, 0), (1, 0, 0), (-1, 1, 0), (0, 1, 0), (1, 1, 0)),
        (10, 10, 10))                                                                  # test/array.jl:113
    h1 = sb.Positional(((-1, -1), (2, -2), (2, 2), (-1, 2), (0, 0)))
    assert sb.radius(h1) == 2 and len(h1) == 5                                         # test/stencils.jl:143-144
    r = sb.Rectangle((-1, 0), (-2, 1))
    assert sb.radius(r) == 2 and len(r) == 8                                           # test/stencils.jl:167-168
    assert len(sb.Rectangle(((-1, 0), (0, 1), (1, 1)))) == 4                           # test/stencils.jl:177-179
    ns = sb.NamedStencil(n=(-1, 0), e=(0, -1), w=(1, 0), s=(0, 1))
    # test/stencils.jl:215 compares two *unfilled* stencils (StaticVector equality on `nothing` neighbours),
    # which only pins the length and the names:
    ns2 = sb.NamedStencil(("n", "e", "w", "s"), sb.VonNeumann())
    assert len(ns) == len(ns2) == 4 and ns.names == ns2.names and ns2.offsets() == sb.VonNeumann().offsets()
    with pytest.raises(sb.ArgumentError):
        sb.NamedStencil(("n", "s"), sb.VonNeumann())                                   # test/stencils.jl:216
    p = sb.merge(sb.Horizontal(), sb.Vertical())
    assert len(p) == 5 and list(p.offsets()) == sorted(p.offsets())                    # test/stencils.jl:319-321
    with pytest.raises(sb.ArgumentError):
        sb.merge(sb.Horizontal(1, 3), sb.Horizontal())                                 # test/stencils.jl:323
    with pytest.raises(sb.ArgumentError):
        sb.Kernel(sb.Window(2), np.arange(9))                                          # test/stencils.jl:272
    assert sb.Kernel(np.arange(1, 10).reshape(3, 3)).stencil == sb.Window(1, 2)        # test/stencils.jl:273-275


def test_filled_stencils_on_host(goldens):
    F = goldens["fills"]
    win = np.array(F["win_5x5"]["v"], dtype=np.int64)
    res1 = sb.stencil(sb.StencilArray(win, sb.Positional(*[tuple(o) for o in F["positional_h1_at_3_3"]["offsets"]])), (3, 3))
    assert list(sb.neighbors(res1)) == F["positional_h1_at_3_3"]["neighbors"] and sb.center(res1) == 0
    ns = sb.NamedStencil(n=(-1, 0), e=(0, -1), w=(1, 0), s=(0, 1))
    res = sb.stencil(sb.StencilArray(win, ns), (3, 3))
    assert list(sb.neighbors(res)) == F["named_h1_at_3_3"]["neighbors"] and res.n == 1 and res.e == 0  # test/stencils.jl:199-203
    A4 = sb.StencilArray(np.zeros((4, 4)), sb.Moore(1), boundary=sb.Wrap())
    assert [list(t) for t in sb.indices(A4, (1, 1))] == goldens["indices"]["wrap_4x4_at_1_1"]["v"]
    A4 = sb.StencilArray(np.zeros((4, 4)), sb.Moore(1), boundary=sb.Reflect())
    assert [list(t) for t in sb.indices(A4, (1, 1))] == goldens["indices"]["reflect_4x4_at_1_1"]["v"]
    S4 = sb.SwitchingStencilArray(np.zeros((4, 4)), sb.Moore(1), boundary=sb.Reflect())
    assert [list(t) for t in sb.indices(S4, (1, 1))] == goldens["indices"]["reflect_4x4_at_1_1"]["v"]


@pytest.mark.parametrize("nd", [1, 2, 3])
def test_size_bookkeeping(nd):
    """test/array.jl:16-107: Conditional keeps the size, Halo{:out} pads the parent, Halo{:in} shrinks the view."""
    n = 100 if nd < 3 else 30
    r = np.asfortranarray(np.random.default_rng(0).random((n,) * nd))
    r0 = r.copy()
    R = 10
    a = sb.StencilArray(r, sb.VonNeumann(R, nd), padding=sb.Conditional(), boundary=sb.Remove(0.0))
    b = sb.StencilArray(r, sb.Window(R, nd), padding=sb.Halo("out"), boundary=sb.Remove(0.0))
    c = sb.StencilArray(r, sb.Moore(R, nd), padding=sb.Halo("in"), boundary=sb.Remove(0.0))
    sa = sb.SwitchingStencilArray(r, sb.Window(R, nd), padding=sb.Conditional(), boundary=sb.Remove(0.0))
    sb_ = sb.SwitchingStencilArray(r, sb.Window(R, nd), padding=sb.Halo("out"), boundary=sb.Remove(0.0))
    sc = sb.SwitchingStencilArray(r, sb.Window(R, nd), padding=sb.Halo("in"), boundary=sb.Remove(0.0))
    assert sa.source is not sa.dest
    assert a.shape == a.parent.shape == sa.shape == sa.parent.shape == (n,) * nd
    assert b.shape == sb_.shape == (n,) * nd and b.parent.shape == sb_.parent.shape == (n + 2 * R,) * nd
    assert c.shape == sc.shape == (n - 2 * R,) * nd and c.parent.shape == sc.parent.shape == (n,) * nd
    assert a.similar().shape == a.shape and b.similar().shape == b.shape and c.similar().shape == c.shape
    c[...] = 0.0
    assert (np.asarray(c) == 0).all()
    assert np.array_equal(np.asarray(b), r0)
    assert c.parent is r and not np.array_equal(r, r0)  # Halo{:in} wraps the user's own array (src/padding.jl:103)
    with pytest.raises(sb.ArgumentError):
        sb.StencilArray(np.zeros((5,) * nd), sb.Window(5, nd))  # src/array.jl:451-453


def test_unsupported_user_function_raises_without_fallback():
    a = sb.StencilArray(np.zeros((8, 8)), sb.Window(1))
    with pytest.raises(sb.ArgumentError, match="no fallback"):
        sb.mapstencil(lambda h: h.center, a)
    with pytest.raises(sb.ArgumentError):
        sb.mapstencil(sb.mean, a, np.zeros((8, 8)))  # extra array args: SURVEY §8f.1 (next)
    with pytest.raises(sb.ArgumentError):
        sb.mapstencil(sb.kernelproduct, a)           # needs a Kernel stencil
    with pytest.raises(sb.ArgumentError):
        sb.mapstencil(sb.Diffusion(0.1), sb.StencilArray(np.zeros((8, 8), dtype=np.int32), sb.VonNeumann(1)))
    with pytest.raises(sb.ArgumentError):
        sb.mapstencil(sb.mean, sb.StencilArray(np.zeros((8, 8)), sb.Window(1), boundary=sb.Remove()))
    with pytest.raises(sb.ArgumentError):
        sb.mapstencil(sb.mean, sb.StencilArray(np.zeros((8, 8), dtype=np.float16), sb.Window(1)))
    with pytest.raises(sb.ArgumentError):
        sb.mapstencil_(sb.mean, np.zeros((7, 8), order="F"), a)  # _checksizes


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "stencils.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("not a cpu fallback", ""), os.path.join(dirpath, f)


def test_scatter_slab_boundaries_are_pass_aligned():
    """split_columns_for_scatter: contiguous cover of the columns, every interior boundary a multiple of the pass stride
    2R+1 (so a column's pass number in _scatterstencil_cpu!, src/scatterstencil.jl:53-58, is the same in local and global
    numbering), remainder units to the first ranks."""
    from stencils_b200.slab import split_columns_for_scatter
    for ncols in (1, 5, 33, 37, 45, 1000, 32768):
        for R in (1, 2, 3, 4):
            S = 2 * R + 1
            for world in (1, 2, 3, 8):
                parts = [split_columns_for_scatter(ncols, world, r, R) for r in range(world)]
                assert parts[0][0] == 0 and parts[-1][1] == ncols
                for (lo, hi), (lo2, _) in zip(parts, parts[1:]):
                    assert hi == lo2 and lo <= hi
                    assert hi % S == 0 or hi == ncols
                units = -(-ncols // S)
                assert max(hi - lo for lo, hi in parts) <= -(-units // world) * S   # remainder units go to the first ranks


def test_slab_scatter_rejects_what_it_cannot_do_exactly():
    """Wrap on the split axis with a column count that is not a multiple of 2R+1 has no defined pass order at the seam
    (SURVEY Appendix A): ArgumentError, not an approximate result."""
    import numpy as np
    import torch
    from stencils_b200 import _abi as A
    from stencils_b200.slab import slab_gather, slab_scatter
    src = torch.zeros((37, 16), dtype=torch.float32)
    with pytest.raises(A.ArgumentError):
        slab_scatter(src, src.clone(), ncols_global=37, offsets=[(0, 1), (1, 0)], radius=1, weights=np.ones(2, np.float32),
                     boundary=(A.WRAP, A.WRAP), eltype=A.F32, rank=0, world=1, compute=lambda *a: None)
    with pytest.raises(A.ArgumentError):
        slab_scatter(src, torch.zeros((36, 16)), ncols_global=36, offsets=[(0, 1)], radius=1, weights=np.ones(1, np.float32),
                     boundary=(A.REMOVE, A.REMOVE), eltype=A.F32, rank=0, world=1, compute=lambda *a: None)
    with pytest.raises(A.ArgumentError):   # mean of UInt8 changes the element type
        slab_gather(torch.zeros((8, 8), dtype=torch.uint8), offsets=[(0, 1)], radius=1, reducer=A.MEAN, boundary=(A.WRAP, A.WRAP),
                    eltype=A.U8, rank=0, world=1, compute=lambda *a: None)


def test_inexact_weights_and_padval_are_refused():
    """The reference promotes (kernelproduct accumulates hood[i] * kernel[i], src/stencils/kernel.jl:37-43; the result type
    follows typeof(padval)); these kernels compute in the source element type, so values that type cannot hold are an
    ArgumentError in the user-facing layer instead of a silent cast (Kernel(Window(1), fill 0.1) on Int32 built an
    all-zero table)."""
    from stencils_b200._desc import require_exact
    from stencils_b200.ops import _desc_for
    for vals, dt in [(np.full(3, 0.1), np.int32), (np.full(3, 0.1), np.float32), ([0.5], np.int64), ([2], np.bool_), ([300], np.uint8),
                     ([-1], np.uint8)]:
        with pytest.raises(A.ArgumentError):
            require_exact(np.asarray(vals), np.dtype(dt), "x")
    for vals, dt in [(np.full(3, 0.1, dtype=np.float32), np.float32), (np.array([1.0, 2.0, -3.0]), np.int32), ([0.1, 0.2], np.float64),
                     ([float("nan"), -0.0], np.float32), ([True], np.bool_), ([1], np.bool_)]:
        require_exact(np.asarray(vals), np.dtype(dt), "x")
    # through the descriptor builder of mapstencil / scatterstencil
    src = np.zeros((8, 8), dtype=np.int32, order="F")
    dst = np.zeros((8, 8), dtype=np.int32, order="F")
    with pytest.raises(A.ArgumentError, match="cannot be represented exactly"):
        _desc_for(sb.kernelproduct, src, 0, dst, 0, sb.Kernel(sb.Window(1), np.full((3, 3), 0.1)), sb.Remove(0))
    with pytest.raises(A.ArgumentError, match="cannot be represented exactly"):
        _desc_for(sb.sum, src, 0, dst, 0, sb.Window(1), sb.Remove(0.5))
    _desc_for(sb.kernelproduct, src, 0, dst, 0, sb.Kernel(sb.Window(1), np.full((3, 3), 2.0)), sb.Remove(0))


def test_layered_goldens():
    """test/stencils.jl:242-264: radius / offsets / indices / neighbors / center of a Layered stencil and nested named layers."""
    p1, p2 = sb.Positional(((-1, -1), (1, 1))), sb.Positional(((-2, -2), (2, 2)))
    layered = sb.Layered(p1, p2)
    assert sb.radius(layered) == 2
    assert sb.offsets(layered) == (((-1, -1), (1, 1)), ((-2, -2), (2, 2)))
    assert sb.indices(layered, (1, 1)) == (((0, 0), (2, 2)), ((-1, -1), (3, 3)))
    arr = np.asfortranarray(np.arange(1, 26).reshape(5, 5, order="F"))
    a = sb.StencilArray(arr, layered)
    filled = a.stencil_at(3, 3)
    assert sb.neighbors(filled) == ((7, 19), (1, 25))
    assert sb.center(filled) == (13, 13)
    l1, l2 = sb.Layered(a=p1, b=p2), sb.Layered(a=p1, b=p2)
    ml = sb.Layered(l1=l1, l2=l2)
    filled2 = sb.StencilArray(arr, ml).stencil_at(3, 3)
    assert filled2.l2.a.neighbors == (7, 19)
    assert ml.lengths() == ((2, 2), (2, 2)) and ml[("l1", "b")] == p2 and sb.radius(ml) == 2
    # Halo padding is sized by the largest layer radius (src/array.jl:471, src/padding.jl:104-110)
    ah = sb.StencilArray(arr.astype(np.float64), layered, padding=sb.Halo("out"))
    assert ah.parent.shape == (9, 9) and ah.shape == (5, 5)
    with pytest.raises(sb.ArgumentError):
        sb.Layered()


def test_iterate_step_split_is_optimal_and_keeps_the_buffer_parity():
    """split_steps (csrc/api.cu) behind sb200_debug_split_steps: the launches sb200_iterate issues for a step count add up to it,
    their number has the parity of the step count (the final state lands in the buffer the contract names), and for step
    counts a brute-force search can cover the split has the least total cost under the library's launch-time table."""
    l = A.lib()
    life_cost = [0, 1.00, 1.13, 1.16, 1.14, 1.07, 1.17, 1.31, 1.45]   # kLifeLaunchCost
    diff_cost = [0, 1.00, 1.38]                                        # kDiffCost

    def split(n, mask, life):
        out = (C.c_int32 * 9)()
        A.check(l.sb200_debug_split_steps(n, mask, life, out))
        return list(out)

    def best(n, sizes, cost):
        INF = float("inf")
        dp = [[INF, INF] for _ in range(n + 1)]
        dp[0][0] = 0.0
        for m in range(1, n + 1):
            for q in (0, 1):
                dp[m][q] = min((dp[m - g][q ^ 1] + cost[g] for g in sizes if g <= m), default=INF)
        return dp[n][n & 1]

    for mask, life, cost in [(0x1FC, 1, life_cost), (0x114, 1, life_cost), (0x14, 1, life_cost), (0x4, 1, life_cost), (0x4, 0, diff_cost), (0, 1, life_cost)]:
        sizes = [1] + [g for g in range(2, 9) if (mask >> g) & 1]
        for n in list(range(0, 150)) + [500, 1000, 1001, 4096, 99999]:
            c = split(n, mask, life)
            assert c[0] == 0 and sum(g * c[g] for g in range(9)) == n, (mask, n, c)
            assert all(c[g] == 0 for g in range(9) if g not in sizes), (mask, n, c)
            assert sum(c) % 2 == n % 2, (mask, n, c)
            total = sum(cost[g] * c[g] for g in sizes)
            assert total <= best(n, sizes, cost) * (1 + 1e-9) + 1e-9, (mask, n, c, total, best(n, sizes, cost))
    assert split(20, 0x1FC, 1) == [0, 0, 0, 0, 0, 4, 0, 0, 0]             # the driver's --steps 20: 5 + 5 + 5 + 5
    assert split(20, 0x114, 1) == [0, 0, 0, 0, 3, 0, 0, 0, 1]             # powers of two only: 4 + 4 + 4 + 8
    assert split(1000, 0x1FC, 1)[8] >= 120                                  # long runs: launches of eight generations
    assert split(100, 0x4, 0) == [0, 0, 50, 0, 0, 0, 0, 0, 0]
    assert split(101, 0x4, 0) == [0, 1, 50, 0, 0, 0, 0, 0, 0]
