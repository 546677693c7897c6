"""CPU models of kernel index logic (tools/model_*.py) as regression tests: the models transliterate the kernels' task / ring /
lane arithmetic thread for thread and are compared with plain NumPy sweeps; here we also check that the constants the models
use are the ones in the CUDA sources, so a kernel change without a model change fails on a box without a GPU."""
import importlib.util
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _cu_const(text, name):
    m = re.search(r"constexpr int (?:[A-Z0-9_]+ = [^,;]+, )?" + name + r" = ([^;,]+)[;,]", text)
    assert m, name
    return m.group(1).strip()


def test_stream3d2_model_uses_the_kernel_constants():
    m = _load("model_stream3d2")
    cu = open(os.path.join(ROOT, "stencils.jl_b200", "csrc", "stream3d2.cu")).read()
    assert re.search(r"constexpr int D2_WX = (\d+), D2_WY = (\d+);", cu).groups() == (str(m.WX), str(m.WY))
    assert _cu_const(cu, "D2_RT") == str(m.RT)
    assert _cu_const(cu, "D2_TY") == "(D2_WY - 1) * D2_RT" and m.TY == (m.WY - 1) * m.RT
    assert _cu_const(cu, "D2_LEFT") == str(m.LEFT)
    assert _cu_const(cu, "D2_ROWB") == "D2_LEFT + D2_TXB + 128" and m.ROWB == m.LEFT + m.TXB + 128
    assert _cu_const(cu, "D2_ROWS") == "D2_TY + 4" and m.ROWS == m.TY + 4
    assert _cu_const(cu, "D2_MROWS") == "D2_TY + 2" and m.MROWS == m.TY + 2
    assert re.search(r"#define SB200_D2_STAGES (\d+)", cu).group(1) == str(m.STAGES)


def test_stream3d2_model_matches_two_sweeps():
    m = _load("model_stream3d2")
    R, W = "remove", "wrap"
    assert m.check((160, 17, 8), np.float32)                                   # Wrap, second x-warp partly active, ragged y
    assert m.check((64, 30, 20), np.float32, ctas=4, force_nz=2)               # z-runs
    assert m.check((40, 15, 12), np.float64)
    assert m.check((128, 14, 16), np.float32, z_lo=3, zn=9, wrap_z=False)      # interior region (slab sweep)
    assert m.check((300, 30, 8), np.float32, bcs=(R, R, None), pad=0.5, ctas=2, force_ty=14)   # PAD variant
    assert m.check((64, 20, 9), np.float64, bcs=(R, W, R), pad=0.25)


def test_life_bit_lane_scheme_needs_one_halo_lane():
    m = _load("model_life_bit_lanes")
    rng = np.random.default_rng(2)
    for G in (2, 4, 8):
        f = (rng.random((40, 1024 + 256)) < 0.4).astype(np.uint8)
        want = f
        for _ in range(G):
            want = m.true_life(want)
        got = m.scheme(f, G, 128)
        exact = (got == want[:, 128:128 + 1024]).all(0).reshape(32, 32).all(1)
        assert exact[1:31].all()            # lanes 1 .. 30 are exact whatever G


def test_life_bit_strip_decomposition_covers_every_column_once():
    """launch_bit's strips / warps / lanes for the measured layout (G - 1 halo lanes) and the one-halo-lane experiment."""
    m = _load("model_life_bit_lanes")
    cu = open(os.path.join(ROOT, "stencils.jl_b200", "csrc", "life_bit.cuh")).read()
    assert "constexpr int LB_WARPS = 6;" in cu and "constexpr int LB_ROWB = 6144;" in cu
    assert "static constexpr int HL = HLN * 32 + 16;" in cu and "static constexpr int VALID = 32 - 2 * HLN;" in cu
    for W in (1024, 4352, 16384):
        for G, one in ((2, False), (4, False), (4, True), (8, True)):
            stored, inside = m.strip_cover(W, G, one)
            assert (stored == 1).all() and inside, (W, G, one)


def test_life_bit_packed_rows_model():
    """The packed source / dest forms (SB200_FLAG_SRC_BITS / _DST_BITS): byte-level model of the producer's three bulk copies per row
    (16-byte aligned, Wrap halos), the lanes' word offsets and the dest word offsets, tied to the constants in the CUDA source."""
    m = _load("model_life_bit_lanes")
    cu = open(os.path.join(ROOT, "stencils.jl_b200", "csrc", "life_bit.cuh")).read()
    assert "constexpr int LB_HLB = 16;" in cu and "constexpr int LB_ROWB_PK = 768;" in cu
    assert "LB_HLB + ((warp * C::WO) >> 3) + (lane - C::HLN) * 4" in cu and "((x0 + warp * C::WO) >> 3) + (lane - C::HLN) * 4" in cu
    for W in (512, 1024, 4352, 5760, 5888, 16384, 32768):
        for G in (1, 6, 8):
            once, words, aligned = m.packed_row_model(W, G)
            assert once and words and aligned, (W, G, once, words, aligned)


def test_life_bit_conway_identity_matches_the_kernel_source():
    """conway_bits in csrc/life_bit.cuh evaluates B3/S23 from the bit-sliced row sums with three immediate LOP3 tables (0x14, 0x42,
    0xCA). The tables are read out of the CUDA source and evaluated over every input combination against the rule
    itself: alive' = (T == 3) | (centre & T == 4), T = the 3 x 3 total including the centre."""
    import itertools
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stencils.jl_b200", "csrc", "life_bit.cuh")).read()
    body = src[src.index("__device__ __forceinline__ unsigned conway_bits"):]
    body = body[:body.index("\n}") + 2]
    imms = [int(v, 16) for v in re.findall(r"lop3_imm<(0x[0-9A-Fa-f]+)>", body)]
    assert imms == [0x14, 0x42, 0xCA], imms
    assert "lop3_imm<0x14>(u1, c0, c1)" in body and "lop3_imm<0x42>(u1, c0, c1) & b.c" in body and "lop3_imm<0xCA>(t0, a3, b4)" in body

    def lop3(a, b, c, imm):
        return (imm >> ((a << 2) | (b << 1) | c)) & 1

    def maj(a, b, c):
        return (a & b) | (a & c) | (b & c)
    for as0, as1, bs0, bs1, ns0, ns1, c in itertools.product([0, 1], repeat=7):
        t0, c0 = as0 ^ bs0 ^ ns0, maj(as0, bs0, ns0)
        u1, c1 = as1 ^ bs1 ^ ns1, maj(as1, bs1, ns1)
        a3 = lop3(u1, c0, c1, imms[0])
        b4 = lop3(u1, c0, c1, imms[1]) & c
        got = lop3(t0, a3, b4, imms[2])
        total = (as0 + 2 * as1) + (bs0 + 2 * bs1) + (ns0 + 2 * ns1)
        assert got == int(total == 3 or (c == 1 and total == 4)), (as0, as1, bs0, bs1, ns0, ns1, c)
