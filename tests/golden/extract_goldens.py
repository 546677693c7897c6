#!/usr/bin/env python
"""Extracts the golden vectors the reference's own tests hold for the hot path into
tests/golden/reference_goldens.json.

The reference is pure Julia and cannot be executed in this image, so the vectors are lifted
mechanically from the literals in /root/reference/test/array.jl and test/stencils.jl (file:line kept
with every entry). Run here (the container with /root/reference); the JSON is committed and is the only
thing the tests read — /root/reference does not exist on the GPU box.

    python tests/golden/extract_goldens.py
"""
import json
import os
import re

REF = "/root/reference/test"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")


def lines(fname, lo, hi):
    with open(os.path.join(REF, fname)) as f:
        all_lines = f.read().split("\n")
    return all_lines[lo - 1:hi]


def tuples(fname, lo, hi):
    """All (a, b[, c]) integer tuples on lines lo..hi, in order."""
    txt = " ".join(lines(fname, lo, hi))
    txt = txt[txt.index("=="):] if "==" in txt else txt
    return [[int(v) for v in m.split(",")] for m in re.findall(r"\(([-\d, ]+)\)", txt)]


def matrix(fname, lo, hi):
    """Julia matrix literal: rows on separate lines lo..hi (exclusive of the bracket lines)."""
    rows = []
    for ln in lines(fname, lo, hi):
        if "[" in ln:
            ln = ln[ln.index("[") + 1:]
        if "]" in ln:
            ln = ln[:ln.index("]")]
        ln = ln.strip()
        if ln:
            rows.append([float(v) for v in ln.split()])
    assert len({len(r) for r in rows}) == 1, (fname, lo, hi)
    return rows


def vector(fname, line):
    ln = lines(fname, line, line)[0]
    body = ln[ln.rindex("[") + 1:ln.rindex("]")]
    return [float(v) for v in body.split(",")]


def svector(fname, line):
    ln = lines(fname, line, line)[0]
    body = ln[ln.rindex("SVector(") + 8:ln.rindex(")")]
    return [int(v) for v in body.split(",")]


def main():
    g = {"_about": "golden vectors lifted from rafaqz/Stencils.jl v0.3.6 test/*.jl by extract_goldens.py"}
    A, S = "array.jl", "stencils.jl"
    # ---- offsets / indices (1-based indices as in the reference) ----
    g["offsets"] = {
        "Moore{1,2}": {"ref": "test/stencils.jl:28", "shape": "Moore", "R": 1, "N": 2, "v": tuples(S, 28, 28)},
        "Window{1,2}": {"ref": "test/stencils.jl:48-49", "shape": "Window", "R": 1, "N": 2, "v": tuples(S, 48, 49)},
        "VonNeumann{1,2}": {"ref": "test/stencils.jl:63", "shape": "VonNeumann", "R": 1, "N": 2, "v": tuples(S, 63, 63)},
        "VonNeumann{2,2}": {"ref": "test/stencils.jl:74-76", "shape": "VonNeumann", "R": 2, "N": 2, "v": tuples(S, 74, 76)},
        "Annulus{1,0,2}": {"ref": "test/stencils.jl:83-84", "shape": "Annulus", "R": 1, "RI": 0, "N": 2, "v": tuples(S, 83, 84)},
        "Annulus{2,1,2}": {"ref": "test/stencils.jl:95-96", "shape": "Annulus", "R": 2, "RI": 1, "N": 2, "v": tuples(S, 95, 96)},
        "Ordinal{1,2}": {"ref": "test/stencils.jl:103", "shape": "Ordinal", "R": 1, "N": 2, "v": tuples(S, 103, 103)},
        "Ordinal{2,2}": {"ref": "test/stencils.jl:114", "shape": "Ordinal", "R": 2, "N": 2, "v": tuples(S, 114, 114)},
        "Cardinal{1,2}": {"ref": "test/stencils.jl:121", "shape": "Cardinal", "R": 1, "N": 2, "v": tuples(S, 121, 121)},
        "Cardinal{2,2}": {"ref": "test/stencils.jl:132", "shape": "Cardinal", "R": 2, "N": 2, "v": tuples(S, 132, 132)},
        "Kernel(Window{1,2})": {"ref": "test/stencils.jl:280-281", "shape": "Window", "R": 1, "N": 2, "v": tuples(S, 280, 281)},
    }
    g["indices"] = {
        "moore_at_1_1": {"ref": "test/stencils.jl:29-30", "center": [1, 1], "v": tuples(S, 30, 30)},
        "kernel_window_at_2_2": {"ref": "test/stencils.jl:282-283", "center": [2, 2], "v": tuples(S, 282, 283)},
        "remove_4x4_at_1_1": {"ref": "test/array.jl:7", "boundary": "remove", "size": [4, 4], "center": [1, 1], "v": tuples(A, 7, 7)},
        "wrap_4x4_at_1_1": {"ref": "test/array.jl:9", "boundary": "wrap", "size": [4, 4], "center": [1, 1], "v": tuples(A, 9, 9)},
        "reflect_4x4_at_1_1": {"ref": "test/array.jl:11", "boundary": "reflect", "size": [4, 4], "center": [1, 1], "v": tuples(A, 11, 11)},
    }
    # ---- mapstencil goldens; input r = (1.0:5.0) * (100.0:105.0)' (test/array.jl:123,187,234) ----
    g["mapstencil"] = {
        "input_2d": {"ref": "test/array.jl:123", "julia": "(1.0:5.0) * (100.0:105.0)'"},
        "input_1d": {"ref": "test/array.jl:167,212", "julia": "collect(1.0:5.0)"},
        "remove_mean_2d": {"ref": "test/array.jl:139-145", "approx": True, "v": matrix(A, 140, 144)},
        "remove_sum_2d_interior": {"ref": "test/array.jl:157-161", "approx": False, "v": matrix(A, 158, 160)},
        "wrap_mean_1d": {"ref": "test/array.jl:181", "approx": False, "v": vector(A, 181)},
        "wrap_mean_1d_halo_in": {"ref": "test/array.jl:182", "approx": False, "v": vector(A, 182)},
        "wrap_mean_2d": {"ref": "test/array.jl:200-206", "approx": True, "v": matrix(A, 201, 205)},
        "reflect_mean_1d": {"ref": "test/array.jl:228", "approx": True, "v": vector(A, 228)},
        "reflect_mean_1d_halo_in": {"ref": "test/array.jl:229", "approx": True, "v": vector(A, 229)},
        "reflect_mean_2d": {"ref": "test/array.jl:247-253", "approx": True, "v": matrix(A, 248, 252)},
    }
    # ---- stencil fills on small integer matrices (test/stencils.jl) ----
    win = matrix(S, 136, 140)
    g["fills"] = {
        "win_5x5": {"ref": "test/stencils.jl:136-140", "v": win},
        "init_6x6": {"ref": "test/stencils.jl:3-8", "v": matrix(S, 3, 8)},
        "positional_h1_at_3_3": {"ref": "test/stencils.jl:141-148", "offsets": tuples(S, 141, 141),
                                 "neighbors": svector(S, 146), "sum": 2},
        "positional_h2_at_3_3": {"ref": "test/stencils.jl:150-157", "offsets": tuples(S, 150, 150),
                                 "neighbors": svector(S, 154), "sum": 0},
        "rectangle_h1_at_3_3": {"ref": "test/stencils.jl:160-174", "A": matrix(S, 160, 164), "axis_ranges": [[-1, 0], [-2, 1]],
                                "neighbors": svector(S, 172), "sum": 2, "radius": 2, "length": 8},
        "vonneumann_init_at_2_2": {"ref": "test/stencils.jl:60-71", "neighbors": svector(S, 70), "sum": 3},
        "named_h1_at_3_3": {"ref": "test/stencils.jl:193-201", "offsets": [[-1, 0], [0, -1], [1, 0], [0, 1]],
                            "neighbors": svector(S, 199), "sum": 3},
    }
    # ---- full-matrix mapstencil goldens with named offsets (test/stencils.jl:205-239) ----
    g["named_maps"] = {
        # s.n + s.w + center(s) with n=(-1,0), w=(1,0)  == golden + win
        "n_plus_w_plus_center": {"ref": "test/stencils.jl:205-213", "offsets": [[-1, 0], [1, 0], [0, 0]],
                                 "golden_minus_input": matrix(S, 208, 212)},
        # NamedStencil(Cardinal(1)) names (:E,:S,:N,:W) over offsets (0,-1),(-1,0),(1,0),(0,1); s.W + s.S
        "cardinal_W_plus_S": {"ref": "test/stencils.jl:219-228 + src/stencils/named.jl:91", "offsets": [[0, 1], [-1, 0]],
                              "v": matrix(S, 223, 227)},
        # NamedStencil(Ordinal(1)) names (:SE,:NE,:SW,:NW) over (-1,-1),(1,-1),(-1,1),(1,1); s.NE + s.NW
        "ordinal_NE_plus_NW": {"ref": "test/stencils.jl:230-239 + src/stencils/named.jl:92", "offsets": [[1, -1], [1, 1]],
                               "v": matrix(S, 234, 238)},
    }
    g["kernelproduct"] = {
        "window_1to9": {"ref": "test/stencils.jl:270-277", "hood": list(range(1, 10)), "kernel": list(range(1, 10)), "v": 285},
        "moore_vals": {"ref": "test/stencils.jl:286-289", "hood": [1, 2, 3, 4, 6, 7, 8, 9], "kernel": [1, 2, 3, 4, 6, 7, 8, 9],
                       "v": sum(v * v for v in [1, 2, 3, 4, 6, 7, 8, 9])},
        "positional_60": {"ref": "test/stencils.jl:296-303", "offsets": [[0, -1], [-1, 0], [1, 0], [0, 1]],
                          "win_3x3_colmajor": list(range(1, 10)), "kernel": [1, 2, 3, 4], "v": 60},
    }
    # ---- scatterstencil! (test/array.jl:385-493): scalar checks on 1-based cells ----
    g["scatter"] = {
        "moore_add_0.1": {"ref": "test/array.jl:386-403", "size": [5, 5], "src_fill": 1.0, "shape": "Moore", "R": 1, "op": "add",
                          "rule": "weights", "w": 0.1, "approx": True,
                          "cells": {"3,3": 0.8, "1,3": 0.5, "3,1": 0.5, "1,1": 0.3, "5,5": 0.3}},
        "moore_max_center": {"ref": "test/array.jl:405-421", "size": [5, 5], "src": "i+j", "shape": "Moore", "R": 1, "op": "max",
                             "rule": "center_weights", "w": 1.0, "approx": False, "cells": {"3,3": 8.0, "1,1": 4.0}},
        "vonneumann_add_0.25": {"ref": "test/array.jl:449-465", "size": [5, 5], "src_fill": 1.0, "shape": "VonNeumann", "R": 1,
                                "op": "add", "rule": "weights", "w": 0.25, "approx": True,
                                "cells": {"3,3": 1.0, "1,3": 0.75, "1,1": 0.5}},
        "moore2_add_0.01": {"ref": "test/array.jl:467-478", "size": [7, 7], "src_fill": 1.0, "shape": "Moore", "R": 2, "op": "add",
                            "rule": "weights", "w": 0.01, "approx": True, "cells": {"4,4": 0.24}},
        "switching_moore_add_0.1": {"ref": "test/array.jl:480-492", "size": [5, 5], "src_fill": 1.0, "shape": "Moore", "R": 1,
                                    "op": "add", "rule": "weights", "w": 0.1, "approx": True, "zero_dest": True,
                                    "cells": {"3,3": 0.8}},
    }
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
