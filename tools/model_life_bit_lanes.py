"""Which lanes of a warp hold exact cells after G generations in life_bit_kernel's scheme (csrc/life.cu)?

A warp row is 32 lanes x 32 cells (one bit each). Level 0 is exact for every lane (the end lanes read one real halo byte);
from generation 1 on the shifted-in edge bit of a lane comes from the adjacent lane by shuffle, and the warp's end lanes get
their OWN word back from __shfl_up/down (no neighbour), i.e. a wrong edge bit. This NumPy emulation runs the scheme on random
fields next to the true B3/S23 evolution of a wider torus and prints, per G, the lanes whose 32 cells are all exact for every
trial: contamination enters at one cell per generation from each end, so ONE halo lane per side is enough for G <= 32
(the kernel currently gives up G-1 lanes per side)."""
import numpy as np


def true_life(a):
    n = sum(np.roll(np.roll(a, dy, 0), dx, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)) - a
    return ((n == 3) | ((a == 1) & (n == 2))).astype(np.uint8)


def scheme(field, G, x0):
    """field: wide torus [rows, cols]; the warp covers columns x0 .. x0+1023. Returns generation G of those columns as the
    kernel's lanes would compute it (rows wrap exactly; only the lane-edge handling is modelled)."""
    rows = field.shape[0]
    cur = field[:, x0:x0 + 1024].copy()                       # level 0 cells
    left = field[:, (x0 - 1) % field.shape[1]].copy()         # real halo cells of level 0
    right = field[:, (x0 + 1024) % field.shape[1]].copy()
    for g in range(G):
        # horizontal neighbours per lane: inside a lane exact; across lanes by shuffle; the end lanes read the halo at level 0 and get
        # their own word (bit 31 of lane 0 / bit 0 of lane 31) afterwards
        L = np.empty_like(cur)
        R = np.empty_like(cur)
        L[:, 1:] = cur[:, :-1]
        R[:, :-1] = cur[:, 1:]
        if g == 0:
            L[:, 0], R[:, -1] = left, right
        else:
            L[:, 0] = cur[:, 31]                               # __shfl_up at lane 0 returns its own word: bit 31
            R[:, -1] = cur[:, 1024 - 32]                       # __shfl_down at lane 31: bit 0 of its own word
        s = L + cur + R                                       # horizontal 3-sums
        t = np.roll(s, 1, 0) + s + np.roll(s, -1, 0)          # 3x3 total including the centre
        cur = ((t == 3) | ((cur == 1) & (t == 4))).astype(np.uint8)
    return cur


def main():
    rng = np.random.default_rng(1)
    for G in (2, 4, 8, 16, 31):
        ok = np.ones(32, bool)
        for _ in range(6):
            f = (rng.random((48, 1024 + 256)) < 0.4).astype(np.uint8)
            want = f
            for _ in range(G):
                want = true_life(want)
            got = scheme(f, G, 128)
            same = (got == want[:, 128:128 + 1024]).all(0).reshape(32, 32).all(1)
            ok &= same
        lanes = np.flatnonzero(ok)
        print(f"G = {G:2d}: exact lanes {lanes.min()} .. {lanes.max()} ({len(lanes)} of 32); the kernel uses {32 - 2 * (G - 1)}")


if __name__ == "__main__":
    main()
