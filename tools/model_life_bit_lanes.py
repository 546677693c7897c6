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


def strip_cover(W, G, one_halo_lane, warps=6, rowb=6144):
    """launch_bit / life_bit_kernel's column decomposition (strips, warps, lanes): returns how often each column of a row is stored
    and whether every active lane's inputs (its 32 cells, its neighbours' words, the end lanes' halo byte) lie inside the halo the
    producer copies."""
    hln = 1 if one_halo_lane else G - 1
    valid = 32 - 2 * hln
    wo, hl = valid * 32, hln * 32 + 16
    cap = warps * wo
    d0 = (128 - hl % 128) % 128
    assert d0 + 2 * hl + cap <= rowb and cap % 128 == 0
    nstrips = (W + cap - 1) // cap
    outb = min(cap, ((W + nstrips - 1) // nstrips + 127) // 128 * 128)
    nstrips = (W + outb - 1) // outb
    stored = np.zeros(W, int)
    inside = True
    for strip in range(nstrips):
        x0 = strip * outb
        wout = min(outb, W - x0)
        for w in range(warps):
            if w * wo >= wout:
                continue
            for lane in range(32):
                cell0 = w * wo + (lane - hln) * 32
                if hln <= lane <= 31 - hln and cell0 < wout:
                    stored[x0 + cell0:x0 + cell0 + 32] += 1
            # the warp reads cells [w*wo - hln*32 - 1, w*wo + (32 - hln)*32 + 1) relative to x0; the row holds [-hl, wout + hl).
            # Lanes whose cells start beyond wout + hl read stale shared memory, which is harmless as long as no ACTIVE lane's
            # dependence cone (G cells to either side after G generations) reaches them.
            last_active = max(l for l in range(32) if hln <= l <= 31 - hln and w * wo + (l - hln) * 32 < wout)
            need_hi = w * wo + (last_active - hln) * 32 + 32 + G
            inside &= (w * wo - hln * 32 - 1 >= -hl) and (need_hi <= wout + hl)
    return stored, inside


def packed_row_model(W, G, warps=6, rowb=768, hlb=16):
    """The packed source / dest forms of life_bit_kernel (IN == LB_BITS / OUT_BITS; csrc/life_bit.cuh): a row is W / 8 bytes, a strip's
    shared-memory row holds [x0 / 8 - 16, (x0 + wout) / 8 + 16) of it (mod W / 8: the producer's main copy + the two wrapped 16-byte
    halo copies), lane l of warp w reads the 32-bit word at byte 16 + w * 120 + (l - 1) * 4 and, when active, stores its word at byte
    (x0 + w * 960) / 8 + (l - 1) * 4 of the dest row. Emulates the byte movement with NumPy on a random row and returns
    (every dest word written exactly once with the word of the same cells, every lane's word is the right global word, every copy is
    16-byte aligned and sized)."""
    assert W % 128 == 0
    rng = np.random.default_rng(W + G)
    Wb = W // 8
    row = rng.integers(0, 256, Wb, dtype=np.uint8)
    wo = 30 * 32
    cap = warps * wo
    nstrips = (W + cap - 1) // cap
    outb = min(cap, ((W + nstrips - 1) // nstrips + 127) // 128 * 128)
    nstrips = (W + outb - 1) // outb
    written = np.zeros(Wb // 4, int)
    dest = np.zeros(Wb, dtype=np.uint8)
    words_ok, aligned = True, True
    for strip in range(nstrips):
        x0 = strip * outb
        wout = min(outb, W - x0)
        xb, wb = x0 // 8, wout // 8
        srow = np.full(rowb, 0xEE, dtype=np.uint8)   # stale shared memory
        lin = hlb if xb >= hlb else 0
        rin = min(hlb, Wb - (xb + wb))
        copies = [(hlb - lin, xb - lin, lin + wb + rin)]
        if not lin:
            copies.append((0, Wb - hlb, hlb))
        if rin < hlb:
            copies.append((hlb + wb + rin, 0, hlb - rin))
        for dst_off, src_off, n in copies:
            aligned &= dst_off % 16 == 0 and src_off % 16 == 0 and n % 16 == 0 and n > 0 and src_off + n <= Wb and dst_off + n <= rowb
            srow[dst_off:dst_off + n] = row[src_off:src_off + n]
        for w in range(warps):
            if w * wo >= wout:
                continue
            for lane in range(32):
                cell0 = w * wo + (lane - 1) * 32
                off = hlb + (w * wo) // 8 + (lane - 1) * 4
                word = srow[off:off + 4]
                active = 1 <= lane <= 30 and cell0 < wout
                # lanes next to an active lane feed its edge bits: their word must be the real neighbour word (mod W)
                need = active or (1 <= lane + 1 <= 30 and cell0 + 32 < wout) or (1 <= lane - 1 <= 30 and 0 <= cell0 - 32 < wout)
                if need:
                    g = ((x0 + cell0) // 8) % Wb
                    words_ok &= bool((word == row[g:g + 4]).all())
                if active:
                    d = (x0 + w * wo) // 8 + (lane - 1) * 4
                    dest[d:d + 4] = word
                    written[d // 4] += 1
    return bool((written == 1).all()) and bool((dest == row).all()), words_ok, aligned


def check_strips():
    ok = True
    for W in (1024, 4096, 4352, 9216, 16384, 32768):
        for G, one in ((2, False), (4, False), (2, True), (4, True), (8, True)):
            stored, inside = strip_cover(W, G, one)
            good = (stored == 1).all() and inside
            ok &= good
            print(f"W = {W:5d}, G = {G}, {'one halo lane ' if one else 'G-1 halo lanes'}: every column stored once = {(stored == 1).all()}, inputs inside the copied halo = {inside}")
    return ok


if __name__ == "__main__":
    raise SystemExit(0 if check_strips() else 1)
