"""CPU model of stream3d2_kernel (csrc/stream3d2.cu): the kernel's index arithmetic transliterated thread for thread
(tasks, ring slots, producer row / halo copies, warp spans, end-lane cells, the two register levels), lanes vectorised
with NumPy. Uncopied shared memory is NaN, so any read the producer did not cover poisons the result. Compared with
two plain diffusion sweeps in the reference's fold order. Run on a box without a GPU to check the kernel's logic:

    python tools/model_stream3d2.py
"""
import sys

import numpy as np

WX, WY, RT = 2, 7, 2
WARPS = WX * WY
TXB, TY, LEFT = WX * 512, WY * RT, 128
ROWB = LEFT + TXB + 128
ROWS = TY + 4
STAGES = 8


def wrap(r, n):
    return r + n if r < 0 else (r - n if r >= n else r)


def update(s, c, alpha, T):
    return (c + (alpha * (s - (T(6) * c)).astype(T)).astype(T)).astype(T)


def reference_step(a, alpha, wrap_z=True):
    """One sweep, offsets (0,0,-1),(0,-1,0),(-1,0,0),(1,0,0),(0,1,0),(0,0,1); array indexed [x, y, z]."""
    T = a.dtype.type
    s = np.roll(a, 1, 2)
    s = (s + np.roll(a, 1, 1)).astype(T)
    s = (s + np.roll(a, 1, 0)).astype(T)
    s = (s + np.roll(a, -1, 0)).astype(T)
    s = (s + np.roll(a, -1, 1)).astype(T)
    s = (s + np.roll(a, -1, 2)).astype(T)
    return update(s, a, T(alpha), T)


def launch_cfg(X, Y, zn, es, ctas):
    ntx = (X * es + TXB - 1) // TXB
    best = (1e300, TY, 1)
    ty = TY
    while ty >= TY // 2:
        nty = (Y + ty - 1) // ty
        for nz in range(1, min(64, max(1, zn // 8)) + 1):
            tasks = ntx * nty * nz
            waves = (tasks + ctas - 1) // ctas
            cost = waves * (ty + 4.0) * (zn / nz + 4.0)
            if cost < best[0] * 0.999:
                best = (cost, ty, nz)
        ty -= RT
    return ntx, best[1], best[2]


def model(src, alpha, z_lo, zn, wrap_z, ctas=3, force_nz=None, force_ty=None):
    T = src.dtype.type
    es = src.dtype.itemsize
    VX = 16 // es
    X, Y, Z = src.shape
    Xb = X * es
    flat = np.ascontiguousarray(src.transpose(2, 1, 0)).view(np.uint8).reshape(Z, Y, Xb)   # [z][y][bytes of the row]
    dst = np.full((X, Y, Z), np.nan, dtype=src.dtype)
    ntx, ty, nzruns = launch_cfg(X, Y, zn, es, ctas)
    if force_nz:
        nzruns = force_nz
    if force_ty:
        ty = force_ty
    nty = (Y + ty - 1) // ty
    ntiles = ntx * nty
    ntasks = ntiles * nzruns
    grid = min(ctas, ntasks)
    alpha = T(alpha)
    lanes = np.arange(32)
    for block in range(grid):
        ring = np.full((STAGES, ROWS, ROWB), 0xFF, dtype=np.uint8)  # all-ones bytes = NaN
        k = 0
        for task in range(block, ntasks, grid):
            tile, zrun = task % ntiles, task // ntiles
            x0b, y0 = (tile % ntx) * TXB, (tile // ntx) * ty
            wbytes = min(TXB, Xb - x0b)
            z0 = z_lo + zn * zrun // nzruns
            z1 = z_lo + zn * (zrun + 1) // nzruns
            nsrc = z1 - z0 + 4
            # consumer state per warp
            st = {}
            for w in range(WARPS):
                st[w] = dict(c1=np.zeros((RT + 2, VX, 32), T), q1=np.zeros((RT + 2, VX, 32), T), c2=np.zeros((RT, VX, 32), T),
                             q2=np.zeros((RT, VX, 32), T), c1e=np.zeros((RT, 32), T), q1e=np.zeros((RT, 32), T))
            for i in range(nsrc):
                slot = k % STAGES
                k += 1
                # ---- producer ----
                ring[slot] = 0xFF
                l_in, r_in = x0b > 0, x0b + wbytes < Xb
                mstart = x0b - (16 if l_in else 0)
                mlen = wbytes + (16 if l_in else 0) + (16 if r_in else 0)
                mdst = LEFT - (16 if l_in else 0)
                zl = z0 - 2 + i
                if wrap_z:
                    zl = wrap(zl, Z)
                assert 0 <= zl < Z
                for lane in range(32):
                    if lane < ty + 4:
                        y = y0 - 2 + lane
                        if y <= Y + 1:
                            g = flat[zl, wrap(y, Y)]
                            ring[slot, lane, mdst:mdst + mlen] = g[mstart:mstart + mlen]
                            if not l_in:
                                ring[slot, lane, LEFT - 16:LEFT] = g[Xb - 16:Xb]
                            if not r_in:
                                ring[slot, lane, LEFT + wbytes:LEFT + wbytes + 16] = g[0:16]
                # ---- consumers ----
                for w in range(WARPS):
                    S = st[w]
                    wx, wy = w % WX, w // WX
                    xtb = (wx * 32 + lanes) * 16
                    ry0 = wy * RT
                    xact = xtb < wbytes
                    gx = (x0b + xtb) // es

                    def lds_vec(row):      # [VX, 32]
                        out = np.empty((VX, 32), T)
                        for l in range(32):
                            o = LEFT + xtb[l]
                            out[:, l] = ring[slot, row, o:o + 16].view(T)
                        return out

                    def lds(row, off):     # per-lane scalar at byte offset `off[l]` relative to the thread's 16 bytes
                        out = np.empty(32, T)
                        for l in range(32):
                            o = LEFT + xtb[l] + off[l]
                            out[l] = ring[slot, row, o:o + es].view(T)[0]
                        return out

                    rowv = [lds_vec(ry0 + q) for q in range(RT + 4)]
                    mid = np.empty((RT + 2, VX, 32), T)
                    with np.errstate(invalid="ignore", over="ignore"):
                        for j in range(RT + 2):
                            row = ry0 + j + 1
                            l_ = np.concatenate([rowv[j + 1][VX - 1][:1], rowv[j + 1][VX - 1][:-1]])
                            r_ = np.concatenate([rowv[j + 1][0][1:], rowv[j + 1][0][-1:]])
                            l_[0] = lds(row, np.full(32, -es))[0]
                            r_[31] = lds(row, np.full(32, 16))[31]
                            for v in range(VX):
                                c = rowv[j + 1][v]
                                cc = S["c1"][j, v]
                                mid[j, v] = update((S["q1"][j, v] + c).astype(T), cc, alpha, T)
                                xm = l_ if v == 0 else rowv[j + 1][v - 1]
                                xp = r_ if v == VX - 1 else rowv[j + 1][v + 1]
                                a = (cc + rowv[j][v]).astype(T)
                                a = (a + xm).astype(T)
                                a = (a + xp).astype(T)
                                a = (a + rowv[j + 2][v]).astype(T)
                                S["q1"][j, v] = a
                                S["c1"][j, v] = c.copy()
                        mide = np.zeros((RT, 32), T)
                        xe = np.where(lanes == 0, -es, 16)
                        for r in range(RT):
                            row = ry0 + r + 2
                            c = lds(row, xe)
                            cc = S["c1e"][r]
                            m = update((S["q1e"][r] + c).astype(T), cc, alpha, T)
                            a = (cc + lds(row - 1, xe)).astype(T)
                            a = (a + lds(row, xe - es)).astype(T)
                            a = (a + lds(row, xe + es)).astype(T)
                            a = (a + lds(row + 1, xe)).astype(T)
                            end = (lanes == 0) | (lanes == 31)
                            mide[r] = np.where(end, m, 0)
                            S["q1e"][r] = np.where(end, a, S["q1e"][r])
                            S["c1e"][r] = np.where(end, c, S["c1e"][r])
                        zo = z0 - 4 + i
                        for r in range(RT):
                            l_ = np.concatenate([mid[r + 1, VX - 1][:1], mid[r + 1, VX - 1][:-1]])
                            r_ = np.concatenate([mid[r + 1, 0][1:], mid[r + 1, 0][-1:]])
                            l_[0] = mide[r][0]
                            r_[31] = mide[r][31]
                            out = np.empty((VX, 32), T)
                            for v in range(VX):
                                c = mid[r + 1, v]
                                cc = S["c2"][r, v]
                                out[v] = update((S["q2"][r, v] + c).astype(T), cc, alpha, T)
                                xm = l_ if v == 0 else mid[r + 1, v - 1]
                                xp = r_ if v == VX - 1 else mid[r + 1, v + 1]
                                a = (cc + mid[r, v]).astype(T)
                                a = (a + xm).astype(T)
                                a = (a + xp).astype(T)
                                a = (a + mid[r + 2, v]).astype(T)
                                S["q2"][r, v] = a
                                S["c2"][r, v] = c.copy()
                            if i >= 4 and y0 + ry0 + r < Y and ry0 + r < ty:
                                for l in range(32):
                                    if xact[l]:
                                        assert np.isnan(dst[gx[l], y0 + ry0 + r, zo]), "cell stored twice"
                                        dst[gx[l]:gx[l] + VX, y0 + ry0 + r, zo] = out[:, l]
    return dst, (ntx, nty, ty, nzruns)


def check(shape, dtype, z_lo=0, zn=None, wrap_z=True, **kw):
    rng = np.random.default_rng(sum(shape))
    a = (rng.random(shape) - 0.3).astype(dtype)
    zn = shape[2] if zn is None else zn
    want = reference_step(reference_step(a, 0.1), 0.1)
    got, cfg = model(a, 0.1, z_lo, zn, wrap_z, **kw)
    u = {4: np.uint32, 8: np.uint64}[a.dtype.itemsize]
    region = slice(z_lo, z_lo + zn)
    ok = np.array_equal(got[:, :, region].view(u), want[:, :, region].view(u))
    outside = np.isnan(got[:, :, :z_lo]).all() and np.isnan(got[:, :, z_lo + zn:]).all()
    print(f"{shape} {np.dtype(dtype).name} z[{z_lo},{z_lo + zn}) wrap_z={wrap_z} cfg(ntx,nty,ty,nz)={cfg} {kw}: "
          f"{'OK' if ok and outside else 'MISMATCH'}")
    if not ok:
        bad = np.argwhere(got[:, :, region].view(u) != want[:, :, region].view(u))
        print("   first bad cells:", bad[:8].tolist(), "of", len(bad))
    return ok and outside


def main():
    ok = True
    ok &= check((64, 20, 9), np.float32)                       # one narrow tile, second x-warp idle
    ok &= check((160, 17, 8), np.float32)                      # second x-warp partly active, ragged y
    ok &= check((300, 16, 8), np.float32, ctas=2)              # two tiles in x (second ragged), several tasks per CTA
    ok &= check((64, 30, 20), np.float32, ctas=4, force_nz=2)  # z-runs
    ok &= check((40, 15, 12), np.float64)                      # Float64: two cells per thread
    ok &= check((128, 14, 16), np.float32, z_lo=3, zn=9, wrap_z=False)   # interior region (slab sweep)
    ok &= check((128, 14, 16), np.float32, z_lo=0, zn=5)       # region touching the wrap seam
    ok &= check((256, 9, 6), np.float32, force_ty=8)
    print("model matches two reference sweeps" if ok else "MODEL MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
