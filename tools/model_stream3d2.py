"""CPU model of stream3d2_kernel (csrc/stream3d2.cu): the kernel's index arithmetic transliterated thread for thread
(tasks, ring slots, producer row / halo copies, warp spans, end-lane cells, the two register levels), lanes vectorised
with NumPy. Uncopied shared memory is NaN, so any read the producer did not cover poisons the result. Compared with
two plain diffusion sweeps in the reference's fold order. Run on a box without a GPU to check the kernel's logic:

    python tools/model_stream3d2.py
"""
import sys

import numpy as np

WX, WY, RT = 2, 8, 2
WARPS = WX * WY
TXB, TY, LEFT = WX * 512, (WY - 1) * RT, 128
ROWB = LEFT + TXB + 128
ROWS = TY + 4
MROWS = TY + 2
STAGES = 6


def wrap(r, n):
    return r + n if r < 0 else (r - n if r >= n else r)


def update(s, c, alpha, T):
    return (c + (alpha * (s - (T(6) * c)).astype(T)).astype(T)).astype(T)


def reference_step(a, alpha, bcs=("wrap", "wrap", "wrap"), pad=0.0):
    """One sweep, offsets (0,0,-1),(0,-1,0),(-1,0,0),(1,0,0),(0,1,0),(0,0,1); array indexed [x, y, z]; per-axis Wrap or
    Remove(pad) (out-of-bounds neighbours read pad)."""
    T = a.dtype.type
    g = a
    for ax, bc in enumerate(bcs):
        w = [(0, 0)] * 3
        w[ax] = (1, 1)
        g = np.pad(g, w, mode="wrap") if bc == "wrap" else np.pad(g, w, mode="constant", constant_values=T(pad))
    X, Y, Z = a.shape
    c = lambda dx, dy, dz: g[1 + dx:1 + dx + X, 1 + dy:1 + dy + Y, 1 + dz:1 + dz + Z]  # noqa: E731
    s = c(0, 0, -1)
    for d in ((0, -1, 0), (-1, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)):
        s = (s + c(*d)).astype(T)
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


def model(src, alpha, z_lo, zn, wrap_z, ctas=3, force_nz=None, force_ty=None, bcs=("wrap", "wrap", None), pad=0.0):
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
    # the PAD variant of the kernel (Remove axes): bcs[2] None = "as wrap_z says"
    padx, pady = bcs[0] == "remove", bcs[1] == "remove"
    padz = bcs[2] == "remove"
    pad = T(pad)
    for block in range(grid):
        ring = np.full((STAGES, ROWS, ROWB), 0xFF, dtype=np.uint8)  # all-ones bytes = NaN
        mbuf = np.full((2, MROWS, ROWB), 0xFF, dtype=np.uint8)
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
                st[w] = dict(c1=np.zeros((RT, VX, 32), T), q1=np.zeros((RT, VX, 32), T), c2=np.zeros((RT, VX, 32), T),
                             q2=np.zeros((RT, VX, 32), T), c1e=np.zeros((RT, 32), T), q1e=np.zeros((RT, 32), T))
            for i in range(nsrc):
                slot = k % STAGES
                par = k & 1
                k += 1
                # ---- producers (lane j of producer w copies row 2j+w) ----
                ring[slot] = 0xFF
                l_in, r_in = x0b > 0, x0b + wbytes < Xb
                mstart = x0b - (16 if l_in else 0)
                mlen = wbytes + (16 if l_in else 0) + (16 if r_in else 0)
                mdst = LEFT - (16 if l_in else 0)
                zl = z0 - 2 + i
                if wrap_z and not padz:
                    zl = wrap(zl, Z)
                zin = not (padz and (zl < 0 or zl >= Z))
                assert (0 <= zl < Z) or not zin
                copied = 0
                for pw in range(2):
                    for lane in range(32):
                        row = 2 * lane + pw
                        if row < ty + 4:
                            y = y0 - 2 + row
                            yrow = wrap(y, Y) if y <= Y + 1 else -1
                            if pady:
                                yrow = y if 0 <= y < Y else -1
                            if yrow >= 0:
                                copied += 1
                                if not zin:
                                    continue
                                g = flat[zl, yrow]
                                ring[slot, row, mdst:mdst + mlen] = g[mstart:mstart + mlen]
                                if not l_in and not padx:
                                    ring[slot, row, LEFT - 16:LEFT] = g[Xb - 16:Xb]
                                if not r_in and not padx:
                                    ring[slot, row, LEFT + wbytes:LEFT + wbytes + 16] = g[0:16]
                nrows = min(ty + 4, Y + 4 - y0)
                if pady:
                    nrows = min(Y - 1, y0 + ty + 1) - max(0, y0 - 2) + 1
                assert copied == nrows, "expect_tx row count"
                sz = z0 - 2 + i
                zs_oob = padz and (sz < 0 or sz >= Z)
                zm_oob = padz and (sz - 1 < 0 or sz - 1 >= Z)
                mbuf[par] = 0xFF

                def geom(w):
                    wx, wy = w % WX, w // WX
                    return wx, wy, (wx * 32 + lanes) * 16, wy * RT

                def lds_vec(buf, row, xtb):      # [VX, 32]
                    out = np.empty((VX, 32), T)
                    for l in range(32):
                        o = LEFT + xtb[l]
                        out[:, l] = buf[row, o:o + 16].view(T)
                    return out

                def lds(buf, row, xtb, off):     # per-lane scalar at byte offset off[l] relative to the thread's 16 bytes
                    out = np.empty(32, T)
                    for l in range(32):
                        o = LEFT + xtb[l] + off[l]
                        out[l] = buf[row, o:o + es].view(T)[0]
                    return out

                def plane(cprev, part, c, ym, yp, l_, r_):
                    done = np.empty((VX, 32), T)
                    for v in range(VX):
                        cc = cprev[v].copy()
                        done[v] = update((part[v] + c[v]).astype(T), cc, alpha, T)
                        xm = l_ if v == 0 else c[v - 1]
                        xp = r_ if v == VX - 1 else c[v + 1]
                        a = (cc + ym[v]).astype(T)
                        a = (a + xm).astype(T)
                        a = (a + xp).astype(T)
                        a = (a + yp[v]).astype(T)
                        part[v] = a
                        cprev[v] = c[v].copy()
                    return done

                mids = {}
                with np.errstate(invalid="ignore", over="ignore"):
                    # ---- level 1 (all warps), then the named barrier ----
                    for w in range(WARPS):
                        S = st[w]
                        wx, wy, xtb, r1 = geom(w)
                        src_ = ring[slot]
                        rowv = [lds_vec(src_, r1 + q, xtb) for q in range(RT + 2)]
                        gx_ = (x0b + xtb) // es
                        xoob = padx & (gx_ >= X)
                        edge_l, edge_r = padx & (gx_ == 0), padx & (gx_ + VX == X)
                        srow_oob = [pady and not (0 <= y0 + r1 - 2 + q < Y) for q in range(RT + 2)]
                        for q in range(RT + 2):
                            if zs_oob or srow_oob[q]:
                                rowv[q] = np.full((VX, 32), pad, T)
                        mid = np.empty((RT, VX, 32), T)
                        for j in range(RT):
                            row = r1 + j + 1
                            l_ = np.concatenate([rowv[j + 1][VX - 1][:1], rowv[j + 1][VX - 1][:-1]])
                            r_ = np.concatenate([rowv[j + 1][0][1:], rowv[j + 1][0][-1:]])
                            l_[0] = lds(src_, row, xtb, np.full(32, -es))[0]
                            r_[31] = lds(src_, row, xtb, np.full(32, 16))[31]
                            if zs_oob or srow_oob[j + 1]:
                                l_[:] = pad
                                r_[:] = pad
                            l_ = np.where(edge_l, pad, l_)
                            r_ = np.where(edge_r, pad, r_)
                            mid[j] = plane(S["c1"][j], S["q1"][j], rowv[j + 1], rowv[j], rowv[j + 2], l_, r_)
                            mrow_oob = pady and not (0 <= y0 + r1 - 1 + j < Y)
                            if zm_oob or mrow_oob:
                                mid[j][:] = pad
                            mid[j][:, xoob] = pad
                        rim = ((lanes == 0) & (wx == 0)) | ((lanes == 31) & (wx == WX - 1))
                        xe = np.where(lanes == 0, -es, 16)
                        for j in range(RT):
                            row = r1 + j + 1
                            yr = y0 - 1 + (r1 + j)          # logical row of intermediate row r1 + j
                            x_oob = padx and ((x0b == 0) if wx == 0 else (x0b + wbytes == Xb))
                            row_oob = pady and not (0 <= yr < Y)
                            up_oob, dn_oob = pady and not (0 <= yr - 1 < Y), pady and not (0 <= yr + 1 < Y)
                            c = lds(src_, row, xtb, xe)
                            yu, yd = lds(src_, row - 1, xtb, xe), lds(src_, row + 1, xtb, xe)
                            if zs_oob or row_oob:
                                c = np.full(32, pad, T)
                            if zs_oob or up_oob:
                                yu = np.full(32, pad, T)
                            if zs_oob or dn_oob:
                                yd = np.full(32, pad, T)
                            cc = S["c1e"][j]
                            m = update((S["q1e"][j] + c).astype(T), cc, alpha, T)
                            if x_oob or row_oob or zm_oob:
                                m = np.full(32, pad, T)
                            a = (cc + yu).astype(T)
                            a = (a + lds(src_, row, xtb, xe - es)).astype(T)
                            a = (a + lds(src_, row, xtb, xe + es)).astype(T)
                            a = (a + yd).astype(T)
                            S["q1e"][j] = np.where(rim, a, S["q1e"][j])
                            S["c1e"][j] = np.where(rim, c, S["c1e"][j])
                            for l in range(32):
                                o = LEFT + xtb[l]
                                mbuf[par, r1 + j, o:o + 16] = np.ascontiguousarray(mid[j][:, l]).view(np.uint8)
                            for l in range(32):
                                if rim[l]:
                                    o = LEFT + xtb[l] + xe[l]
                                    mbuf[par, r1 + j, o:o + es] = np.array([m[l]], T).view(np.uint8)
                        mids[w] = mid
                    # ---- level 2 ----
                    zo = z0 - 4 + i
                    for w in range(WARPS):
                        S = st[w]
                        wx, wy, xtb, r1 = geom(w)
                        if wy >= WY - 1:
                            continue
                        xact = xtb < wbytes
                        gx = (x0b + xtb) // es
                        M = mbuf[par]
                        m2 = lds_vec(M, r1 + 2, xtb)
                        m3 = lds_vec(M, r1 + 3, xtb)
                        mid = mids[w]
                        for r in range(RT):
                            ym = mid[0] if r == 0 else mid[1]
                            cc = mid[1] if r == 0 else m2
                            yp = m2 if r == 0 else m3
                            row = r1 + r + 1
                            l_ = np.concatenate([cc[VX - 1][:1], cc[VX - 1][:-1]])
                            r_ = np.concatenate([cc[0][1:], cc[0][-1:]])
                            l_[0] = lds(M, row, xtb, np.full(32, -es))[0]
                            r_[31] = lds(M, row, xtb, np.full(32, 16))[31]
                            out = plane(S["c2"][r], S["q2"][r], cc, ym, yp, l_, r_)
                            if i >= 4 and y0 + r1 + r < Y and r1 + r < ty:
                                for l in range(32):
                                    if xact[l]:
                                        assert np.isnan(dst[gx[l], y0 + r1 + r, zo]), "cell stored twice"
                                        dst[gx[l]:gx[l] + VX, y0 + r1 + r, zo] = out[:, l]
    return dst, (ntx, nty, ty, nzruns)


def check(shape, dtype, z_lo=0, zn=None, wrap_z=True, bcs=("wrap", "wrap", None), pad=0.0, **kw):
    rng = np.random.default_rng(sum(shape))
    a = (rng.random(shape) - 0.3).astype(dtype)
    zn = shape[2] if zn is None else zn
    ref_bcs = (bcs[0], bcs[1], bcs[2] or "wrap")
    want = reference_step(reference_step(a, 0.1, ref_bcs, pad), 0.1, ref_bcs, pad)
    got, cfg = model(a, 0.1, z_lo, zn, wrap_z, bcs=bcs, pad=pad, **kw)
    u = {4: np.uint32, 8: np.uint64}[a.dtype.itemsize]
    region = slice(z_lo, z_lo + zn)
    ok = np.array_equal(got[:, :, region].view(u), want[:, :, region].view(u))
    outside = np.isnan(got[:, :, :z_lo]).all() and np.isnan(got[:, :, z_lo + zn:]).all()
    print(f"{shape} {np.dtype(dtype).name} z[{z_lo},{z_lo + zn}) wrap_z={wrap_z} bcs={bcs} cfg(ntx,nty,ty,nz)={cfg} {kw}: "
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
    # the PAD variant (Remove axes, SB200_D2_REMOVE=1): padval outside the array at BOTH time levels
    R, W = "remove", "wrap"
    ok &= check((64, 20, 9), np.float32, bcs=(R, R, R), pad=0.25)
    ok &= check((160, 17, 8), np.float32, bcs=(R, W, R), pad=-1.5)
    ok &= check((300, 16, 8), np.float32, bcs=(W, R, None), pad=0.5, ctas=2)
    ok &= check((300, 30, 8), np.float32, bcs=(R, R, None), pad=0.5, ctas=2, force_ty=14)
    ok &= check((512, 15, 12), np.float32, bcs=(R, R, R), pad=2.0, force_nz=2, ctas=5)
    ok &= check((40, 15, 12), np.float64, bcs=(R, R, R), pad=0.125)
    ok &= check((128, 14, 16), np.float32, z_lo=3, zn=9, wrap_z=False, bcs=(R, R, None), pad=1.0)   # slab sweep, Remove on x and y
    print("model matches two reference sweeps" if ok else "MODEL MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
