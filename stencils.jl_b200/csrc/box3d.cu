// box3d.cu — 3-D box stencils at compile time: Window(1,3) (27 taps) and Moore(1,3) (26 taps) x sum / mean / minimum /
// maximum for Float32 / Float64, 2.5-D streaming like stream3d.cu.
//
// Replaces gatherstencil_kernel! (src/gatherstencil.jl:105-109) for the N-D box shapes (src/stencils/window.jl:4-8,
// src/stencils/moore.jl:5-18; offsets = the box (-1:1)^3 with axis 0 fastest, Moore without the centre). The run-time
// table kernel (gather_stream3d.cu) pays one shared-memory load per tap (27 LDS per cell: 0.21 of the HBM roofline on
// Window(1,3) 768^3); here a thread owns 16 bytes of x by B3_RT rows, reads every source row ONCE per plane as a 128-bit
// load (x neighbours by warp shuffle) and keeps the folds in registers.
//
// sum / mean: the reference's left fold runs over the offsets in order — plane z-1 (rows y-1, y, y+1; x-1, x, x+1 inside a
// row), then plane z, then plane z+1 — every addition rounded separately, so partial sums cannot be shared between cells.
// But the chain of an output cell consumes the planes in arrival order, so a thread keeps TWO running chains per cell (the
// output whose z-1 plane has been folded, and the one whose z-1 and z planes have) and, when plane p arrives, starts the
// chain of output p+1 (8 adds), continues that of output p (9 adds; 8 for Moore, which skips the centre) and finishes
// that of output p-1 (9 adds) — 26 (25) dependent additions per cell, bit-identical to the Julia fold.
// minimum / maximum: Julia's max / min are associative and commutative (NaN wins, -0 < +0), so the box maximum is
// separable: row maxima (3-input FMNMX3), column maxima of those, then the maximum over three planes — 3 (Window) or 5
// (Moore) operations per cell.
// Algorithmic traffic: sizeof(T) read + sizeof(T) written per cell.
#include <algorithm>
#include "common.cuh"
#include "tma.cuh"

namespace sb {

constexpr int B3_WX = 2, B3_WY = 4;             // consumer warps across x and y
constexpr int B3_WARPS = B3_WX * B3_WY;
constexpr int B3_PRODUCERS = 2;                 // producer warps (rows of a stage dealt round-robin; stream3d.cu lesson 3)
constexpr int B3_THREADS = (B3_WARPS + B3_PRODUCERS) * 32;
constexpr int B3_TXB = B3_WX * 512;             // tile width in bytes
#ifndef SB200_B3_RT
#define SB200_B3_RT 4
#endif
#ifndef SB200_B3_BACKOFF_NS
#define SB200_B3_BACKOFF_NS 0      // producer back-off while the ring is full: measured r02g, 0 / 1000 / 4000 ns all 556-558 Gcell/s
#endif
#ifndef SB200_B3_UNROLL
#define SB200_B3_UNROLL 1   // unroll factor of the consumers' plane loop. Measured r02an, Window(1,3) mean 768^3: 2 spills at the 96-register cap
                            // (570 -> 497 Gcell/s) — unlike stream3d2_kernel, whose loop gained 3 % from the renaming. Stays 1.
#endif
constexpr int B3_UNROLL = SB200_B3_UNROLL;
#ifndef SB200_B3_PACKED
#define SB200_B3_PACKED 0          // Float32 sums with packed add.rn.f32x2 (two cells per issue slot). Measured r02h, Window(1,3) mean
                                   // 768^3: scalar 571 Gcell/s, packed 504 — assembling the (x-1, x) / (x+1, x+2) operand pairs costs more
                                   // moves than the 208 saved FADD issue slots (stream3d2's pairs are register-aligned, these are not).
                                   // 2 = mixed fold: the centre taps as packed adds, the x-neighbour taps as scalar adds on the halves of
                                   // the accumulator pair (no pair assembly; 418 FADD -> 290 FADD + 64 FADD2 per plane and thread).
                                   // Measured r02az: bit-exact (29 GPU parity tests), 572.5 / 566.9 against 570.8 / 562.4 Gcell/s: no gain,
                                   // the kernel is not bound by its FP issue slots alone. Stays 0.
#endif
constexpr int B3_RT = SB200_B3_RT;              // rows per thread
constexpr int B3_TY = B3_WY * B3_RT;            // tile height in rows
constexpr int B3_LEFT = 128;                    // margin (halo at its end): global and shared addresses agree mod 128
constexpr int B3_ROWB = B3_LEFT + B3_TXB + 128; // shared-memory row: margin | tile | margin
constexpr int B3_STAGE = (B3_TY + 2) * B3_ROWB;
constexpr int B3_STAGES = 4;
constexpr int B3_SMEM = 128 + B3_STAGES * B3_STAGE;
static_assert(B3_TY + 2 <= 32 * B3_PRODUCERS, "one producer lane per row");

template <typename T> struct B3Params {
    const T* src;
    T* dst;
    long long sp1, sp2, dp1, dp2;  // source / dest pitches (elements) of axes 1 and 2
    int X, Y, Z;                   // logical size
    int so1, so2, do0, do1, do2;   // ring / ghost offsets (source axis 0 is unpadded)
    int bc0, bc1, bc2;
    T pad;
    int z_lo, zn;                  // output planes [z_lo, z_lo + zn)
    int ntx, nty, nzruns, ty;
};

__device__ __forceinline__ long long b3_map(int r, int n, int off, int bc) {
    if (off > 0) return (long long)r + off;
    if (r >= 0 && r < n) return r;
    if (bc == SB200_WRAP) return r < 0 ? r + n : r - n;
    if (bc == SB200_REFLECT) return r < 0 ? -r : 2 * (n - 1) - r;
    return -1;
}

template <typename T> struct B3Vec;
template <> struct B3Vec<float> { using type = float4; };
template <> struct B3Vec<double> { using type = double2; };

template <typename T, bool ISMAX> __device__ __forceinline__ T b3_ext3(T a, T b, T c) { return ISMAX ? jl_max3(a, b, c) : jl_min3(a, b, c); }
template <typename T, bool ISMAX> __device__ __forceinline__ T b3_ext2(T a, T b) { return ISMAX ? jl_max(a, b) : jl_min(a, b); }

// The nine taps of one plane for the cell in column v of row r (rows r-1, r, r+1 of the window = rowv[r], rowv[r+1],
// rowv[r+2]), added to `acc` in offset order. FIRST: the chain starts with the first tap. SKIPC: Moore's middle plane.
template <typename T, int VX, int NR, bool FIRST, bool SKIPC>
__device__ __forceinline__ T b3_fold9(T acc, const T (&rowv)[NR][VX], const T (&xl)[NR], const T (&xr)[NR], int r, int v) {
#pragma unroll
    for (int dy = 0; dy < 3; dy++) {
        const T m = v == 0 ? xl[r + dy] : rowv[r + dy][v == 0 ? 0 : v - 1];
        const T c = rowv[r + dy][v];
        const T q = v == VX - 1 ? xr[r + dy] : rowv[r + dy][v == VX - 1 ? v : v + 1];
        if (FIRST && dy == 0) acc = m; else acc = add_rn(acc, m);
        if (!(SKIPC && dy == 1)) acc = add_rn(acc, c);
        acc = add_rn(acc, q);
    }
    return acc;
}

// Packed Float32 adds (SASS FADD2): each lane rounds like the scalar add, so the folds stay bit-exact.
__device__ __forceinline__ unsigned long long b3_pk(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void b3_upk(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long b3_add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// MOORE: centre excluded. RED: SB200_SUM / SB200_MEAN / SB200_MAX / SB200_MIN. PAD: some axis is Remove (out-of-bounds rows /
// planes / columns read padval); the other instantiation carries no padval code at all.
template <typename T, bool MOORE, int RED, bool PAD>
__global__ void __launch_bounds__(B3_THREADS, 2) box3d_kernel(const __grid_constant__ B3Params<T> p) {
    constexpr int VX = 16 / (int)sizeof(T);
    constexpr int NR = B3_RT + 2;
    constexpr bool EXT = RED == SB200_MAX || RED == SB200_MIN;
    constexpr bool ISMAX = RED == SB200_MAX;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + B3_STAGES;
    unsigned char* ring = smem + 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < B3_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], B3_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    const int ntiles = p.ntx * p.nty;
    const int ntasks = ntiles * p.nzruns;
    const int Xb = p.X * (int)sizeof(T);
    unsigned k = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int tile = task % ntiles, zrun = task / ntiles;
        const int x0b = (tile % p.ntx) * B3_TXB, y0 = (tile / p.ntx) * p.ty;
        const int wbytes = min(B3_TXB, Xb - x0b);
        const int z0 = p.z_lo + (int)((long long)p.zn * zrun / p.nzruns);
        const int z1 = p.z_lo + (int)((long long)p.zn * (zrun + 1) / p.nzruns);
        const int nsrc = z1 - z0 + 2;  // source planes z0-1 .. z1
        if (warp >= B3_WARPS) {
            // ---------------- producer warps: lane j of producer w copies row P*j+w of the plane (as in stream3d.cu) ----------------
            const int pw = warp - B3_WARPS;
            const bool l_in = x0b > 0, r_in = x0b + wbytes < Xb;
            const bool l_wrap = !l_in && p.bc0 == SB200_WRAP, r_wrap = !r_in && p.bc0 == SB200_WRAP;
            const int mstart = x0b - (l_in ? 16 : 0);
            const unsigned mlen = wbytes + (l_in ? 16 : 0) + (r_in ? 16 : 0);
            const int mdst = B3_LEFT - (l_in ? 16 : 0);
            const unsigned rowbytes = mlen + (l_wrap ? 16 : 0) + (r_wrap ? 16 : 0);
            long long yrow = -1, yany = -1;   // yany: row `lane` of the stage (every producer counts all rows for expect_tx)
            if (lane < p.ty + 2) {
                const int y = y0 - 1 + lane;
                if (y <= p.Y) yany = b3_map(y, p.Y, p.so1, p.bc1);
            }
            const unsigned nrows = __popc(__ballot_sync(0xffffffffu, yany >= 0));
            const int srow_i = B3_PRODUCERS * lane + pw;
            if (srow_i < p.ty + 2) {
                const int y = y0 - 1 + srow_i;
                if (y <= p.Y) yrow = b3_map(y, p.Y, p.so1, p.bc1);
            }
            for (int i = 0; i < nsrc; i++, k++) {
                const int slot = k % B3_STAGES;
                const long long zpl = b3_map(z0 - 1 + i, p.Z, p.so2, p.bc2);
                if (lane == 0) {
                    mbar_wait_producer(&empty[slot], ((k / B3_STAGES) & 1) ^ 1, SB200_B3_BACKOFF_NS);
                    if (pw == 0) mbar_arrive_expect_tx(&full[slot], zpl >= 0 ? nrows * rowbytes : 0u);
                }
                __syncwarp();
                if (zpl >= 0 && yrow >= 0) {
                    const unsigned char* g = reinterpret_cast<const unsigned char*>(p.src + zpl * p.sp2 + yrow * p.sp1);
                    unsigned char* srow = ring + slot * B3_STAGE + srow_i * B3_ROWB;
                    bulk_g2s(srow + mdst, g + mstart, mlen, &full[slot]);
                    if (l_wrap) bulk_g2s(srow + B3_LEFT - 16, g + Xb - 16, 16, &full[slot]);
                    if (r_wrap) bulk_g2s(srow + B3_LEFT + wbytes, g, 16, &full[slot]);
                }
            }
            continue;
        }
        // ---------------- consumers ----------------
        const int wx = warp % B3_WX, wy = warp / B3_WX;
        const int xtb = (wx * 32 + lane) * 16;         // byte offset inside the tile
        const int ry0 = wy * B3_RT;                    // first tile row of this thread
        const bool xact = xtb < wbytes;
        const int gx = (x0b + xtb) / (int)sizeof(T);
        const bool edge_l = xact && p.bc0 != SB200_WRAP && gx == 0;
        const bool edge_r = xact && p.bc0 != SB200_WRAP && gx + VX == p.X;
        const bool pad1 = PAD && p.so1 == 0 && p.bc1 == SB200_REMOVE;   // OOB rows read padval
        const bool pad2 = PAD && p.so2 == 0 && p.bc2 == SB200_REMOVE;   // OOB planes read padval
        // sums: a1 = chain of the output whose z-1 plane is folded, a2 = z-1 and z planes folded
        // extrema: a1 = box-plane extremum M of the previous plane, a2 = of the plane before; a3 (Moore) = ring of the previous plane
        T a1[B3_RT][VX], a2[B3_RT][VX], a3[MOORE && EXT ? B3_RT : 1][VX];
#pragma unroll
        for (int r = 0; r < B3_RT; r++)
#pragma unroll
            for (int v = 0; v < VX; v++) { a1[r][v] = T(0); a2[r][v] = T(0); if constexpr (MOORE && EXT) a3[r][v] = T(0); }
        unsigned long long a1p[B3_RT][2], a2p[B3_RT][2];   // the same two chains as packed cell pairs (Float32 sums)
#pragma unroll
        for (int r = 0; r < B3_RT; r++) { a1p[r][0] = a1p[r][1] = 0ull; a2p[r][0] = a2p[r][1] = 0ull; }
        T* __restrict__ dbase = p.dst + (long long)(y0 + ry0 + p.do1) * p.dp1 + p.do0 + gx;
#pragma unroll B3_UNROLL
        for (int i = 0; i < nsrc; i++, k++) {
            const int slot = k % B3_STAGES;
            const int z = z0 - 1 + i;                  // logical plane held by this stage
            const bool zpad = pad2 && (z < 0 || z >= p.Z);
            mbar_wait(&full[slot], (k / B3_STAGES) & 1);
            const unsigned char* sb_ = ring + slot * B3_STAGE + B3_LEFT + xtb;
            // rows ry0-1 .. ry0+RT of the tile (shared-memory row index = tile row + 1), with their x neighbours
            T rowv[NR][VX];
            T xl[NR], xr[NR];
#pragma unroll
            for (int r = 0; r < NR; r++) {
                const unsigned char* t = sb_ + (ry0 + r) * B3_ROWB;
                // (a row or plane outside a Remove axis was never copied: the stale shared-memory cells are read and replaced)
                const typename B3Vec<T>::type q = *reinterpret_cast<const typename B3Vec<T>::type*>(t);
                if constexpr (VX == 4) { rowv[r][0] = q.x; rowv[r][1] = q.y; rowv[r][2] = q.z; rowv[r][3] = q.w; }
                else { rowv[r][0] = q.x; rowv[r][1] = q.y; }
                // x neighbours across the 16-byte vectors come from the adjacent lanes; only the warp's end lanes read the
                // halo cells (predicated loads: no divergent branch)
                T l_ = __shfl_up_sync(0xffffffffu, rowv[r][VX - 1], 1);
                T r_ = __shfl_down_sync(0xffffffffu, rowv[r][0], 1);
                lds_if(l_, t - sizeof(T), lane == 0);
                lds_if(r_, t + 16, lane == 31);
                if (edge_l) l_ = p.bc0 == SB200_REFLECT ? rowv[r][1] : p.pad;
                if (edge_r) r_ = p.bc0 == SB200_REFLECT ? rowv[r][VX - 2] : p.pad;
                if constexpr (PAD) {
                    const int y = y0 + ry0 - 1 + r;
                    const bool ypad = zpad || (pad1 && (y < 0 || y >= p.Y));
#pragma unroll
                    for (int v = 0; v < VX; v++) rowv[r][v] = ypad ? p.pad : rowv[r][v];
                    l_ = ypad ? p.pad : l_;
                    r_ = ypad ? p.pad : r_;
                }
                xl[r] = l_;
                xr[r] = r_;
            }
            const int zo = z - 1;  // output plane completed by this stage
            const bool store = i >= 2;
            if constexpr (EXT) {
                // row extrema of every window row (the row itself: x-1, x, x+1), then box-plane extrema per output row
                T rm[NR][VX];
#pragma unroll
                for (int r = 0; r < NR; r++)
#pragma unroll
                    for (int v = 0; v < VX; v++) {
                        const T m = v == 0 ? xl[r] : rowv[r][v == 0 ? 0 : v - 1];
                        const T q = v == VX - 1 ? xr[r] : rowv[r][v == VX - 1 ? v : v + 1];
                        rm[r][v] = b3_ext3<T, ISMAX>(m, rowv[r][v], q);
                    }
#pragma unroll
                for (int r = 0; r < B3_RT; r++) {
                    T out[VX];
#pragma unroll
                    for (int v = 0; v < VX; v++) {
                        const T M = b3_ext3<T, ISMAX>(rm[r][v], rm[r + 1][v], rm[r + 2][v]);
                        if constexpr (MOORE) {
                            // output z-1 = ext(M(z-2), ring(z-1), M(z)); ring = the 8 in-plane neighbours without the centre
                            out[v] = b3_ext3<T, ISMAX>(a2[r][v], a3[r][v], M);
                            const T m = v == 0 ? xl[r + 1] : rowv[r + 1][v == 0 ? 0 : v - 1];
                            const T q = v == VX - 1 ? xr[r + 1] : rowv[r + 1][v == VX - 1 ? v : v + 1];
                            a3[r][v] = b3_ext3<T, ISMAX>(rm[r][v], rm[r + 2][v], b3_ext2<T, ISMAX>(m, q));
                        } else {
                            out[v] = b3_ext3<T, ISMAX>(a2[r][v], a1[r][v], M);
                        }
                        a2[r][v] = a1[r][v];
                        a1[r][v] = M;
                    }
                    const int y = y0 + ry0 + r;
                    if (store && xact && y < p.Y && ry0 + r < p.ty) {
                        T* d = dbase + (long long)(zo + p.do2) * p.dp2 + (long long)r * p.dp1;
                        if constexpr (VX == 4) *reinterpret_cast<float4*>(d) = make_float4(out[0], out[1], out[2], out[3]);
                        else *reinterpret_cast<double2*>(d) = make_double2(out[0], out[1]);
                    }
                }
            } else if constexpr (SB200_B3_PACKED && sizeof(T) == 4) {
                constexpr int L = MOORE ? 26 : 27;
                // Window rows are streamed: row w is turned into its cell pairs (0,1) / (2,3) — left neighbours, centres, right
                // neighbours — and folded at once into every chain that uses it (output rows w, w-1, w-2 x the three time levels),
                // so only one row of pairs is live at a time. Every chain still sees its taps in offset order.
                unsigned long long fin[B3_RT][2], n2[B3_RT][2], n1[B3_RT][2];
#pragma unroll
                for (int r = 0; r < B3_RT; r++)
#pragma unroll
                    for (int h = 0; h < 2; h++) { fin[r][h] = a2p[r][h]; n2[r][h] = a1p[r][h]; n1[r][h] = 0ull; }
#pragma unroll
                for (int w = 0; w < NR; w++) {
                    const unsigned long long pc[2] = {b3_pk(rowv[w][0], rowv[w][1]), b3_pk(rowv[w][2], rowv[w][3])};
#if SB200_B3_PACKED == 2
                    // mixed form: only the centre taps (register-aligned pairs) are packed adds, the x-neighbour taps are scalar
                    // adds on the halves of the accumulator pair (no pair assembly)
                    const float m0[2] = {xl[w], rowv[w][1]}, m1[2] = {rowv[w][0], rowv[w][2]};   // left neighbours of a pair's cells
                    const float q0[2] = {rowv[w][1], rowv[w][3]}, q1[2] = {rowv[w][2], xr[w]};   // right neighbours
                    auto adds = [](unsigned long long a, float x0, float x1) {
                        float a0, a1;
                        b3_upk(a, a0, a1);
                        return b3_pk(__fadd_rn(a0, x0), __fadd_rn(a1, x1));
                    };
#pragma unroll
                    for (int dy = 0; dy < 3; dy++) {
                        const int r = w - dy;
                        if (r < 0 || r >= B3_RT) continue;
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            fin[r][h] = adds(b3_add2(adds(fin[r][h], m0[h], m1[h]), pc[h]), q0[h], q1[h]);
                            unsigned long long b = adds(n2[r][h], m0[h], m1[h]);
                            if (!(MOORE && dy == 1)) b = b3_add2(b, pc[h]);
                            n2[r][h] = adds(b, q0[h], q1[h]);
                            if (dy == 0) {
                                const float c0 = rowv[w][2 * h], c1 = rowv[w][2 * h + 1];
                                n1[r][h] = b3_pk(__fadd_rn(__fadd_rn(m0[h], c0), q0[h]), __fadd_rn(__fadd_rn(m1[h], c1), q1[h]));
                            } else {
                                n1[r][h] = adds(b3_add2(adds(n1[r][h], m0[h], m1[h]), pc[h]), q0[h], q1[h]);
                            }
                        }
                    }
                    continue;
#endif
                    const unsigned long long mid = b3_pk(rowv[w][1], rowv[w][2]);
                    const unsigned long long pm[2] = {b3_pk(xl[w], rowv[w][0]), mid};
                    const unsigned long long pq[2] = {mid, b3_pk(rowv[w][3], xr[w])};
#pragma unroll
                    for (int dy = 0; dy < 3; dy++) {
                        const int r = w - dy;
                        if (r < 0 || r >= B3_RT) continue;
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            fin[r][h] = b3_add2(b3_add2(b3_add2(fin[r][h], pm[h]), pc[h]), pq[h]);           // plane z+1 of output z-1
                            unsigned long long b = b3_add2(n2[r][h], pm[h]);                                  // plane z of output z
                            if (!(MOORE && dy == 1)) b = b3_add2(b, pc[h]);
                            n2[r][h] = b3_add2(b, pq[h]);
                            const unsigned long long c = dy == 0 ? pm[h] : b3_add2(n1[r][h], pm[h]);        // plane z-1 of output z+1
                            n1[r][h] = b3_add2(b3_add2(c, pc[h]), pq[h]);
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < B3_RT; r++) {
                    float out[4];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        float f0, f1;
                        b3_upk(fin[r][h], f0, f1);
                        out[2 * h] = RED == SB200_MEAN ? div_rn(f0, (float)L) : f0;
                        out[2 * h + 1] = RED == SB200_MEAN ? div_rn(f1, (float)L) : f1;
                        a2p[r][h] = n2[r][h];
                        a1p[r][h] = n1[r][h];
                    }
                    const int y = y0 + ry0 + r;
                    if (store && xact && y < p.Y && ry0 + r < p.ty) {
                        T* d = dbase + (long long)(zo + p.do2) * p.dp2 + (long long)r * p.dp1;
                        *reinterpret_cast<float4*>(d) = make_float4(out[0], out[1], out[2], out[3]);
                    }
                }
            } else {
                constexpr int L = MOORE ? 26 : 27;
#pragma unroll
                for (int r = 0; r < B3_RT; r++) {
                    T out[VX];
#pragma unroll
                    for (int v = 0; v < VX; v++) {
                        const T fin = b3_fold9<T, VX, NR, false, false>(a2[r][v], rowv, xl, xr, r, v);   // plane z+1 of output z-1
                        out[v] = RED == SB200_MEAN ? div_rn(fin, (T)L) : fin;
                        a2[r][v] = b3_fold9<T, VX, NR, false, MOORE>(a1[r][v], rowv, xl, xr, r, v);      // plane z of output z
                        a1[r][v] = b3_fold9<T, VX, NR, true, false>(T(0), rowv, xl, xr, r, v);            // plane z-1 of output z+1
                    }
                    const int y = y0 + ry0 + r;
                    if (store && xact && y < p.Y && ry0 + r < p.ty) {
                        T* d = dbase + (long long)(zo + p.do2) * p.dp2 + (long long)r * p.dp1;
                        if constexpr (VX == 4) *reinterpret_cast<float4*>(d) = make_float4(out[0], out[1], out[2], out[3]);
                        else *reinterpret_cast<double2*>(d) = make_double2(out[0], out[1]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
    }
}

template <typename T, bool MOORE, int RED, bool PAD> static int b3_launch_p(B3Params<T>& p, cudaStream_t st) {
    static thread_local int cfg_dev = -1, ctas_per_sm = 0;
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (dev != cfg_dev) {
        SB_CUDA(cudaFuncSetAttribute(box3d_kernel<T, MOORE, RED, PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, B3_SMEM));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, box3d_kernel<T, MOORE, RED, PAD>, B3_THREADS, B3_SMEM) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        ctas_per_sm = per_sm;
        cfg_dev = dev;
    }
    const long long ctas = (long long)ctas_per_sm * num_sms();
    // Tile height and z-runs: a task loads (ty + 2) rows x (zn / nz + 2) planes; pick the pair that minimises waves x rows
    // loaded per task (the idle tail of a partial last wave against re-read halo rows / planes), as stream3d.cu does.
    int best_ty = B3_TY, best = 1;
    double best_cost = 1e300;
    for (int ty = B3_TY; ty >= B3_TY / 2; ty--) {
        const long long nty = (p.Y + ty - 1) / ty;
        for (int nz = 1; nz <= 64 && nz <= std::max(1, p.zn / 4); nz++) {
            const long long tasks = (long long)p.ntx * nty * nz;
            const long long waves = (tasks + ctas - 1) / ctas;
            const double cost = (double)waves * (ty + 2.0) * ((double)p.zn / nz + 2.0);
            if (cost < best_cost * 0.999) { best_cost = cost; best = nz; best_ty = ty; }
        }
    }
    p.ty = best_ty;
    p.nty = (p.Y + best_ty - 1) / best_ty;
    p.nzruns = best;
    const long long grid = std::min<long long>(ctas, (long long)p.ntx * p.nty * p.nzruns);
    box3d_kernel<T, MOORE, RED, PAD><<<(unsigned)grid, B3_THREADS, B3_SMEM, st>>>(p);
    SB_LAUNCH_CHECK();
    return SB200_OK;
}

template <typename T, bool MOORE, int RED> static int b3_launch(B3Params<T>& p, cudaStream_t st) {
    const bool pad = p.bc0 == SB200_REMOVE || (p.so1 == 0 && p.bc1 == SB200_REMOVE) || (p.so2 == 0 && p.bc2 == SB200_REMOVE);
    return pad ? b3_launch_p<T, MOORE, RED, true>(p, st) : b3_launch_p<T, MOORE, RED, false>(p, st);
}

template <typename T, bool MOORE> static int b3_try(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const sb200_desc& d = pl.d;
    if (d.src_off[0] != 0) return -1;
    if ((d.size[0] * sizeof(T)) % 16 || (d.src_ext[0] * sizeof(T)) % 16 || (d.dst_ext[0] * sizeof(T)) % 16 || (d.dst_off[0] * sizeof(T)) % 16)
        return -1;
    if (((uintptr_t)src | (uintptr_t)dst) & 15) return -1;
    if (d.size[0] * (long long)sizeof(T) < 32 || d.size[0] > (1 << 28) || d.size[1] > (1 << 28) || d.size[2] > (1 << 28)) return -1;
    if (pl.dd.lo[0] != 0 || pl.dd.n[0] != d.size[0] || pl.dd.lo[1] != 0 || pl.dd.n[1] != d.size[1]) return -1;  // z regions only
    for (int a = 0; a < 3; a++)
        if (d.src_off[a] == 0 && d.boundary[a] == SB200_USE) return -1;
    if (pl.dd.n[2] == 0) return SB200_OK;
    B3Params<T> p;
    p.src = (const T*)src; p.dst = (T*)dst;
    p.sp1 = d.src_ext[0]; p.sp2 = d.src_ext[0] * d.src_ext[1];
    p.dp1 = d.dst_ext[0]; p.dp2 = d.dst_ext[0] * d.dst_ext[1];
    p.X = (int)d.size[0]; p.Y = (int)d.size[1]; p.Z = (int)d.size[2];
    p.so1 = d.src_off[1]; p.so2 = d.src_off[2];
    p.do0 = d.dst_off[0]; p.do1 = d.dst_off[1]; p.do2 = d.dst_off[2];
    p.bc0 = d.boundary[0]; p.bc1 = d.boundary[1]; p.bc2 = d.boundary[2];
    memcpy(&p.pad, &d.padval_bits, sizeof(T));
    p.z_lo = (int)pl.dd.lo[2]; p.zn = (int)pl.dd.n[2];
    const long long Xb = d.size[0] * (long long)sizeof(T);
    p.ntx = (int)((Xb + B3_TXB - 1) / B3_TXB);
    p.nty = 0; p.nzruns = 1; p.ty = B3_TY;
    switch (d.reducer) {
    case SB200_SUM: return b3_launch<T, MOORE, SB200_SUM>(p, st);
    case SB200_MEAN: return b3_launch<T, MOORE, SB200_MEAN>(p, st);
    case SB200_MAX: return b3_launch<T, MOORE, SB200_MAX>(p, st);
    case SB200_MIN: return b3_launch<T, MOORE, SB200_MIN>(p, st);
    default: return -1;
    }
}

int try_box3d(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const sb200_desc& d = pl.d;
    if (d.flags & SB200_FLAG_NO_TMA) return -1;
    if (d.ndim != 3 || pl.shape_ndim != 3 || d.radius != 1) return -1;
    const bool window = pl.shape_tag == SB200_WINDOW && d.noffsets == 27, moore = pl.shape_tag == SB200_MOORE && d.noffsets == 26;
    if (!window && !moore) return -1;
    if (d.eltype != d.out_eltype) return -1;
    int rc = -1;
    if (d.eltype == SB200_F32) rc = moore ? b3_try<float, true>(pl, src, dst, st) : b3_try<float, false>(pl, src, dst, st);
    else if (d.eltype == SB200_F64) rc = moore ? b3_try<double, true>(pl, src, dst, st) : b3_try<double, false>(pl, src, dst, st);
    if (rc == SB200_OK) set_kernel_name(moore ? "box3d_kernel<moore>" : "box3d_kernel<window>");
    return rc;
}

}  // namespace sb
