// stream2d.cuh — 2-D gather for Float32/Float64 over compile-time stencil shapes, TMA-fed row streaming.
//
// Replaces gatherstencil_kernel! (src/gatherstencil.jl:105-109) + the neighbour read path (src/array.jl:91-138)
// for the named 2-D shapes (src/stencils/*.jl) and the reducers sum / mean / minimum / maximum / kernelproduct /
// diffusion. Structure (same ring as life.cu):
//   * a CTA owns a strip of S2_WARPS*512 bytes of columns and streams down a run of rows;
//   * one producer thread issues cp.async.bulk (UBLKCP) copies of CH source rows per stage into a ring of shared
//     memory stages and resolves the row boundary (Wrap / Reflect / ring rows) and the Wrap column halo by
//     choosing source addresses; full/empty mbarriers recycle the stages;
//   * every consumer thread owns VX = 16/sizeof(T) consecutive cells (one 128-bit store per output row). Each
//     source row is read from shared memory once (its 16 bytes + R halo cells per side) and folded into the
//     2R+1 output rows it belongs to: accumulators rotate through registers with a period of 2R+1 rows, so the
//     fold of one output visits its taps in exactly the reference's offset order (row by row, first axis
//     fastest) — bit-identical to the left fold of StaticArrays — with no contraction (explicit *_rn ops).
// Algorithmic traffic: sizeof(T) read + sizeof(T) written per cell; each source row is fetched from HBM once
// per run (+2R rows per run of rows).
#pragma once
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "tma.cuh"

namespace sb {

#ifndef SB200_S2_PAIRFOLD_MAXR
#define SB200_S2_PAIRFOLD_MAXR 3   // nested extrema fold rows in pairs through `cen` (three-input ops) up to this radius; beyond it the
                                   // pending halves push the 9-row fold of Circle(4) over the 96-register cap (188 + 160 bytes of spills
                                   // per thread and row): one two-input op per row instead — r02o: 620 -> 683 Gcell/s, 0.835 of the roofline
#endif
#ifndef SB200_S2_EDGE_WARP
#define SB200_S2_EDGE_WARP 1   // Remove padval selects only in the warps that touch the array edge (r02j: Circle(4) max 620 -> 660, 7x7 246 -> 263 Gcell/s)
#endif
constexpr int S2_WARPS = 8;
// kernelproduct with the multiply-add contracted into one FMA (SB200_FLAG_ALLOW_FMA): a reducer code of its own for the templates
constexpr int S2_KDOT_FMA = 100;
__device__ __forceinline__ float s2_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double s2_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
// EXPERIMENT (default 1 = the measured kernels; not yet run on a GPU with another value): producer warps of the element-granular
// cp.async mode (Halo rings / unaligned rows, mean F64 + Halo{:out} 0.74 of peak). One warp issues 16-32 LDGSTS per lane and
// row; the 3-D kernels were producer-bound with far less (lesson 3). Each producer warp copies every
// S2_PRODUCERS-th group of 32 elements of a row; the bulk-copy mode keeps using one lane of the first producer warp.
#ifndef SB200_S2_PRODUCERS
#define SB200_S2_PRODUCERS 1
#endif
constexpr int S2_PRODUCERS = SB200_S2_PRODUCERS;
constexpr int S2_THREADS = (S2_WARPS + S2_PRODUCERS) * 32;
__device__ __forceinline__ bool s2_is_producer(int warp) { return S2_PRODUCERS == 1 ? warp == S2_WARPS : warp >= S2_WARPS; }
constexpr int S2_BXB = S2_WARPS * 32 * 16;  // strip width in bytes

// 2-D shape predicates (same as shape_keep in api.cu for N = 2), usable at compile time.
__host__ __device__ constexpr int s2_abs(int v) { return v < 0 ? -v : v; }
__host__ __device__ constexpr bool s2_has(int shape, int R, int dx, int dy) {
    const int ax = s2_abs(dx), ay = s2_abs(dy), manh = ax + ay, mx = ax > ay ? ax : ay, sq = dx * dx + dy * dy;
    switch (shape) {
    case SB200_WINDOW: return true;
    case SB200_MOORE: return manh != 0;
    case SB200_VONNEUMANN: return manh >= 1 && manh <= R;
    case SB200_CROSS: return dx == 0 || dy == 0;
    case SB200_DIAMOND: return manh <= R;
    case SB200_CIRCLE: return 4 * sq < (2 * R + 1) * (2 * R + 1);  // sqrt(sq) < R + 0.5
    case SB200_CARDINAL: return manh == R && mx == R;
    case SB200_ORDINAL: return manh == 2 * R && mx == R;
    default: return false;
    }
}
// index of tap (dx,dy) in the reference's offset order (dy-major, dx fastest)
__host__ __device__ constexpr int s2_tap_index(int shape, int R, int dx, int dy) {
    int k = 0;
    for (int y = -R; y <= R; y++)
        for (int x = -R; x <= R; x++) {
            if (y == dy && x == dx) return k;
            if (s2_has(shape, R, x, y)) k++;
        }
    return k;
}
__host__ __device__ constexpr int s2_count(int shape, int R) { return s2_tap_index(shape, R, R + 1, R); }
// does stream row-offset dy hold the first / last tap of the fold?
__host__ __device__ constexpr int s2_first_dy(int shape, int R) {
    for (int y = -R; y <= R; y++)
        for (int x = -R; x <= R; x++)
            if (s2_has(shape, R, x, y)) return y;
    return 0;
}
__host__ __device__ constexpr int s2_last_dy(int shape, int R) {
    for (int y = R; y >= -R; y--)
        for (int x = -R; x <= R; x++)
            if (s2_has(shape, R, x, y)) return y;
    return 0;
}

// Half-width of row dy when the row is one contiguous symmetric segment [-hw, hw] (else -1): Window, Circle, Diamond.
__host__ __device__ constexpr int s2_row_halfwidth(int shape, int R, int dy) {
    int hw = -1;
    for (int x = 0; x <= R; x++)
        if (s2_has(shape, R, x, dy)) hw = x;
    if (hw < 0) return -1;
    for (int x = -R; x <= R; x++)
        if (s2_has(shape, R, x, dy) != (s2_abs(x) <= hw)) return -1;
    return hw;
}
__host__ __device__ constexpr bool s2_convex_rows(int shape, int R) {
    for (int y = -R; y <= R; y++)
        if (s2_row_halfwidth(shape, R, y) < 0) return false;
    return true;
}

template <typename T> struct S2Params {
    const T* src;
    T* dst;
    long long spitch, dpitch;  // elements per row
    int W, H;                  // logical size, axis 0 = W
    int soff0, soff1, doff0, doff1;
    int cpasync;               // 1: element-granular cp.async producer (rows / base not 16-byte aligned)
    int delta;                 // bulk copies of a Halo-padded source (ring on axis 0): (src_off0 * sizeof(T)) % 16, the byte shift
                               // of the shared-memory image that keeps global and shared addresses congruent mod 16
    int bc0, bc1;
    T pad;
    int y_lo, rows;
    int nstrips, nruns;
    T alpha;
    T weights[81];             // KERNELDOT: in offset order (kernel parameter space = constant bank operands)
};

template <typename T, int R> struct S2Cfg {
    static constexpr int VX = 16 / (int)sizeof(T);
    static constexpr int HLB = ((R * (int)sizeof(T) + 15) / 16) * 16;   // halo bytes per side in a shared-memory row
    static constexpr int HL = HLB / (int)sizeof(T);                      // ... in elements
    static constexpr int LEFT = 128;   // margin (halo at its end): global and shared addresses of every copy agree mod 128
    static constexpr int ROWB = LEFT + S2_BXB + 128;
    static constexpr int P = 2 * R + 1;                                   // accumulator rotation period
    static constexpr int CH = P * ((R == 1) ? 2 : 1);                     // source rows per stage
    static constexpr int FIT = (113 * 1024 - 128) / (CH * ROWB);          // stages that still allow 2 CTAs per SM
    static constexpr int STAGES = FIT >= 4 ? 4 : (FIT < 2 ? 2 : FIT);
    static constexpr int SMEM = 128 + STAGES * CH * ROWB;
    static constexpr int SEG = VX + 2 * R;
};

template <typename T> __device__ __forceinline__ T s2_ldvec(const unsigned char* p, T* out);
template <> __device__ __forceinline__ float s2_ldvec<float>(const unsigned char* p, float* out) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    return 0.f;
}
template <> __device__ __forceinline__ double s2_ldvec<double>(const unsigned char* p, double* out) {
    const double2 v = *reinterpret_cast<const double2*>(p);
    out[0] = v.x; out[1] = v.y;
    return 0.0;
}
__device__ __forceinline__ void s2_stvec(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void s2_stvec(double* p, const double* v) { *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]); }

template <typename T> __device__ __forceinline__ long long s2_map_row(const S2Params<T>& p, int r) {
    if (p.soff1 > 0) return (long long)r + p.soff1;
    if (r >= 0 && r < p.H) return r;
    if (p.bc1 == SB200_WRAP) return r < 0 ? r + p.H : r - p.H;
    if (p.bc1 == SB200_REFLECT) return r < 0 ? -r : 2 * (p.H - 1) - r;
    return -1;
}

// Per-thread constants of one (strip, run) task.
template <typename T> struct S2Thread {
    int xtb;        // byte offset of the thread's cells inside the strip
    int x0b;        // byte offset of the strip in the row
    int gx;         // global column of the thread's first cell
    int y0, nout;   // first output row of the run, number of output rows
    bool active, edge_l, edge_r, may_pad;
    bool edge_warp; // Remove on axis 0: some lane of this warp has out-of-bounds halo cells (warp-uniform)
    int nl, rlim;   // Remove on axis 0: halo cells e < nl (left) and e >= rlim (right) of the segment are out of bounds
    T* dt;          // dest pointer of (row y0, column gx)
};

// Large folds (R >= 2) keep ONE copy of the row code and shift the accumulators through registers after every row
// (2R+1 register moves per cell against L folds): the fully unrolled rotation of a 7x7 or Circle(4) fold is > 100 KB
// of SASS and stalls on instruction fetch. R == 1 keeps the rotation by renaming (period-P unrolled rows).
template <int SHAPE, int R, int RED> struct S2Roll {
    static constexpr bool nested = (RED == SB200_MAX || RED == SB200_MIN) && s2_convex_rows(SHAPE, R);  // small code
    static constexpr bool value = R >= 2 && !nested;
};

// Fold source row J of the current stage (stream index i0 + J) into the 2R+1 outputs it belongs to.
// SH: the shared-memory image is shifted by p.delta bytes (bulk copies of a Halo-padded source whose ring is not a multiple of
// 16 bytes thick); a compile-time switch because the issue-bound folds (7 x 7 kernelproduct, Circle(4) maximum) lost 5 % to the
// run-time form of it (r02f).
template <typename T, int SHAPE, int R, int RED, int J_, bool SH>
__device__ __forceinline__ void s2_row(const S2Params<T>& p, const S2Thread<T>& th, const unsigned char* sbase, int i0,
                                       T (&acc)[2 * R + 1][16 / sizeof(T)], T (&cen)[2 * R + 1][16 / sizeof(T)], int jrt = 0) {
    using C = S2Cfg<T, R>;
    constexpr int VX = C::VX, P = C::P, SEG = C::SEG, L = s2_count(SHAPE, R);
    constexpr int DY0 = s2_first_dy(SHAPE, R), DY1 = s2_last_dy(SHAPE, R);
    constexpr bool ROLL = S2Roll<SHAPE, R, RED>::value;
    const int J = ROLL ? jrt : J_;   // row of the stage: run time in the rolled form
    const int i = i0 + J;            // stream index of this source row
    const int r = th.y0 - R + i;     // logical source row
    // ---- segment: cells gx-R .. gx+VX-1+R of the row ----
    T seg[SEG];
    if (th.may_pad && (r < 0 || r >= p.H)) {
#pragma unroll
        for (int e = 0; e < SEG; e++) seg[e] = p.pad;
    } else {
        const unsigned char* t = sbase + J * C::ROWB + C::LEFT + (SH ? p.delta : 0) + th.xtb;
        if constexpr (!SH) {
            s2_ldvec<T>(t, &seg[R]);
        } else {   // ring on axis 0 whose thickness is not a multiple of 16 bytes: the thread's cells straddle two vectors
#pragma unroll
            for (int v = 0; v < VX; v++) seg[R + v] = *reinterpret_cast<const T*>(t + v * (int)sizeof(T));
        }
        // (halo cells as whole 128-bit vectors of the neighbouring lanes' cells instead of 2R scalar loads: measured r02l, no gain —
        // Circle(4) 638 -> 618, 7 x 7 unchanged — and removed again)
#pragma unroll
        for (int e = 0; e < R; e++) {
            seg[e] = *reinterpret_cast<const T*>(t - (R - e) * (int)sizeof(T));
            seg[R + VX + e] = *reinterpret_cast<const T*>(t + (VX + e) * (int)sizeof(T));
        }
        if (p.bc0 == SB200_REMOVE) {
            // branch-free inside the warp (a divergent patch loop on the one edge warp would pace its whole CTA), skipped by the
            // warps that have no out-of-bounds halo cell at all (SB200_S2_EDGE_WARP: all but two warps of a 32768-wide row)
            if (!SB200_S2_EDGE_WARP || th.edge_warp) {
#pragma unroll
                for (int e = 0; e < R; e++) {
                    seg[e] = e < th.nl ? p.pad : seg[e];
                    seg[R + VX + e] = e >= th.rlim ? p.pad : seg[R + VX + e];
                }
            }
        } else if (th.edge_l || th.edge_r) {
            const unsigned char* row0 = sbase + J * C::ROWB + C::LEFT + (SH ? p.delta : 0) - th.x0b;  // address of global column 0
#pragma unroll
            for (int e = 0; e < SEG; e++) {
                const int x = th.gx - R + e;
                if (x < 0 || x >= p.W) {
                    if (p.bc0 == SB200_REFLECT) {
                        const int xm = x < 0 ? -x : 2 * (p.W - 1) - x;
                        seg[e] = *reinterpret_cast<const T*>(row0 + (long long)xm * (int)sizeof(T));
                    } else {
                        seg[e] = p.pad;
                    }
                }
            }
        }
    }
    // ---- maximum / minimum over shapes whose rows are contiguous segments: the fold is exact in any order, so the
    // running extrema m[w] over [-w, w] are built once per source row (2 ops per width) and every output takes the
    // one matching its row of the shape: 2R + (2R+1) ops per cell instead of L.
    constexpr bool NESTED = (RED == SB200_MAX || RED == SB200_MIN) && s2_convex_rows(SHAPE, R);
    if constexpr (NESTED) {
        // The running extremum over [-w, w] is widened one step at a time and folded AT ONCE into every output whose row of the
        // shape has that half-width (the outputs are independent accumulators, and max / min do not care about the order), so
        // only one width is live at a time. (Round 1 built all R + 1 widths first: 20 more live registers, and the Circle(4)
        // fold spilled 136 + 140 bytes per thread and row at the 96-register cap — more local-memory instructions than math.)
        T mw[VX];
#pragma unroll
        for (int v = 0; v < VX; v++) mw[v] = seg[R + v];
#pragma unroll
        for (int w_ = 0; w_ <= R; w_++) {
            if (w_ > 0) {
#pragma unroll
                for (int v = 0; v < VX; v++)
                    mw[v] = RED == SB200_MAX ? jl_max3(mw[v], seg[R + v - w_], seg[R + v + w_]) : jl_min3(mw[v], seg[R + v - w_], seg[R + v + w_]);
            }
#pragma unroll
            for (int d = 0; d < P; d++) {
                const int dy = d - R;
                if (dy < DY0 || dy > DY1 || s2_row_halfwidth(SHAPE, R, dy) != w_) continue;
                const int s = ROLL ? d : ((J_ - d) % P + P) % P;
                const int q = dy - DY0;
#pragma unroll
                for (int v = 0; v < VX; v++) {
                    // rows fold in pairs: the row's extremum waits in `cen` (unused by max / min) until the next row arrives and
                    // both enter one three-input instruction; convex shapes span 2R+1 rows, so the last row is a pair's end
                    if (q == 0) acc[s][v] = mw[v];
                    else if (R > SB200_S2_PAIRFOLD_MAXR) acc[s][v] = RED == SB200_MAX ? jl_max(acc[s][v], mw[v]) : jl_min(acc[s][v], mw[v]);
                    else if (q & 1) cen[s][v] = mw[v];
                    else acc[s][v] = RED == SB200_MAX ? jl_max3(acc[s][v], cen[s][v], mw[v]) : jl_min3(acc[s][v], cen[s][v], mw[v]);
                }
            }
        }
    }
    // ---- fold into the outputs o = i - d, d = dy + R ----
    // Rolled form (R >= 2, one copy of the row body): after source row i slot d holds the fold of output o = i - d. Instead of
    // updating in place and then shifting every accumulator one slot (2R x VX moves per row — IMAD.MOV on the same FMA pipe
    // the folds run on: 5 % of the 7 x 7 kernel's instructions, ncu r02p), the slots are visited from the last to the first and
    // the FIRST tap of a row reads the neighbouring slot: acc[d] = acc[d-1] (+) tap — the shift rides on a three-operand add.
#pragma unroll
    for (int dd = 0; dd < P; dd++) {
        const int d = ROLL ? P - 1 - dd : dd;
        const int dy = d - R;
        if (dy < DY0 || dy > DY1) continue;
        // accumulator of output o = i - d: slot d in the rolled form, else the compile-time rotation (stages hold a multiple of P rows)
        const int s = ROLL ? d : ((J_ - d) % P + P) % P;
        const int sp = ROLL ? (d > 0 ? d - 1 : 0) : s;   // where that output's fold stood before this row
        if (NESTED) {
            // folded above, width by width
        } else
        {
        if (RED == SB200_DIFFUSION) {
#pragma unroll
            for (int v = 0; v < VX; v++) {
                if (dy == 0) cen[s][v] = seg[R + v];
                else if (ROLL && dy > 0) cen[s][v] = cen[sp][v];
            }
        }
        bool first = true;   // compile time after unrolling: the first tap of this row of the shape
#pragma unroll
        for (int dx = -R; dx <= R; dx++) {
            if (!s2_has(SHAPE, R, dx, dy)) continue;
            const int kk = s2_tap_index(SHAPE, R, dx, dy);
            const int sr = first ? sp : s;
#pragma unroll
            for (int v = 0; v < VX; v++) {
                const T x = seg[R + v + dx];
                if (RED == SB200_SUM || RED == SB200_MEAN || RED == SB200_DIFFUSION) acc[s][v] = kk == 0 ? x : add_rn(acc[sr][v], x);
                else if (RED == SB200_MAX) acc[s][v] = kk == 0 ? x : jl_max(acc[sr][v], x);
                else if (RED == SB200_MIN) acc[s][v] = kk == 0 ? x : jl_min(acc[sr][v], x);
                else if (RED == SB200_KERNELDOT) acc[s][v] = add_rn(kk == 0 ? T(0) : acc[sr][v], mul_rn(x, p.weights[kk]));
                else if (RED == S2_KDOT_FMA) acc[s][v] = s2_fma(x, p.weights[kk], kk == 0 ? T(0) : acc[sr][v]);
            }
            first = false;
        }
        if (ROLL && first && dy > DY0) {   // a row of the shape without taps: the fold just moves on
#pragma unroll
            for (int v = 0; v < VX; v++) acc[s][v] = acc[sp][v];
        }
        }
        if (dy == DY1) {  // last row of the fold: output o = i - d is complete
            const int o = i - d;
            T out[VX];
#pragma unroll
            for (int v = 0; v < VX; v++) {
                if (RED == SB200_MEAN) out[v] = div_rn(acc[s][v], (T)L);
                else if (RED == SB200_DIFFUSION) {
                    const T cc = cen[s][v];
                    out[v] = add_rn(cc, mul_rn(p.alpha, sub_rn(acc[s][v], mul_rn((T)L, cc))));
                } else out[v] = acc[s][v];
            }
            if (o >= 0 && o < th.nout && th.active) s2_stvec(th.dt + (long long)o * p.dpitch, out);
        }
    }
}

template <typename T, int SHAPE, int R, int RED, int J, bool SH> struct S2Rows {
    static __device__ __forceinline__ void run(const S2Params<T>& p, const S2Thread<T>& th, const unsigned char* sbase, int i0,
                                               T (&acc)[2 * R + 1][16 / sizeof(T)], T (&cen)[2 * R + 1][16 / sizeof(T)]) {
        if constexpr (S2Roll<SHAPE, R, RED>::value) {
#pragma unroll 1
            for (int j = 0; j < S2Cfg<T, R>::CH; j++) s2_row<T, SHAPE, R, RED, 0, SH>(p, th, sbase, i0, acc, cen, j);
        } else {
            s2_row<T, SHAPE, R, RED, J, SH>(p, th, sbase, i0, acc, cen);
            if constexpr (J + 1 < S2Cfg<T, R>::CH) S2Rows<T, SHAPE, R, RED, J + 1, SH>::run(p, th, sbase, i0, acc, cen);
        }
    }
};

template <typename T, int SHAPE, int R, int RED, bool SH>
__global__ void __launch_bounds__(S2_THREADS, 2) stream2d_kernel(const __grid_constant__ S2Params<T> p) {
    using C = S2Cfg<T, R>;
    constexpr int VX = C::VX, P = C::P, CH = C::CH;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + C::STAGES;
    unsigned char* ring = smem + 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; s++) { mbar_init(&full[s], p.cpasync ? 32 * S2_PRODUCERS : 1); mbar_init(&empty[s], S2_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    const int ntasks = p.nstrips * p.nruns;
    const int Wb = p.W * (int)sizeof(T);
    unsigned k = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int strip = task % p.nstrips, run = task / p.nstrips;
        const int x0b = strip * S2_BXB;
        const int wbytes = min(S2_BXB, Wb - x0b);
        const int y0 = p.y_lo + (int)((long long)p.rows * run / p.nruns);
        const int y1 = p.y_lo + (int)((long long)p.rows * (run + 1) / p.nruns);
        const int nout = y1 - y0;
        const int nsrc = nout + 2 * R;  // source rows y0-R .. y1-1+R
        const int nchunks = (nsrc + CH - 1) / CH;
        if (s2_is_producer(warp) && p.cpasync) {
            // ---------------- producer, element-granular (same shared-memory layout: strip cell 0 at LEFT) ----------------
            const int xs = x0b / (int)sizeof(T), wc = wbytes / (int)sizeof(T);
            const bool ring0 = p.soff0 > 0;                       // axis 0 has a ring: every neighbour is a cell of the parent
            const int lA = (xs > 0 || ring0) ? R : 0;
            const int rA = ring0 ? R : min(R, p.W - (xs + wc));
            const bool wrap0 = !ring0 && p.bc0 == SB200_WRAP;
            for (int c = 0; c < nchunks; c++, k++) {
                const int slot = k % C::STAGES;
                mbar_wait_producer(&empty[slot], ((k / C::STAGES) & 1) ^ 1);
                unsigned char* sbase = ring + slot * (CH * C::ROWB);
                for (int j = 0; j < CH; j++) {
                    const int i = c * CH + j;
                    const long long prow = i < nsrc ? s2_map_row(p, y0 - R + i) : -1;
                    if (prow < 0) continue;
                    const T* g = p.src + prow * p.spitch + p.soff0;                        // logical cell 0 of the row
                    T* srow = reinterpret_cast<T*>(sbase + j * C::ROWB + C::LEFT);          // strip cell 0
                    for (int e = lane - lA + 32 * (warp - S2_WARPS); e < wc + rA; e += 32 * S2_PRODUCERS) cp_async_elem(srow + e, g + xs + e);
                    if (S2_PRODUCERS > 1 && warp != S2_WARPS) continue;   // the wrapped halo cells: first producer warp only
                    if (wrap0 && lA == 0)
                        for (int e = lane; e < R; e += 32) cp_async_elem(srow - R + e, g + p.W - R + e);
                    if (wrap0 && rA < R)
                        for (int e = lane; e < R - rA; e += 32) cp_async_elem(srow + wc + rA + e, g + e);
                }
                cp_async_arrive_noinc(&full[slot]);
            }
            continue;
        }
        if (s2_is_producer(warp) && p.soff0 > 0) {
            // ---------------- producer, bulk copies of a Halo-padded source (ring on axis 0: every neighbour is a parent cell) ------
            // The cells the strip needs, logical [x0 - R, x0 + w + R), are parent columns shifted by the ring thickness; the copy
            // is widened to 16-byte boundaries of the PARENT row and lands at LEFT + delta + ..., so that both addresses are
            // 16-byte aligned although logical cell x0 is not (r01: this layout went through per-element cp.async at 0.73 of peak).
            if (lane == 0 && (S2_PRODUCERS == 1 || warp == S2_WARPS)) {
                const int es = (int)sizeof(T);
                const int qlo = (x0b + (p.soff0 - R) * es) & ~15;
                const int qhi = (x0b + wbytes + (p.soff0 + R) * es + 15) & ~15;
                const unsigned mlen = qhi - qlo;
                const int mdst = C::LEFT + p.delta + qlo - p.soff0 * es - x0b;
                for (int c = 0; c < nchunks; c++, k++) {
                    const int slot = k % C::STAGES;
                    mbar_wait_producer(&empty[slot], ((k / C::STAGES) & 1) ^ 1);
                    unsigned char* sbase = ring + slot * (CH * C::ROWB);
                    unsigned bytes = 0;
                    long long prow[CH];
#pragma unroll
                    for (int j = 0; j < CH; j++) {
                        const int i = c * CH + j;
                        prow[j] = i < nsrc ? s2_map_row(p, y0 - R + i) : -1;
                        if (prow[j] >= 0) bytes += mlen;
                    }
                    mbar_arrive_expect_tx(&full[slot], bytes);
#pragma unroll
                    for (int j = 0; j < CH; j++) {
                        if (prow[j] < 0) continue;
                        const unsigned char* g = reinterpret_cast<const unsigned char*>(p.src + prow[j] * p.spitch);
                        bulk_g2s(sbase + j * C::ROWB + mdst, g + qlo, mlen, &full[slot]);
                    }
                }
            } else {
                k += nchunks;
            }
            continue;
        }
        if (s2_is_producer(warp)) {
            // ---------------- producer ----------------
            if (lane == 0 && (S2_PRODUCERS == 1 || warp == S2_WARPS)) {
                // One bulk copy per row covers the strip plus the halo cells that are ordinary neighbours in the row
                // (a narrow last strip may end inside the right halo); only the wrapped halo of an array-edge strip
                // needs its own copy.
                const bool l_in = x0b > 0, r_in = x0b + wbytes < Wb;
                const bool l_wrap = !l_in && p.bc0 == SB200_WRAP, r_wrap = !r_in && p.bc0 == SB200_WRAP;
                const int r_in_bytes = r_in ? min(C::HLB, Wb - (x0b + wbytes)) : 0;
                const int mstart = x0b - (l_in ? C::HLB : 0), mdst = C::LEFT - (l_in ? C::HLB : 0);
                const unsigned mlen = wbytes + (l_in ? C::HLB : 0) + r_in_bytes;
                const int wrap_bytes = min(C::HLB, Wb);
                const unsigned rowbytes = mlen + (l_wrap ? wrap_bytes : 0) + (r_wrap ? wrap_bytes : 0);
                for (int c = 0; c < nchunks; c++, k++) {
                    const int slot = k % C::STAGES;
                    mbar_wait_producer(&empty[slot], ((k / C::STAGES) & 1) ^ 1);
                    unsigned char* sbase = ring + slot * (CH * C::ROWB);
                    unsigned bytes = 0;
                    long long prow[CH];
#pragma unroll
                    for (int j = 0; j < CH; j++) {
                        const int i = c * CH + j;
                        prow[j] = i < nsrc ? s2_map_row(p, y0 - R + i) : -1;
                        if (prow[j] >= 0) bytes += rowbytes;
                    }
                    mbar_arrive_expect_tx(&full[slot], bytes);
#pragma unroll
                    for (int j = 0; j < CH; j++) {
                        if (prow[j] < 0) continue;
                        const unsigned char* g = reinterpret_cast<const unsigned char*>(p.src + prow[j] * p.spitch);
                        unsigned char* srow = sbase + j * C::ROWB;
                        bulk_g2s(srow + mdst, g + mstart, mlen, &full[slot]);
                        if (l_wrap) bulk_g2s(srow + C::LEFT - wrap_bytes, g + Wb - wrap_bytes, wrap_bytes, &full[slot]);
                        if (r_wrap) bulk_g2s(srow + C::LEFT + wbytes, g, wrap_bytes, &full[slot]);
                    }
                }
            } else {
                k += nchunks;
            }
            continue;
        }
        // ---------------- consumers ----------------
        S2Thread<T> th;
        th.xtb = (warp * 32 + lane) * 16;
        th.x0b = x0b;
        th.active = th.xtb < wbytes;
        th.gx = (x0b + th.xtb) / (int)sizeof(T);
        // Threads whose segment crosses the array edge under Remove / Reflect patch their halo cells themselves
        // (under Wrap the producer already copied the wrapped columns).
        const bool ring0 = p.soff0 > 0;   // neighbours beyond the logical edge are ring cells the producer copied
        th.edge_l = th.active && !ring0 && p.bc0 != SB200_WRAP && th.gx - R < 0;
        th.edge_r = th.active && !ring0 && p.bc0 != SB200_WRAP && th.gx + VX - 1 + R >= p.W;
        th.nl = ring0 ? 0 : R - th.gx;                 // > 0 only next to the left edge
        th.rlim = ring0 ? R : p.W - th.gx - VX;        // < R only next to the right edge
        th.edge_warp = __any_sync(0xffffffffu, th.active && (th.nl > 0 || th.rlim < R));
        th.y0 = y0; th.nout = nout;
        th.may_pad = p.soff1 == 0 && p.bc1 == SB200_REMOVE;
        th.dt = p.dst + (long long)(y0 + p.doff1) * p.dpitch + p.doff0 + th.gx;
        T acc[P][VX];
        T cen[P][VX];
#pragma unroll
        for (int s = 0; s < P; s++)
#pragma unroll
            for (int v = 0; v < VX; v++) { acc[s][v] = T(0); cen[s][v] = T(0); }
        for (int c = 0; c < nchunks; c++, k++) {
            const int slot = k % C::STAGES;
            mbar_wait(&full[slot], (k / C::STAGES) & 1);
            const unsigned char* sbase = ring + slot * (CH * C::ROWB);
            S2Rows<T, SHAPE, R, RED, 0, SH>::run(p, th, sbase, c * CH, acc, cen);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
    }
}

// Launch one instantiation; returns SB200_OK or an error.
template <typename T, int SHAPE, int R, int RED, bool SH>
int s2_launch_sh(const S2Params<T>& p0, cudaStream_t st) {
    using C = S2Cfg<T, R>;
    static thread_local int cfg_dev = -1, ctas_per_sm = 0;
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (dev != cfg_dev) {
        SB_CUDA(cudaFuncSetAttribute(stream2d_kernel<T, SHAPE, R, RED, SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stream2d_kernel<T, SHAPE, R, RED, SH>, S2_THREADS, C::SMEM) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        ctas_per_sm = per_sm;
        cfg_dev = dev;
    }
    S2Params<T> p = p0;
    const int Wb = p.W * (int)sizeof(T);
    p.nstrips = (Wb + S2_BXB - 1) / S2_BXB;
    const long long ctas = (long long)ctas_per_sm * num_sms();
    // four tasks per CTA, interleaved: every SM stays busy and the tail is a quarter of a task
    long long nruns = std::max<long long>(1, 4 * ctas / p.nstrips);
    const long long cap = std::max(1, p.rows / (4 * C::P));                  // keep the 2R re-read rows per run small
    if (nruns > cap) {
        // small grid: the cap alone would leave resident CTA slots idle (1000 x 1000 Float64: 166 tasks for 296 slots); one task
        // per slot while a run keeps at least one rotation period of rows (r02m: 10.4 -> 8.3 us per sweep)
        nruns = std::min<long long>(std::max<long long>(cap, ctas / p.nstrips), std::max(1, p.rows / C::P));
    }
    static const int nruns_env = getenv("SB200_S2_NRUNS") ? atoi(getenv("SB200_S2_NRUNS")) : 0;   // A/B knob (runs per strip)
    if (nruns_env > 0) nruns = std::min<long long>(nruns_env, std::max(1, p.rows / C::P));
    p.nruns = (int)nruns;
    const long long grid = std::min<long long>(ctas, (long long)p.nstrips * p.nruns);
    stream2d_kernel<T, SHAPE, R, RED, SH><<<(unsigned)grid, S2_THREADS, C::SMEM, st>>>(p);
    SB_LAUNCH_CHECK();
    return SB200_OK;
}

template <typename T, int SHAPE, int R, int RED>
int s2_launch(const S2Params<T>& p, cudaStream_t st) {
    return p.delta ? s2_launch_sh<T, SHAPE, R, RED, true>(p, st) : s2_launch_sh<T, SHAPE, R, RED, false>(p, st);
}

// Fill the parameters from a plan; false when the plan is outside what the streaming kernels accept.
template <typename T> bool s2_accepts(const Plan& pl, const void* src, void* dst, S2Params<T>& p) {
    const sb200_desc& d = pl.d;
    if (d.ndim != 2 || pl.shape_tag < 0 || pl.shape_ndim != 2) return false;
    // bulk copies need an unpadded axis 0 and 16-byte aligned source rows; a Halo ring on axis 0 or unaligned rows take
    // the element-granular producer. The consumers' 128-bit stores need aligned dest rows either way.
    // (a Halo ring on axis 0 with 16-byte aligned parent rows also takes bulk copies: the shared-memory image is shifted by delta)
    const bool aligned = (d.src_ext[0] * sizeof(T)) % 16 == 0 && ((uintptr_t)src & 15) == 0;
    if ((d.size[0] * sizeof(T)) % 16 || ((uintptr_t)src % sizeof(T))) return false;
    if ((d.dst_ext[0] * sizeof(T)) % 16 || (d.dst_off[0] * sizeof(T)) % 16 || ((uintptr_t)dst & 15)) return false;
    if (d.size[0] > (1LL << 28) || d.size[1] > (1LL << 30)) return false;
    if (d.size[0] * (long long)sizeof(T) < 16 * 2) return false;
    if (pl.dd.lo[0] != 0 || pl.dd.n[0] != d.size[0]) return false;  // regions only along axis 1
    if (d.src_off[1] == 0 && d.boundary[1] == SB200_USE) return false;
    if (d.src_off[0] == 0 && d.boundary[0] == SB200_USE) return false;
    if (d.radius >= d.size[0]) return false;
    const long long Wb_ = d.size[0] * (long long)sizeof(T);
    if (Wb_ < 64) return false;
    if (Wb_ > S2_BXB && (Wb_ % S2_BXB) != 0 && (Wb_ % S2_BXB) < 64) return false;  // last strip narrower than a halo
    p.src = (const T*)src; p.dst = (T*)dst;
    p.spitch = d.src_ext[0]; p.dpitch = d.dst_ext[0];
    p.W = (int)d.size[0]; p.H = (int)d.size[1];
    p.soff0 = d.src_off[0]; p.soff1 = d.src_off[1]; p.doff0 = d.dst_off[0]; p.doff1 = d.dst_off[1];
    p.cpasync = aligned ? 0 : 1;
    p.delta = aligned ? (int)((d.src_off[0] * sizeof(T)) % 16) : 0;
    p.bc0 = d.boundary[0]; p.bc1 = d.boundary[1];
    memcpy(&p.pad, &d.padval_bits, sizeof(T));
    p.y_lo = (int)pl.dd.lo[1]; p.rows = (int)pl.dd.n[1];
    p.alpha = (T)d.alpha;
    if (d.reducer == SB200_KERNELDOT) {
        if (d.noffsets > 81) return false;
        memcpy(p.weights, d.weights_host, sizeof(T) * d.noffsets);
    }
    return true;
}

// One dispatcher per (shape, R) group, defined in stream2d_*.cu; returns -1 when the reducer is not compiled.
template <typename T, int SHAPE, int R> int s2_dispatch_reducer(const S2Params<T>& p, int reducer, cudaStream_t st) {
    switch (reducer) {
    case SB200_SUM: return s2_launch<T, SHAPE, R, SB200_SUM>(p, st);
    case SB200_MEAN: return s2_launch<T, SHAPE, R, SB200_MEAN>(p, st);
    case SB200_MIN: return s2_launch<T, SHAPE, R, SB200_MIN>(p, st);
    case SB200_MAX: return s2_launch<T, SHAPE, R, SB200_MAX>(p, st);
    case SB200_KERNELDOT: return s2_launch<T, SHAPE, R, SB200_KERNELDOT>(p, st);
    case SB200_DIFFUSION: return s2_launch<T, SHAPE, R, SB200_DIFFUSION>(p, st);
    default: return -1;
    }
}

int s2_group_a(const Plan& pl, const void* src, void* dst, cudaStream_t st);  // Window R=1..3
int s2_group_b(const Plan& pl, const void* src, void* dst, cudaStream_t st);  // Moore, VonNeumann, Cross, Diamond R=1..2
int s2_group_c(const Plan& pl, const void* src, void* dst, cudaStream_t st);  // Circle R=2..4
int s2_group_d(const Plan& pl, const void* src, void* dst, cudaStream_t st);  // Cross, Diamond R=1..2

}  // namespace sb
