// stream3d2.cu — TWO diffusion steps per launch for 3-D VonNeumann(1) (SB200_FLAG_DOUBLE_STEP with SB200_DIFFUSION):
// dest = f(f(src)), the intermediate state never touches memory, so an iterated run (SwitchingStencilArray,
// src/gatherstencil.jl:77-83 called in a loop) moves half the HBM bytes per step.
//
// Same 2.5-D streaming structure as stream3d.cu: a CTA owns an (x,y) tile and marches along z; three producer warps feed
// whole source rows (tile + 2 halo rows above / below, + 16 B halo left / right) through a ring of shared-memory
// stages with cp.async.bulk. Per arriving source plane the 16 consumer warps
//   level 1: complete one plane of the INTERMEDIATE state over the tile plus a one-cell rim (every cell once: a thread
//            owns 16 bytes of x by 2 rows; the two rim columns, 2 x 16 cells, are one lane each of a 17th "rim" warp)
//            and put it into a double-buffered shared-memory plane,
//   (one named barrier)
//   level 2: complete one plane of the FINAL state from that plane (own rows from registers, the rows of the
//            neighbouring warp from shared memory, x-neighbours by shuffle) and store it.
// Both levels keep, per cell, the centre of the previous plane and the partial fold (((zm + ym) + xm) + xp) + yp in
// registers, exactly like stream3d.cu does for one level, so a plane costs 9 flops per cell per level and nothing is
// recomputed except the rim. The end lanes of a warp fetch the cells next to its span with predicated loads (inline
// PTX: no divergent branch), ring cursor and dest pointer advance incrementally: 256 instructions per source plane and
// warp, 144 of them the flops. (Measured: r01g, first version, warp-private with a rim recomputed per warp: 27
// instructions per cell-update, issue-bound, 3 % SLOWER than two single sweeps; r01h, shared intermediate plane: 855
// Gcell-updates/s against 635 for single sweeps; r01i rim warp + predicated loads 938; r01j three producer warps 1092;
// r01l final build 1095 Gcell-updates/s on 1024^3 Float32 = 1.36x the one-sweep HBM roofline.)
// Fold order = the reference's offset order (src/stencils/vonneumman.jl:5-15), every operation rounded separately:
// bit-identical to two single sweeps.
//
// Accepted: unpadded Float32 / Float64 parents, Wrap on axes 0 and 1, axis 2 Wrap or an output region that stays two
// planes inside the parent (slab runs). Wrapped halo cells of the intermediate state are recomputed from wrapped source
// cells in the same order, so they equal the cells they stand for bit for bit (a Reflect image would fold its
// neighbours in the opposite order, which is why Reflect is not accepted). Remove axes: the PAD instantiation below
// (padval selects at both time levels), bit-exact against two single CPU sweeps in the GPU tests (r02a), on by default.
// Algorithmic traffic: sizeof(T) read + sizeof(T) written per cell per TWO steps.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "tma.cuh"

namespace sb {

constexpr int D2_WX = 2, D2_WY = 8;               // consumer warps across x and y
constexpr int D2_WARPS = D2_WX * D2_WY;
#ifndef SB200_D2_PRODUCERS
#define SB200_D2_PRODUCERS 3
#endif
#ifndef SB200_D2_STAGES
#define SB200_D2_STAGES 6
#endif
#ifndef SB200_D2_UNROLL
#define SB200_D2_UNROLL 2   // unroll factor of the plane loop: the loop-carried centre / partial-fold registers are renamed instead of copied
                            // (222 -> 190 instructions per plane and warp, a third fewer MOVs; r02aj / r02ak, 1024^3 Float32 under the power cap:
                            // 1117 -> 1151 Gcell-updates/s; 3: 1143, 4: 1158)
#endif
constexpr int D2_UNROLL = SB200_D2_UNROLL;
#ifndef SB200_D2_XSCALAR
#define SB200_D2_XSCALAR 1   // scalar FADDs for the x-neighbour adds of the packed fold (no pair assembly; the operand pairs
                             // (x-1, x) / (x+1, x+2) are not register-aligned). r02at, 1024^3 under the power cap:
                             // 1135 -> 1164 Gcell-updates/s, two runs each, parity green; 0 = packed adds
#endif
#ifndef SB200_D2_PACKED
#define SB200_D2_PACKED 1
#endif
#ifndef SB200_D2_SPLIT
#define SB200_D2_SPLIT 0     // split-phase level barrier (A/B): a warp ARRIVES on an mbarrier once its part of intermediate plane i is in
                             // shared memory, goes on, and only WAITS for the other warps' arrivals of plane i-1 when it starts level 2 of
                             // that plane, one iteration later (four intermediate planes instead of two; level 2 re-reads its own rows).
                             // Measured r02ay (one box, two runs each): bit-exact (13 GPU parity tests incl. the slab plans), but 1206.2 /
                             // 1206.0 against 1248.6 / 1250.0 Gcell-updates/s: the loop grows from 356 to 433 instructions per two planes
                             // (two more LDS.128, the mbarrier polls, the loop skew's predicates), which costs more than the barrier slack
                             // returns. Off.
#endif
// producer warps (the rows of a stage dealt round-robin). Measured on 1024^3 Float32 (r01j): 2 producers 935, 3 producers
// 1092, 4 producers (register cap 80) 1073 Gcell-updates/s: with two, the consumers waited for data 23 % of the time
// (each bulk copy costs ~14 issue slots of lane-by-lane serialisation: ELECT / R2UR / UBLKCP / BRA.U.ANY).
// Re-measured after the instruction cuts of round 2 (r02av, one box, two runs each): 3 producers / 6 stages 1181.2, 1179.9;
// 4 producers 1181.9, 1181.3; 8 stages (225 KB) 1185.7, 1187.6; 4 producers + 8 stages 1171.9, 1172.2: the defaults stay.
constexpr int D2_PRODUCERS = SB200_D2_PRODUCERS;
constexpr int D2_RIMWARP = D2_WARPS;               // one more consumer warp: the rim columns of the intermediate plane, one cell per lane
constexpr int D2_CONSUMERS = D2_WARPS + 1;
constexpr int D2_THREADS = (D2_CONSUMERS + D2_PRODUCERS) * 32;
constexpr int D2_TXB = D2_WX * 512;               // tile width in bytes
constexpr int D2_RT = 2;                          // rows per thread and level
constexpr int D2_TY = (D2_WY - 1) * D2_RT;        // final rows per tile (14: 1024 rows = 74 tiles = whole waves of 148);
                                                  // level 1 covers D2_TY + 2 = D2_WY * D2_RT rows
constexpr int D2_LEFT = 128;                      // margin (halo at its end): global and shared addresses agree mod 128
constexpr int D2_ROWB = D2_LEFT + D2_TXB + 128;   // shared-memory row: margin | tile | margin
constexpr int D2_ROWS = D2_TY + 4;                // source rows of a stage: tile rows + two halo rows on each side
constexpr int D2_STAGE = D2_ROWS * D2_ROWB;
constexpr int D2_STAGES = SB200_D2_STAGES;
constexpr int D2_MROWS = D2_TY + 2;               // intermediate plane: tile rows + one rim row on each side
constexpr int D2_MSTAGE = D2_MROWS * D2_ROWB;
constexpr int D2_MBUFS = SB200_D2_SPLIT ? 4 : 2;   // intermediate planes in shared memory
constexpr int D2_SMEM = 128 + D2_STAGES * D2_STAGE + D2_MBUFS * D2_MSTAGE;
static_assert(D2_SMEM <= 227 * 1024, "ring does not fit");
static_assert(D2_ROWS <= 32 * D2_PRODUCERS, "one producer lane per row");
static_assert((2 * D2_STAGES + 2) * 8 <= 128, "barrier header");

template <typename T> struct D2Params {
    const T* src;
    T* dst;
    long long p1, p2;        // pitches (elements) of axes 1 and 2 (source and dest parents have the same layout)
    int X, Y, Z;
    int bc2;                 // boundary of axis 2 (axes 0 and 1 are Wrap)
    T alpha;
    int z_lo, zn;            // output planes [z_lo, z_lo + zn)
    int ntx, nty, nzruns, ty;
    // PAD kernel only (appended so that the Wrap kernel's parameter layout is unchanged)
    int bc0, bc1;            // boundaries of axes 0 and 1: SB200_WRAP or SB200_REMOVE
    T pad;                   // Remove(padval)
    // MIRROR kernel only: fused ghost-plane push, output planes [m_lo, m_hi) are also stored here (plane m_lo first)
    T* mirror;
    int m_lo, m_hi;
};

__device__ __forceinline__ long long d2_wrap(int r, int n) { return r < 0 ? r + n : (r >= n ? r - n : r); }

template <typename T> struct D2Vec;
template <> struct D2Vec<float> { using type = float4; };
template <> struct D2Vec<double> { using type = double2; };

// one diffusion update, every operation rounded separately:  c + alpha * (s - 6 c)
template <typename T> __device__ __forceinline__ T d2_update(T s, T c, T alpha) {
    return add_rn(c, mul_rn(alpha, sub_rn(s, mul_rn((T)6, c))));
}

template <typename T, int VX> __device__ __forceinline__ void d2_lds(T (&v)[VX], const unsigned char* p) {
    const typename D2Vec<T>::type w = *reinterpret_cast<const typename D2Vec<T>::type*>(p);
    if constexpr (VX == 4) { v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w; }
    else { v[0] = w.x; v[1] = w.y; }
}
template <typename T, int VX> __device__ __forceinline__ void d2_st(void* p, const T (&v)[VX]) {
    if constexpr (VX == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
}

// Predicated shared-memory accesses as single instructions (`@p LDS` / `@p STS`): the end lanes of a warp fetch the cells
// next to its span without a divergent branch (ptxas turns `if (lane == 0) x = *p;` into BSSY / BRA / BSYNC sequences).
__device__ __forceinline__ void d2_lds_if(float& v, const void* p, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.shared.f32 %0, [%1];\n\t}" : "+f"(v) : "r"(smem_u32(p)), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void d2_lds_if(double& v, const void* p, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.shared.f64 %0, [%1];\n\t}" : "+d"(v) : "r"(smem_u32(p)), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void d2_sts_if(void* p, float v, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q st.shared.f32 [%0], %1;\n\t}" ::"r"(smem_u32(p)), "f"(v), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void d2_sts_if(void* p, double v, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q st.shared.f64 [%0], %1;\n\t}" ::"r"(smem_u32(p)), "d"(v), "r"((int)pred) : "memory");
}

// One plane of one level for one row of a thread: `c` = this plane's centre cells, ym / yp = the rows above / below,
// l_ / r_ = the cells left of c[0] / right of c[VX-1]. Completes the previous plane (returned in `done`) and starts this one.
template <typename T, int VX>
__device__ __forceinline__ void d2_plane(T (&done)[VX], T (&cprev)[VX], T (&part)[VX], const T (&c)[VX], const T (&ym)[VX],
                                         const T (&yp)[VX], T l_, T r_, T alpha) {
#pragma unroll
    for (int v = 0; v < VX; v++) {
        const T cc = cprev[v];
        done[v] = d2_update(add_rn(part[v], c[v]), cc, alpha);
        const T xm = v == 0 ? l_ : c[v == 0 ? 0 : v - 1];
        const T xp = v == VX - 1 ? r_ : c[v == VX - 1 ? v : v + 1];
        T a = add_rn(cc, ym[v]);
        a = add_rn(a, xm);
        a = add_rn(a, xp);
        a = add_rn(a, yp[v]);
        part[v] = a;
        cprev[v] = c[v];
    }
}

// The named barrier between the two levels. One (non-inlined) instruction for the main warps and the rim warp alike:
// compute-sanitizer's synccheck reports "divergent thread(s) in block" when the warps of a block meet at the same named
// barrier through different BAR.SYNC instructions (r01l), which the hardware allows but the tool does not model.
__device__ __noinline__ void d2_level_barrier() {
    asm volatile("bar.sync 1, %0;" ::"n"(D2_CONSUMERS * 32) : "memory");
}

#if SB200_D2_PACKED
// The same plane step for Float32 with packed add / sub / fma.rn.f32x2 (SASS FADD2 / FFMA2; default since r02b, -DSB200_D2_PACKED=0
// restores the scalar folds): the kernel is bound by issue slots (65 % busy, FMA pipe 34 %), and a packed instruction advances
// two cells per slot. Every lane of a packed instruction rounds like the scalar one. Products are written as
// fma(x, w, -0.0) == rn(x * w) (exact product plus -0 changes nothing, signed zeros and NaN included) with the -0.0 pair read
// from constant memory, so that ptxas can neither fold the fma into a multiply nor contract it with the following add /
// subtract into one FFMA2 (single rounding) — which it does for mul.rn.f32x2 + add.rn.f32x2 and for a literal -0.0 addend.
// Measured r02b, 1024^3 Float32: 1076 -> 1184 Gcell-updates/s, bit-identical to two single sweeps of the CPU restatement.
__device__ __forceinline__ unsigned long long d2_pk(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void d2_upk(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long d2_add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long d2_sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// (-0.0f, -0.0f) read from constant memory: ptxas cannot know the value (the host may overwrite a __constant__), so it can
// neither simplify fma(a, b, -0) to a multiply nor contract it with the next add. r02a: with the literal it did both —
// `FFMA2 R26, -R74, 6, R26` = s - 6 cc in ONE rounding, 1-ulp differences in 40 % of the cells.
__constant__ unsigned long long d2_negzero2 = 0x8000000080000000ull;
__device__ __forceinline__ unsigned long long d2_mul2_opaque(unsigned long long a, unsigned long long b) {   // rn(a * b), never contracted
    unsigned long long r;
    const unsigned long long nz = d2_negzero2;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
    return r;
}
template <>
__device__ __forceinline__ void d2_plane<float, 4>(float (&done)[4], float (&cprev)[4], float (&part)[4], const float (&c)[4],
                                                   const float (&ym)[4], const float (&yp)[4], float l_, float r_, float alpha) {
    const unsigned long long six = d2_pk(6.0f, 6.0f), al = d2_pk(alpha, alpha);
    const unsigned long long mid12 = d2_pk(c[1], c[2]);                      // xp of cells 0,1 and xm of cells 2,3
    const unsigned long long xm[2] = {d2_pk(l_, c[0]), mid12}, xp[2] = {mid12, d2_pk(c[3], r_)};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const unsigned long long cc = d2_pk(cprev[2 * h], cprev[2 * h + 1]), cv = d2_pk(c[2 * h], c[2 * h + 1]);
        // done = cc + alpha * ((part + c) - 6 * cc), every operation rounded separately
        const unsigned long long s = d2_add2(d2_pk(part[2 * h], part[2 * h + 1]), cv);
        const unsigned long long u = d2_sub2(s, d2_mul2_opaque(six, cc));
        d2_upk(d2_add2(cc, d2_mul2_opaque(al, u)), done[2 * h], done[2 * h + 1]);
        // part = (((cc + ym) + xm) + xp) + yp
        unsigned long long a = d2_add2(cc, d2_pk(ym[2 * h], ym[2 * h + 1]));
#if SB200_D2_XSCALAR
        {   // the two x-neighbour adds as scalar FADDs: their operand pairs (x-1, x) / (x+1, x+2) are not register-aligned
            float a0, a1;
            d2_upk(a, a0, a1);
            const float xm0 = h == 0 ? l_ : c[1], xm1 = h == 0 ? c[0] : c[2], xp0 = h == 0 ? c[1] : c[3], xp1 = h == 0 ? c[2] : r_;
            a0 = __fadd_rn(__fadd_rn(a0, xm0), xp0);
            a1 = __fadd_rn(__fadd_rn(a1, xm1), xp1);
            a = d2_pk(a0, a1);
        }
#else
        a = d2_add2(a, xm[h]);
        a = d2_add2(a, xp[h]);
#endif
        a = d2_add2(a, d2_pk(yp[2 * h], yp[2 * h + 1]));
        d2_upk(a, part[2 * h], part[2 * h + 1]);
        cprev[2 * h] = c[2 * h];
        cprev[2 * h + 1] = c[2 * h + 1];
    }
}
#endif

// PAD = true (Remove axes; SB200_D2_REMOVE=0 makes diffusion2_accepts decline them): out-of-bounds
// source cells AND out-of-bounds cells of the intermediate state read padval (Remove boundary: the second sweep sees
// padval outside the array, not an update of it). Rows / planes / edge halos outside the array are not copied; the values
// are substituted by selects. PAD = false is the measured Wrap kernel, its code is untouched (if constexpr).
// MIRROR = true (default for boundary sweeps of slab runs since r02i; SB200_D2_MIRROR=0 falls back to the copy): the planes a
// neighbour GPU needs are also stored into its landing slot by the sweep itself (sb200_desc.mirror_*), like stream3d_kernel's
// MIRROR variant; without it do_gather copies them after the sweep. Bit-exact on 2 GPUs (r02d); 2 x 1024^3 weak scaling 0.980
// with, 0.973 without (r02i; round 1 measured 0.954 with the copy and the slower exchange of the Python iterator).
template <typename T, bool PAD, bool MIRROR>
__global__ void __launch_bounds__(D2_THREADS, 1) stream3d2_kernel(const __grid_constant__ D2Params<T> p) {
    constexpr int VX = 16 / (int)sizeof(T);
    const bool padx = PAD && p.bc0 == SB200_REMOVE, pady = PAD && p.bc1 == SB200_REMOVE, padz = PAD && p.bc2 == SB200_REMOVE;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + D2_STAGES;
    unsigned char* ring = smem + 128;
    uint64_t* midbar = empty + D2_STAGES;                // SPLIT: arrivals of the consumer warps per intermediate plane (two, alternating)
    unsigned char* mbuf = ring + D2_STAGES * D2_STAGE;   // D2_MBUFS intermediate planes
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < D2_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], D2_CONSUMERS); }
        if (SB200_D2_SPLIT) { mbar_init(&midbar[0], D2_CONSUMERS); mbar_init(&midbar[1], D2_CONSUMERS); }
        mbar_fence_init();
    }
    __syncthreads();
    const int ntiles = p.ntx * p.nty;
    const int ntasks = ntiles * p.nzruns;
    const int Xb = p.X * (int)sizeof(T);
    unsigned k = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int tile = task % ntiles, zrun = task / ntiles;
        const int x0b = (tile % p.ntx) * D2_TXB, y0 = (tile / p.ntx) * p.ty;
        const int wbytes = min(D2_TXB, Xb - x0b);
        const int z0 = p.z_lo + (int)((long long)p.zn * zrun / p.nzruns);
        const int z1 = p.z_lo + (int)((long long)p.zn * (zrun + 1) / p.nzruns);
        const int nsrc = z1 - z0 + 4;  // source planes z0-2 .. z1+1
        if (warp >= D2_CONSUMERS) {
            // ---------------- producer warps: lane j of producer w copies shared-memory row P*j+w = logical row y0-2+P*j+w ----------------
            const int pw = warp - D2_CONSUMERS;
            const bool l_in = x0b > 0, r_in = x0b + wbytes < Xb;   // else: the Wrap image of the other array edge
            const int mstart = x0b - (l_in ? 16 : 0);
            const unsigned mlen = wbytes + (l_in ? 16 : 0) + (r_in ? 16 : 0);
            const int mdst = D2_LEFT - (l_in ? 16 : 0);
            unsigned rowbytes = wbytes + 32;
            if constexpr (PAD) if (padx) rowbytes = mlen;   // no Wrap image at an array edge
            const int srow_i = D2_PRODUCERS * lane + pw;
            long long yrow = -1;
            if (srow_i < p.ty + 4) {
                const int y = y0 - 2 + srow_i;
                if (y <= p.Y + 1) yrow = d2_wrap(y, p.Y);
                if constexpr (PAD) if (pady) yrow = (y >= 0 && y < p.Y) ? y : -1;
            }
            unsigned nrows = (unsigned)min(p.ty + 4, p.Y + 4 - y0);   // rows of the stage that are copied (both producers)
            if constexpr (PAD) if (pady) nrows = (unsigned)(min(p.Y - 1, y0 + p.ty + 1) - max(0, y0 - 2) + 1);
            for (int i = 0; i < nsrc; i++, k++) {
                const int slot = k % D2_STAGES;
                int zl = z0 - 2 + i;
                if (p.bc2 == SB200_WRAP) zl = (int)d2_wrap(zl, p.Z);
                bool zin = true;
                if constexpr (PAD) zin = !(padz && (zl < 0 || zl >= p.Z));   // a plane outside the array is not copied
                if (lane == 0) {
                    mbar_wait_producer(&empty[slot], ((k / D2_STAGES) & 1) ^ 1);
                    if (pw == 0) mbar_arrive_expect_tx(&full[slot], zin ? nrows * rowbytes : 0u);
                }
                __syncwarp();
                if (yrow >= 0 && zin) {
                    const unsigned char* g = reinterpret_cast<const unsigned char*>(p.src + (long long)zl * p.p2 + yrow * p.p1);
                    unsigned char* srow = ring + slot * D2_STAGE + srow_i * D2_ROWB;
                    bulk_g2s(srow + mdst, g + mstart, mlen, &full[slot]);
                    if (!l_in && !padx) bulk_g2s(srow + D2_LEFT - 16, g + Xb - 16, 16, &full[slot]);
                    if (!r_in && !padx) bulk_g2s(srow + D2_LEFT + wbytes, g, 16, &full[slot]);
                }
            }
            continue;
        }
        if (warp == D2_RIMWARP) {
            // ---------------- rim warp: level 1 of the two columns next to the tile (x = -1 and x = tile width) ----------------
            // lane = 16 * side + intermediate row; only level 2 of the tile's first / last cell of a row reads these cells
            static_assert(D2_MROWS == 16, "one lane per rim cell");
            const int mrow = lane & 15;
            const int pos = (lane >> 4) ? D2_LEFT + D2_TXB : D2_LEFT - (int)sizeof(T);
            T ce = T(0), qe = T(0);
            int slot = k % D2_STAGES;
            unsigned phase = (k / D2_STAGES) & 1;
            // PAD: which of this rim cell's inputs lie outside the array (task constants; the plane test is per iteration)
            const int yr = y0 - 1 + mrow;
            const bool x_oob = padx && ((lane >> 4) ? x0b + wbytes == Xb : x0b == 0);
            const bool row_oob = pady && (yr < 0 || yr >= p.Y), up_oob = pady && (yr - 1 < 0 || yr - 1 >= p.Y), dn_oob = pady && (yr + 1 < 0 || yr + 1 >= p.Y);
#if SB200_D2_SPLIT
            for (int i = 0; i <= nsrc; i++) {
                if (i >= 1) mbar_wait(&midbar[(k - 1) & 1], ((k - 1) >> 1) & 1);   // keeps this warp within the other warps' window of planes
                if (i == nsrc) break;
#else
            for (int i = 0; i < nsrc; i++, k++) {
#endif
                mbar_wait(&full[slot], phase);
                const unsigned char* t = ring + slot * D2_STAGE + (mrow + 1) * D2_ROWB + pos;   // intermediate row m <-> source row m + 1
                T c = *reinterpret_cast<const T*>(t);
                T yu = *reinterpret_cast<const T*>(t - D2_ROWB), yd = *reinterpret_cast<const T*>(t + D2_ROWB);
                bool m_oob = false;
                if constexpr (PAD) {
                    const int sz = z0 - 2 + i;                       // source plane of this stage; it completes intermediate plane sz - 1
                    const bool zs_oob = padz && (sz < 0 || sz >= p.Z);
                    m_oob = x_oob || row_oob || (padz && (sz - 1 < 0 || sz - 1 >= p.Z));
                    if (zs_oob || row_oob) c = p.pad;
                    if (zs_oob || up_oob) yu = p.pad;
                    if (zs_oob || dn_oob) yd = p.pad;
                }
                T m = d2_update(add_rn(qe, c), ce, p.alpha);
                if constexpr (PAD) if (m_oob) m = p.pad;
                T a = add_rn(ce, yu);
                a = add_rn(a, *reinterpret_cast<const T*>(t - sizeof(T)));
                a = add_rn(a, *reinterpret_cast<const T*>(t + sizeof(T)));
                a = add_rn(a, yd);
                qe = a;
                ce = c;
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
#if SB200_D2_SPLIT
                *reinterpret_cast<T*>(mbuf + (k & 3) * D2_MSTAGE + mrow * D2_ROWB + pos) = m;
                __syncwarp();
                if (lane == 0) mbar_arrive(&midbar[k & 1]);
                k++;
#else
                *reinterpret_cast<T*>(mbuf + (k & 1) * D2_MSTAGE + mrow * D2_ROWB + pos) = m;
                d2_level_barrier();
#endif
                if (++slot == D2_STAGES) { slot = 0; phase ^= 1; }
            }
            continue;
        }
        // ---------------- consumers ----------------
        const int wx = warp % D2_WX, wy = warp / D2_WX;
        const int xtb = (wx * 32 + lane) * 16;         // byte offset inside the tile
        const int r1 = wy * D2_RT;                     // level 1: tile rows r1-1, r1 (intermediate-plane rows r1, r1+1); level 2: tile rows r1, r1+1
        const bool xact = xtb < wbytes;
        const int gx = (x0b + xtb) / (int)sizeof(T);
        const bool lvl2 = wy < D2_WY - 1;
        T c1[D2_RT][VX], q1[D2_RT][VX], c2[D2_RT][VX], q2[D2_RT][VX];
#pragma unroll
        for (int r = 0; r < D2_RT; r++) {
#pragma unroll
            for (int v = 0; v < VX; v++) { c1[r][v] = T(0); q1[r][v] = T(0); c2[r][v] = T(0); q2[r][v] = T(0); }
        }
        // per-thread constants of the march: store predicates, the running dest pointer (plane z0-4 at i = 0), ring cursor
        bool rowok[D2_RT];
#pragma unroll
        for (int r = 0; r < D2_RT; r++) rowok[r] = lvl2 && xact && y0 + r1 + r < p.Y && r1 + r < p.ty;
        constexpr int LAG = SB200_D2_SPLIT ? 1 : 0;   // SPLIT: level 2 of a plane runs one iteration after its level 1
        T* __restrict__ dptr = p.dst + (long long)(y0 + r1) * p.p1 + gx + (long long)(z0 - 4 - LAG) * p.p2;
        T* mptr = nullptr;   // MIRROR: where output plane z0-4+i of this thread's rows lands in the neighbour's slot
        if constexpr (MIRROR) mptr = p.mirror + (long long)(y0 + r1) * p.p1 + gx + (long long)(z0 - 4 - LAG - p.m_lo) * p.p2;
        const bool l0 = lane == 0, l31 = lane == 31;
        int slot = k % D2_STAGES;
        unsigned phase = (k / D2_STAGES) & 1;
        // PAD: out-of-bounds tests that do not depend on the plane
        const bool xoob = padx && gx >= p.X;                              // lanes past the right array edge of a ragged edge tile
        const bool edge_l = padx && gx == 0, edge_r = padx && gx + VX == p.X;
        bool srow_oob[D2_RT + 2], mrow_oob[D2_RT];
#pragma unroll
        for (int q = 0; q < D2_RT + 2; q++) srow_oob[q] = pady && (y0 + r1 - 2 + q < 0 || y0 + r1 - 2 + q >= p.Y);
#pragma unroll
        for (int j = 0; j < D2_RT; j++) mrow_oob[j] = xoob || (pady && (y0 + r1 - 1 + j < 0 || y0 + r1 - 1 + j >= p.Y));
#if SB200_D2_SPLIT
#pragma unroll D2_UNROLL
        for (int i = 0; i <= nsrc; i++) {
            const unsigned kprev = k - 1;              // the intermediate plane of the previous iteration
            if (i < nsrc) {
            mbar_wait(&full[slot], phase);
            const unsigned char* sb_ = ring + slot * D2_STAGE + D2_LEFT + xtb + r1 * D2_ROWB;   // source row r1
            unsigned char* mb_ = mbuf + (k & 3) * D2_MSTAGE + D2_LEFT + xtb + r1 * D2_ROWB;    // intermediate row r1
#else
#pragma unroll D2_UNROLL
        for (int i = 0; i < nsrc; i++, k++) {
            mbar_wait(&full[slot], phase);
            // this thread's 16 bytes in shared-memory row 0; tile row t lives in source row t + 2 and intermediate row t + 1
            const unsigned char* sb_ = ring + slot * D2_STAGE + D2_LEFT + xtb + r1 * D2_ROWB;   // source row r1
            unsigned char* mb_ = mbuf + (k & 1) * D2_MSTAGE + D2_LEFT + xtb + r1 * D2_ROWB;    // intermediate row r1
#endif
            // ---- level 1: source plane s completes intermediate plane s-1 on tile rows r1-1, r1 ----
            T rowv[D2_RT + 2][VX];                     // source tile rows r1-2 .. r1+1 = source rows r1 .. r1+3
#pragma unroll
            for (int q = 0; q < D2_RT + 2; q++) d2_lds<T, VX>(rowv[q], sb_ + q * D2_ROWB);
            bool zs_oob = false, zm_oob = false;       // PAD: source plane s / intermediate plane s-1 outside the array
            if constexpr (PAD) {
                const int sz = z0 - 2 + i;
                zs_oob = padz && (sz < 0 || sz >= p.Z);
                zm_oob = padz && (sz - 1 < 0 || sz - 1 >= p.Z);
#pragma unroll
                for (int q = 0; q < D2_RT + 2; q++)
                    if (zs_oob || srow_oob[q]) {
#pragma unroll
                        for (int v = 0; v < VX; v++) rowv[q][v] = p.pad;
                    }
            }
            T mid[D2_RT][VX];
#pragma unroll
            for (int j = 0; j < D2_RT; j++) {
                const unsigned char* t = sb_ + (j + 1) * D2_ROWB;   // source centre row of intermediate tile row r1-1+j
                T l_ = __shfl_up_sync(0xffffffffu, rowv[j + 1][VX - 1], 1);
                T r_ = __shfl_down_sync(0xffffffffu, rowv[j + 1][0], 1);
                d2_lds_if(l_, t - sizeof(T), l0);
                d2_lds_if(r_, t + 16, l31);
                if constexpr (PAD) {
                    if (zs_oob || srow_oob[j + 1] || edge_l) l_ = p.pad;
                    if (zs_oob || srow_oob[j + 1] || edge_r) r_ = p.pad;
                }
                d2_plane<T, VX>(mid[j], c1[j], q1[j], rowv[j + 1], rowv[j], rowv[j + 2], l_, r_, p.alpha);
                if constexpr (PAD)
                    if (zm_oob || mrow_oob[j]) {
#pragma unroll
                        for (int v = 0; v < VX; v++) mid[j][v] = p.pad;
                    }
            }
            __syncwarp();
            if (l0) mbar_arrive(&empty[slot]);          // every read of the source stage is done
#pragma unroll
            for (int j = 0; j < D2_RT; j++) {
                d2_st<T, VX>(mb_ + j * D2_ROWB, mid[j]);
            }
#if SB200_D2_SPLIT
            __syncwarp();
            if (l0) mbar_arrive(&midbar[k & 1]);        // this warp's part of intermediate plane k is in shared memory
            k++;
            if (++slot == D2_STAGES) { slot = 0; phase ^= 1; }
            }
            if (i >= 1) mbar_wait(&midbar[kprev & 1], (kprev >> 1) & 1);   // every warp's part of the previous plane
            // ---- level 2 of the PREVIOUS plane: intermediate plane m completes final plane m-1 on tile rows r1, r1+1 ----
            if (lvl2 && i >= 1) {
                const unsigned char* mb_ = mbuf + (kprev & 3) * D2_MSTAGE + D2_LEFT + xtb + r1 * D2_ROWB;
                T mid[D2_RT][VX], m2[VX], m3[VX];       // own rows (re-read) and intermediate tile rows r1+1, r1+2 (the warp below)
                d2_lds<T, VX>(mid[0], mb_);
                d2_lds<T, VX>(mid[1], mb_ + D2_ROWB);
                d2_lds<T, VX>(m2, mb_ + 2 * D2_ROWB);
                d2_lds<T, VX>(m3, mb_ + 3 * D2_ROWB);
                const bool store = i >= 4 + LAG;
#else
            d2_level_barrier();
            // ---- level 2: intermediate plane m = s-1 completes final plane m-1 = s-2 on tile rows r1, r1+1 ----
            if (lvl2) {
                T m2[VX], m3[VX];                       // intermediate tile rows r1+1, r1+2 (owned by the warp below)
                d2_lds<T, VX>(m2, mb_ + 2 * D2_ROWB);
                d2_lds<T, VX>(m3, mb_ + 3 * D2_ROWB);
                const bool store = i >= 4;
#endif
#pragma unroll
                for (int r = 0; r < D2_RT; r++) {      // final tile row r1+r: centre = intermediate row r1+r+1
                    const T(&ym)[VX] = r == 0 ? mid[0] : mid[1];
                    const T(&cc)[VX] = r == 0 ? mid[1] : m2;
                    const T(&yp)[VX] = r == 0 ? m2 : m3;
                    const unsigned char* t = mb_ + (r + 1) * D2_ROWB;
                    T l_ = __shfl_up_sync(0xffffffffu, cc[VX - 1], 1);
                    T r_ = __shfl_down_sync(0xffffffffu, cc[0], 1);
                    d2_lds_if(l_, t - sizeof(T), l0);
                    d2_lds_if(r_, t + 16, l31);
                    T out[VX];
                    d2_plane<T, VX>(out, c2[r], q2[r], cc, ym, yp, l_, r_, p.alpha);
                    if (store && rowok[r]) d2_st<T, VX>(dptr + (long long)r * p.p1, out);
                    if constexpr (MIRROR) {
                        const int zo = z0 - 4 - LAG + i;
                        if (store && rowok[r] && zo >= p.m_lo && zo < p.m_hi) d2_st<T, VX>(mptr + (long long)r * p.p1, out);
                    }
                }
            }
            if constexpr (MIRROR) mptr += p.p2;
            dptr += p.p2;
#if !SB200_D2_SPLIT
            if (++slot == D2_STAGES) { slot = 0; phase ^= 1; }
#endif
        }
    }
}

template <typename T, bool PAD, bool MIRROR> static int d2_launch(D2Params<T>& p, cudaStream_t st) {
    static thread_local int cfg_dev = -1;
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (dev != cfg_dev) {
        SB_CUDA(cudaFuncSetAttribute(stream3d2_kernel<T, PAD, MIRROR>, cudaFuncAttributeMaxDynamicSharedMemorySize, D2_SMEM));
        cfg_dev = dev;
    }
    const long long ctas = num_sms();   // one CTA per SM (the ring takes most of the shared memory)
    // Tile height and z-runs: a task loads (ty + 4) rows x (zn / nz + 4) planes; pick the pair that minimises
    // waves x rows loaded per task (idle tail of a partial last wave against re-read halo rows / planes).
    int best_ty = D2_TY, best = 1;
    double best_cost = 1e300;
    for (int ty = D2_TY; ty >= D2_TY / 2; ty -= D2_RT) {
        const long long nty = (p.Y + ty - 1) / ty;
        for (int nz = 1; nz <= 64 && nz <= std::max(1, p.zn / 8); nz++) {
            const long long tasks = (long long)p.ntx * nty * nz;
            const long long waves = (tasks + ctas - 1) / ctas;
            const double cost = (double)waves * (ty + 4.0) * ((double)p.zn / nz + 4.0);
            if (cost < best_cost * 0.999) { best_cost = cost; best = nz; best_ty = ty; }
        }
    }
    p.ty = best_ty;
    p.nty = (p.Y + best_ty - 1) / best_ty;
    p.nzruns = best;
    const long long ntasks = (long long)p.ntx * p.nty * p.nzruns;
    const long long grid = std::min<long long>(ctas, ntasks);
    stream3d2_kernel<T, PAD, MIRROR><<<(unsigned)grid, D2_THREADS, D2_SMEM, st>>>(p);
    SB_LAUNCH_CHECK();
    return SB200_OK;
}

// Can this sweep run two diffusion steps per launch?
bool diffusion2_accepts(const sb200_desc& d, const Plan& pl) {
    if (d.reducer != SB200_DIFFUSION || d.ndim != 3 || (d.eltype != SB200_F32 && d.eltype != SB200_F64)) return false;
    if (pl.shape_tag != SB200_VONNEUMANN || pl.shape_ndim != 3 || d.radius != 1 || d.noffsets != 6) return false;
    if (d.flags & (SB200_FLAG_NO_TMA | SB200_FLAG_FORCE_GENERIC | SB200_FLAG_QUAD_STEP | SB200_FLAG_OCT_STEP)) return false;
    const long long es = (long long)elsize(d.eltype);
    for (int a = 0; a < 3; a++) {
        if (d.src_off[a] != 0 || d.dst_off[a] != 0 || d.dst_ext[a] != d.size[a] || d.src_ext[a] != d.size[a]) return false;
        if (d.size[a] < 4 || d.size[a] > (1 << 28)) return false;
    }
    if ((d.size[0] * es) % 16 || d.size[0] * es < 64 || d.size[0] * es >= (1LL << 30)) return false;   // row bytes are held in an int
    // Remove axes run the PAD variant of the kernel (bit-exact on the GPU in r02a; SB200_D2_REMOVE=0 declines them again)
    static const bool remove_ok = !(getenv("SB200_D2_REMOVE") && atoi(getenv("SB200_D2_REMOVE")) == 0);
    for (int a = 0; a < 2; a++)
        if (d.boundary[a] != SB200_WRAP && !(remove_ok && d.boundary[a] == SB200_REMOVE)) return false;
    if (pl.dd.lo[0] != 0 || pl.dd.n[0] != d.size[0] || pl.dd.lo[1] != 0 || pl.dd.n[1] != d.size[1]) return false;  // z regions only
    const long long lo = pl.dd.lo[2], hi = lo + pl.dd.n[2];
    const bool inside = lo >= 2 && hi + 2 <= d.size[2];   // never leaves the parent: the boundary rule of axis 2 is not exercised
    if (!inside && d.boundary[2] != SB200_WRAP && !(remove_ok && d.boundary[2] == SB200_REMOVE)) return false;
    return true;
}

template <typename T> static int d2_try(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const sb200_desc& d = pl.d;
    if (((uintptr_t)src | (uintptr_t)dst) & 15) return -1;
    if (pl.dd.n[2] == 0) return SB200_OK;
    D2Params<T> p;
    p.src = (const T*)src; p.dst = (T*)dst;
    p.p1 = d.size[0]; p.p2 = d.size[0] * d.size[1];
    p.X = (int)d.size[0]; p.Y = (int)d.size[1]; p.Z = (int)d.size[2];
    p.bc2 = d.boundary[2];
    p.alpha = (T)d.alpha;
    p.z_lo = (int)pl.dd.lo[2]; p.zn = (int)pl.dd.n[2];
    const long long Xb = d.size[0] * (long long)sizeof(T);
    p.ntx = (int)((Xb + D2_TXB - 1) / D2_TXB);
    p.nty = 0; p.nzruns = 1; p.ty = D2_TY;
    p.bc0 = d.boundary[0]; p.bc1 = d.boundary[1];
    memcpy(&p.pad, &d.padval_bits, sizeof(T));
    const long long lo = pl.dd.lo[2], hi = lo + pl.dd.n[2];
    const bool z_remove = d.boundary[2] == SB200_REMOVE && !(lo >= 2 && hi + 2 <= d.size[2]);
    p.mirror = nullptr; p.m_lo = p.m_hi = 0;
    if (d.boundary[0] == SB200_REMOVE || d.boundary[1] == SB200_REMOVE || z_remove) return d2_launch<T, true, false>(p, st);
    static const bool mirror_ok = !(getenv("SB200_D2_MIRROR") && atoi(getenv("SB200_D2_MIRROR")) == 0);
    if (g_mirror.ptr && mirror_ok) {   // fused ghost-plane push (Wrap kernel; bit-exact on 2 GPUs r02d, +0.7 % at N = 2 r02i)
        p.mirror = (T*)g_mirror.ptr; p.m_lo = (int)g_mirror.lo; p.m_hi = (int)g_mirror.hi;
        g_mirror.honoured = true;
        return d2_launch<T, false, true>(p, st);
    }
    return d2_launch<T, false, false>(p, st);
}

int try_diffusion3d_double(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const sb200_desc& d = pl.d;
    if (!diffusion2_accepts(d, pl)) return -1;
    int rc = d.eltype == SB200_F32 ? d2_try<float>(pl, src, dst, st) : d2_try<double>(pl, src, dst, st);
    if (rc == SB200_OK) set_kernel_name("stream3d2_kernel");
    return rc;
}

}  // namespace sb
