// gather_stream3d.cu — 3-D gather for ANY offset table of radius <= 2 (Window(1,3) / Moore(1,3) / Cross / Circle /
// Positional ... in three dimensions) as a TMA-fed 2.5-D streaming kernel with a run-time tap table.
//
// Replaces gatherstencil_kernel! (src/gatherstencil.jl:105-109) + the neighbour read path (src/array.jl:91-138) for the
// 3-D sweeps stream3d.cu (VonNeumann(1,3) only) does not instantiate. A CTA owns an (x, y) tile of G3_BXB bytes x
// G3_TY rows and marches along z; the 32 lanes of a producer warp issue cp.async.bulk (UBLKCP) copies of one z-plane of
// the tile (+R halo rows above / below, +halo cells left / right) per stage into a ring of shared-memory planes, and
// resolve Wrap / Reflect / ring rows and planes and the Wrap halo of axis 0 by choosing source addresses. The 2R+1
// planes an output plane needs stay resident in the ring, so every cell is read from HBM once (+ tile halos). Lane l
// owns cells l, l+32, ... of the tile row, so the shared-memory read of a tap with any offset is conflict-free and
// every global store is a coalesced 128-byte access. Taps fold in table order (the reference's offset order).
// The R cells next to each end of axis 0 under Remove / Reflect go to gather_generic as two thin bands.
// Algorithmic traffic: sizeof(T) read + sizeof(T) written per cell; the fold itself is shared-memory bound for large L
// (one LDS per tap and cell).
#include <algorithm>
#include <type_traits>
#include "common.cuh"
#include "tma.cuh"

namespace sb {

constexpr int G3_WARPS = 8;
// EXPERIMENT (default 1 = the measured kernel; other values have not been on a GPU yet): producer warps, the rows of a plane
// dealt round-robin. This kernel copies 544-byte rows — one bulk copy (~14 issue slots of lane-by-lane serialisation) per 512
// bytes of tile, twice stream3d's rate, and stream3d went from 0.79 to 0.94 of the HBM roofline with a second producer warp
// (lesson 3).
#ifndef SB200_G3_PRODUCERS
#define SB200_G3_PRODUCERS 1
#endif
constexpr int G3_PRODUCERS = SB200_G3_PRODUCERS;
constexpr int G3_THREADS = (G3_WARPS + G3_PRODUCERS) * 32;
constexpr int G3_BXB = 512;                   // tile width in bytes: one lane-strided warp row
constexpr int G3_TY = 16;                     // tile height in rows; warp w owns rows w and w + 8
constexpr int G3_RT = G3_TY / G3_WARPS;
constexpr int G3_MAXR = 2;
constexpr int G3_LEFT = 128;                  // margin: global and shared addresses of the main copy agree mod 128
constexpr int G3_ROWB = G3_LEFT + G3_BXB + 128;
constexpr int G3_ROWS = G3_TY + 2 * G3_MAXR;
constexpr int G3_PLANE = G3_ROWS * G3_ROWB;
constexpr int G3_NS = 7;                      // ring of planes: 2R+1 resident + prefetch
constexpr int G3_HDR = 512;
constexpr int G3_TAB = 128;                   // taps (Window(2,3) = 125)
constexpr int G3_SMEM = G3_HDR + G3_TAB * 16 + G3_NS * G3_PLANE;

// byte shift inside the plane tile (o1 rows + o0 cells), ring distance of the tap's plane (R + o2), the offsets, weight
template <typename T> struct G3Tap { short boff, d2, o1, o2; T w; };

template <typename T> struct G3Params {
    const T* src;
    T* dst;
    long long sp1, sp2, dp1, dp2;  // pitches (elements) of axes 1 and 2
    int X, Y, Z;
    int so1, so2, do0, do1, do2;
    int bc0, bc1, bc2;
    T pad, alpha;
    int x_lo, x_hi;                // cells handled here along axis 0
    int z_lo, zn;                  // output planes [z_lo, z_lo + zn)
    int ntx, nty, nzruns;
    int R, L;
    const int* offs;
    const T* weights;
};

__device__ __forceinline__ long long g3_map(int r, int n, int off, int bc) {
    if (off > 0) return (long long)r + off;
    if (r >= 0 && r < n) return r;
    if (bc == SB200_WRAP) return r < 0 ? r + n : r - n;
    if (bc == SB200_REFLECT) return r < 0 ? -r : 2 * (n - 1) - r;
    return -1;
}

template <typename T, int RED> __device__ __forceinline__ T g3_fold(T acc, T v, T w) {
    if (RED == SB200_MAX) return jl_max(acc, v);
    if (RED == SB200_MIN) return jl_min(acc, v);
    if (RED == SB200_KERNELDOT) return add_rn(acc, mul_rn(v, w));
    return add_rn(acc, v);
}
template <typename T, int RED> __device__ __forceinline__ T g3_first(T v, T w) {
    if (RED == SB200_KERNELDOT) return add_rn(T(0), mul_rn(v, w));
    return v;
}

template <typename T, int RED>
__global__ void __launch_bounds__(G3_THREADS, 2) gather_stream3d_kernel(const __grid_constant__ G3Params<T> p) {
    constexpr int VX = 16 / (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + G3_NS;
    unsigned char* tabraw = smem + G3_HDR;
    unsigned char* ring = smem + G3_HDR + G3_TAB * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = p.R, L = p.L;
    if (threadIdx.x == 0) {
        for (int s = 0; s < G3_NS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], G3_WARPS); }
        mbar_fence_init();
    }
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
        G3Tap<T>* t = reinterpret_cast<G3Tap<T>*>(tabraw + k * 16);
        const int o0 = p.offs[3 * k], o1 = p.offs[3 * k + 1], o2 = p.offs[3 * k + 2];
        t->boff = (short)(o1 * G3_ROWB + o0 * (int)sizeof(T));
        t->d2 = (short)(R + o2);
        t->o1 = (short)o1; t->o2 = (short)o2;
        t->w = p.weights ? p.weights[k] : T(0);
    }
    __syncthreads();
    const int ntiles = p.ntx * p.nty;
    const int ntasks = ntiles * p.nzruns;
    const int Xb = p.X * (int)sizeof(T);
    const int HLB = ((R * (int)sizeof(T) + 15) / 16) * 16;
    const bool pad1 = p.so1 == 0 && p.bc1 == SB200_REMOVE, pad2 = p.so2 == 0 && p.bc2 == SB200_REMOVE;
    unsigned kb = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int tile = task % ntiles, zrun = task / ntiles;
        const int x0b = (tile % p.ntx) * G3_BXB, y0 = (tile / p.ntx) * G3_TY;
        const int wbytes = min(G3_BXB, Xb - x0b);
        const int z0 = p.z_lo + (int)((long long)p.zn * zrun / p.nzruns);
        const int z1 = p.z_lo + (int)((long long)p.zn * (zrun + 1) / p.nzruns);
        const int nout = z1 - z0;
        const int nst = nout + 2 * R;  // stage i holds source plane z0 - R + i
        if (G3_PRODUCERS == 1 ? warp == G3_WARPS : warp >= G3_WARPS) {
            // ---------------- producer warps: lane j of producer w copies row P*j+w of the plane tile (logical row y0 - R + P*j+w) ----------------
            const int pw = G3_PRODUCERS == 1 ? 0 : warp - G3_WARPS;
            const bool l_in = x0b > 0, r_in = x0b + wbytes < Xb;
            const bool l_wrap = !l_in && p.bc0 == SB200_WRAP, r_wrap = !r_in && p.bc0 == SB200_WRAP;
            const int r_in_bytes = r_in ? min(HLB, Xb - (x0b + wbytes)) : 0;
            const int mstart = x0b - (l_in ? HLB : 0), mdst = G3_LEFT - (l_in ? HLB : 0);
            const unsigned mlen = wbytes + (l_in ? HLB : 0) + r_in_bytes;
            const int wrap_bytes = min(HLB, Xb);
            const unsigned rowbytes = mlen + (l_wrap ? wrap_bytes : 0) + (r_wrap ? wrap_bytes : 0);
            long long yrow = -1;
            if (lane < G3_TY + 2 * R) {
                const int y = y0 - R + lane;
                if (y < p.Y + R) yrow = g3_map(y, p.Y, p.so1, p.bc1);
            }
            const unsigned nrows = __popc(__ballot_sync(0xffffffffu, yrow >= 0));   // all rows of the stage (expect_tx)
            int srow_i = lane;
            if constexpr (G3_PRODUCERS > 1) {   // this producer's row
                srow_i = G3_PRODUCERS * lane + pw;
                yrow = -1;
                if (srow_i < G3_TY + 2 * R) {
                    const int y = y0 - R + srow_i;
                    if (y < p.Y + R) yrow = g3_map(y, p.Y, p.so1, p.bc1);
                }
            }
            for (int i = 0; i < nst; i++) {
                const unsigned k = kb + i;
                const int slot = k % G3_NS;
                const long long zpl = g3_map(z0 - R + i, p.Z, p.so2, p.bc2);
                if (lane == 0) {
                    mbar_wait_producer(&empty[slot], ((k / G3_NS) & 1) ^ 1);
                    if (pw == 0) mbar_arrive_expect_tx(&full[slot], zpl >= 0 ? nrows * rowbytes : 0u);
                }
                __syncwarp();
                if (zpl >= 0 && yrow >= 0) {
                    const unsigned char* g = reinterpret_cast<const unsigned char*>(p.src + zpl * p.sp2 + yrow * p.sp1);
                    unsigned char* srow = ring + slot * G3_PLANE + srow_i * G3_ROWB;
                    bulk_g2s(srow + mdst, g + mstart, mlen, &full[slot]);
                    if (l_wrap) bulk_g2s(srow + G3_LEFT - wrap_bytes, g + Xb - wrap_bytes, wrap_bytes, &full[slot]);
                    if (r_wrap) bulk_g2s(srow + G3_LEFT + wbytes, g, wrap_bytes, &full[slot]);
                }
            }
            kb += nst;
            continue;
        }
        // ---------------- consumers ----------------
        const int gx0 = x0b / (int)sizeof(T) + lane;            // global index of the lane's first cell along axis 0
        bool inr[VX];
#pragma unroll
        for (int v = 0; v < VX; v++) {
            const int x = gx0 + v * 32;
            inr[v] = x >= p.x_lo && x < p.x_hi;
        }
        for (int i = 0; i < 2 * R; i++) {
            const unsigned k = kb + i;
            mbar_wait(&full[k % G3_NS], (k / G3_NS) & 1);
        }
        unsigned slot_t = kb % G3_NS;                            // slot of stage t (source plane z - R)
        unsigned slot_n = (kb + 2 * R) % G3_NS;                  // slot of stage t + 2R (source plane z + R)
        unsigned par_n = ((kb + 2 * R) / G3_NS) & 1;
        // this lane's first cell of tile row 0 (shared-memory row R) in slot 0
        const unsigned char* mine0 = ring + R * G3_ROWB + G3_LEFT + lane * (int)sizeof(T);
        for (int t = 0; t < nout; t++) {
            const int z = z0 + t;
            mbar_wait(&full[slot_n], par_n);
            const bool zplain = !pad2 || (z - R >= 0 && z + R < p.Z);
#pragma unroll
            for (int rr = 0; rr < G3_RT; rr++) {
                const int ry = warp + rr * G3_WARPS;             // tile row
                const int y = y0 + ry;
                if (y >= p.Y) continue;                           // ragged last tile (warp-uniform)
                const unsigned char* mine = mine0 + ry * G3_ROWB;
                T acc[VX];
                if (zplain && (!pad1 || (y - R >= 0 && y + R < p.Y))) {
                    {
                        const G3Tap<T> tp = *reinterpret_cast<const G3Tap<T>*>(tabraw);
                        unsigned sl = slot_t + tp.d2;
                        sl = sl >= G3_NS ? sl - G3_NS : sl;
                        const T* srow = reinterpret_cast<const T*>(mine + sl * G3_PLANE + tp.boff);
#pragma unroll
                        for (int v = 0; v < VX; v++) acc[v] = g3_first<T, RED>(srow[v * 32], tp.w);
                    }
#pragma unroll 4
                    for (int q = 1; q < L; q++) {
                        const G3Tap<T> tp = *reinterpret_cast<const G3Tap<T>*>(tabraw + q * 16);
                        unsigned sl = slot_t + tp.d2;
                        sl = sl >= G3_NS ? sl - G3_NS : sl;
                        const T* srow = reinterpret_cast<const T*>(mine + sl * G3_PLANE + tp.boff);
#pragma unroll
                        for (int v = 0; v < VX; v++) acc[v] = g3_fold<T, RED>(acc[v], srow[v * 32], tp.w);
                    }
                } else {
                    // Remove on axis 1 / 2 next to the array faces: rows / planes outside the array read padval
                    for (int q = 0; q < L; q++) {
                        const G3Tap<T> tp = *reinterpret_cast<const G3Tap<T>*>(tabraw + q * 16);
                        const int yy = y + tp.o1, zz = z + tp.o2;
                        const bool oob = (pad1 && (yy < 0 || yy >= p.Y)) || (pad2 && (zz < 0 || zz >= p.Z));
                        unsigned sl = slot_t + tp.d2;
                        sl = sl >= G3_NS ? sl - G3_NS : sl;
                        const T* srow = reinterpret_cast<const T*>(mine + sl * G3_PLANE + tp.boff);
#pragma unroll
                        for (int v = 0; v < VX; v++) {
                            const T x = oob ? p.pad : srow[v * 32];
                            acc[v] = q == 0 ? g3_first<T, RED>(x, tp.w) : g3_fold<T, RED>(acc[v], x, tp.w);
                        }
                    }
                }
                if constexpr (RED == SB200_MEAN) {
#pragma unroll
                    for (int v = 0; v < VX; v++) acc[v] = div_rn(acc[v], (T)L);
                }
                if constexpr (RED == SB200_DIFFUSION) {
                    unsigned sl = slot_t + R;
                    sl = sl >= G3_NS ? sl - G3_NS : sl;
                    const T* crow = reinterpret_cast<const T*>(mine + sl * G3_PLANE);
#pragma unroll
                    for (int v = 0; v < VX; v++) {
                        const T c = crow[v * 32];
                        acc[v] = add_rn(c, mul_rn(p.alpha, sub_rn(acc[v], mul_rn((T)L, c))));
                    }
                }
                T* drow = p.dst + (long long)(z + p.do2) * p.dp2 + (long long)(y + p.do1) * p.dp1 + p.do0 + gx0;
#pragma unroll
                for (int v = 0; v < VX; v++)
                    if (inr[v]) drow[v * 32] = acc[v];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot_t]);  // source plane z-R is done
            if (++slot_t == G3_NS) slot_t = 0;
            if (++slot_n == G3_NS) { slot_n = 0; par_n ^= 1; }
        }
        __syncwarp();
        if (lane == 0)
            for (int i = nout; i < nst; i++) mbar_arrive(&empty[(kb + i) % G3_NS]);
        kb += nst;
    }
}

template <typename T, int RED> static int g3_launch(G3Params<T>& p, cudaStream_t st) {
    static thread_local int cfg_dev = -1, ctas_per_sm = 0;
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (dev != cfg_dev) {
        SB_CUDA(cudaFuncSetAttribute(gather_stream3d_kernel<T, RED>, cudaFuncAttributeMaxDynamicSharedMemorySize, G3_SMEM));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_stream3d_kernel<T, RED>, G3_THREADS, G3_SMEM) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        ctas_per_sm = per_sm;
        cfg_dev = dev;
    }
    const long long ctas = (long long)ctas_per_sm * num_sms();
    const long long ntiles = (long long)p.ntx * p.nty;
    // z-runs: the 2R re-read planes per run against the idle tail of the last wave of tasks
    int best = 1;
    double best_cost = 1e300;
    for (int nz = 1; nz <= 64 && nz <= std::max(1, p.zn / (4 * p.R)); nz++) {
        const long long tasks = ntiles * nz;
        const long long waves = (tasks + ctas - 1) / ctas;
        const double cost = (double)waves * ((double)p.zn / nz + 2.0 * p.R);
        if (cost < best_cost * 0.999) { best_cost = cost; best = nz; }
    }
    p.nzruns = best;
    const long long grid = std::min<long long>(ctas, ntiles * p.nzruns);
    gather_stream3d_kernel<T, RED><<<(unsigned)grid, G3_THREADS, G3_SMEM, st>>>(p);
    SB_LAUNCH_CHECK();
    return SB200_OK;
}

template <typename T> static int g3_try(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const sb200_desc& d = pl.d;
    const DevDesc& dd = pl.dd;
    const int R = d.radius, L = d.noffsets;
    if (R < 1 || R > G3_MAXR || L < 1 || L > G3_TAB) return -1;
    if (d.src_off[0] != 0 || d.boundary[0] == SB200_USE) return -1;                 // axis 0 must be unpadded
    for (int a = 1; a < 3; a++)
        if (d.src_off[a] == 0 && d.boundary[a] == SB200_USE) return -1;
    const long long Xb = d.size[0] * (long long)sizeof(T);
    if (Xb % 16 || Xb < 64 || d.size[0] <= 4 * R) return -1;
    if (Xb > G3_BXB && (Xb % G3_BXB) != 0 && (Xb % G3_BXB) < 32) return -1;         // last tile narrower than a wrap halo
    if ((d.src_ext[0] * sizeof(T)) % 16 || ((uintptr_t)src & 15) || ((uintptr_t)dst % sizeof(T))) return -1;
    if (d.size[0] > (1 << 28) || d.size[1] > (1 << 28) || d.size[2] > (1 << 28) || R >= d.size[1] || R >= d.size[2]) return -1;
    if (dd.lo[0] != 0 || dd.n[0] != d.size[0] || dd.lo[1] != 0 || dd.n[1] != d.size[1]) return -1;  // z regions only
    if (dd.n[2] == 0) return SB200_OK;
    G3Params<T> p;
    p.src = (const T*)src; p.dst = (T*)dst;
    p.sp1 = d.src_ext[0]; p.sp2 = d.src_ext[0] * d.src_ext[1];
    p.dp1 = d.dst_ext[0]; p.dp2 = d.dst_ext[0] * d.dst_ext[1];
    p.X = (int)d.size[0]; p.Y = (int)d.size[1]; p.Z = (int)d.size[2];
    p.so1 = d.src_off[1]; p.so2 = d.src_off[2];
    p.do0 = d.dst_off[0]; p.do1 = d.dst_off[1]; p.do2 = d.dst_off[2];
    p.bc0 = d.boundary[0]; p.bc1 = d.boundary[1]; p.bc2 = d.boundary[2];
    memcpy(&p.pad, &d.padval_bits, sizeof(T));
    p.alpha = (T)d.alpha;
    const int band = d.boundary[0] == SB200_WRAP ? 0 : R;
    p.x_lo = band; p.x_hi = p.X - band;
    p.z_lo = (int)dd.lo[2]; p.zn = (int)dd.n[2];
    p.ntx = (int)((Xb + G3_BXB - 1) / G3_BXB);
    p.nty = (int)((d.size[1] + G3_TY - 1) / G3_TY);
    p.R = R; p.L = L;
    p.offs = dd.offs; p.weights = d.reducer == SB200_KERNELDOT ? (const T*)dd.weights : nullptr;
    int rc = -1;
    switch (d.reducer) {
    case SB200_SUM: rc = g3_launch<T, SB200_SUM>(p, st); break;
    case SB200_MIN: rc = g3_launch<T, SB200_MIN>(p, st); break;
    case SB200_MAX: rc = g3_launch<T, SB200_MAX>(p, st); break;
    case SB200_KERNELDOT: rc = g3_launch<T, SB200_KERNELDOT>(p, st); break;
    case SB200_MEAN:
        if constexpr (std::is_floating_point<T>::value) rc = g3_launch<T, SB200_MEAN>(p, st);
        break;
    case SB200_DIFFUSION:
        if constexpr (std::is_floating_point<T>::value) rc = g3_launch<T, SB200_DIFFUSION>(p, st);
        break;
    default: break;
    }
    if (rc != SB200_OK) return rc;
    if (band > 0) {  // the two edge bands of axis 0 (all rows and planes of the region)
        Plan edge = pl;
        edge.dd.lo[0] = 0; edge.dd.n[0] = band;
        if ((rc = launch_generic_gather(edge, src, dst, st))) return rc;
        edge.dd.lo[0] = p.X - band;
        if ((rc = launch_generic_gather(edge, src, dst, st))) return rc;
    }
    set_kernel_name("gather_stream3d_kernel");
    return SB200_OK;
}

int try_gather_stream3d(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const sb200_desc& d = pl.d;
    if ((d.flags & SB200_FLAG_NO_TMA) || d.ndim != 3 || d.out_eltype != d.eltype) return -1;
    switch (d.eltype) {
    case SB200_F32: return g3_try<float>(pl, src, dst, st);
    case SB200_F64: return g3_try<double>(pl, src, dst, st);
    case SB200_I32: return g3_try<int32_t>(pl, src, dst, st);
    case SB200_I64: return g3_try<int64_t>(pl, src, dst, st);
    default: return -1;
    }
}

}  // namespace sb
