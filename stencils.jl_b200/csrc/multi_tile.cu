// multi_tile.cu — multi-array gathers in ONE pass: dest = c_1 g_1(hood(A_1)) + c_2 g_2(hood(A_2)) + ... for 2-D arrays.
//
// Replaces gatherstencil!(f, dest, A1, A2, ...) (src/gatherstencil.jl:84-88, 112-113) for the user functions of
// test/array.jl:312-383 (a linear combination of per-argument reducers, evaluated left to right, every operation rounded
// separately). sb200_gather_multi used to run one streaming sweep per argument into a scratch parent plus a combine kernel per
// argument: 7 array transits for two arguments where 3 are needed (2 reads + 1 write). Here a CTA owns a tile of MT_TX x MT_TY
// destination cells; per argument it stages the tile plus the argument's own radius-R halo in shared memory — boundary rule,
// Halo ring and padval of THAT argument resolved at load time (src/array.jl:91-138) —, every thread folds the argument's taps in
// table order for its four cells (the reference's offset order: bit-identical to the per-argument kernels)
// and adds the term to its running results; the destination is written once. Every source cell is read from HBM once plus
// the tile halo (1.08x for R = 1 at 128 x 32; neighbouring CTAs find it in L2), no scratch parent is needed.
// 3-D arrays, integer element types and reducers outside the menu below keep the sweep-per-argument path (api.cu) — which is also
// still the DEFAULT: see try_multi_tile2d for the measurement.
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "common.cuh"

namespace sb {

constexpr int MT_TX = 128, MT_TY = 32, MT_MAXR = 4, MT_THREADS = 256;
constexpr int MT_RG = MT_TY / (MT_THREADS / 32);   // rows per warp: a thread owns MT_RG x 4 cells (rows 8 apart, cells 32 apart)
constexpr int MT_TILE = (MT_TX + 2 * MT_MAXR) * (MT_TY + 2 * MT_MAXR);
constexpr int MT_MAXL = 128;   // taps per argument

struct MtTerm {
    DevDesc d;
    const void* src;
    double coef;
    int has_coef;
};
struct MtParams {
    int nterms;
    long long n0, n1;       // logical size
    long long dstr1;        // dest parent: elements per row of axis 1
    int doff0, doff1;
    void* dst;
    MtTerm t[SB200_MAX_TERMS];
};

template <typename T> __device__ __forceinline__ T mt_pad(unsigned long long bits) {
    T v;
    memcpy(&v, &bits, sizeof(T));
    return v;
}

// parent index of logical index j on axis a (ring read, or the boundary rule on the fly), -1 = padval
__device__ __forceinline__ long long mt_index(const DevDesc& d, int a, long long j) {
    if (d.soff[a] > 0) return j + d.soff[a];
    return bounded(j, d.size[a], d.bc[a]);
}

// The taps of one argument over the four cells of a thread (cells 32 apart: every tap is a conflict-free LDS), strict left fold
// in table order; the four chains are independent.
template <typename T, int RED>
__device__ __forceinline__ void mt_fold(const T* __restrict__ c, const int* __restrict__ toff, const T* __restrict__ tw, int L, T (&acc)[4]) {
    {
        const int o = toff[0];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const T v = c[o + 32 * i];
            acc[i] = RED == SB200_KERNELDOT ? add_rn(T(0), mul_rn(v, tw[0])) : v;
        }
    }
#pragma unroll 4
    for (int k = 1; k < L; k++) {
        const int o = toff[k];
        const T wk = RED == SB200_KERNELDOT ? tw[k] : T(0);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const T v = c[o + 32 * i];
            if (RED == SB200_MAX) acc[i] = jl_max(acc[i], v);
            else if (RED == SB200_MIN) acc[i] = jl_min(acc[i], v);
            else if (RED == SB200_KERNELDOT) acc[i] = add_rn(acc[i], mul_rn(v, wk));
            else acc[i] = add_rn(acc[i], v);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(MT_THREADS) multi_tile2d_kernel(const __grid_constant__ MtParams p) {
    __shared__ T tile[MT_TILE];
    __shared__ int toff[MT_MAXL];   // tap -> offset inside the tile
    __shared__ T tw[MT_MAXL];       // kernelproduct weights
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n0 = (int)p.n0, n1 = (int)p.n1;
    const int tiles_x = (n0 + MT_TX - 1) / MT_TX, tiles_y = (n1 + MT_TY - 1) / MT_TY;
    for (int tid = blockIdx.x; tid < tiles_x * tiles_y; tid += gridDim.x) {
        const int x0 = (tid % tiles_x) * MT_TX, y0 = (tid / tiles_x) * MT_TY;
        T res[MT_RG][4];
        for (int j = 0; j < p.nterms; j++) {
            const DevDesc& d = p.t[j].d;
            const T* __restrict__ src = (const T*)p.t[j].src;
            const int R = d.R, W = MT_TX + 2 * R, H = MT_TY + 2 * R, L = d.L;
            const T pv = mt_pad<T>(d.padbits);
            __syncthreads();   // the previous argument's taps have been read
            for (int k = threadIdx.x; k < L; k += MT_THREADS) {
                toff[k] = d.offs[3 * k + 1] * W + d.offs[3 * k];
                tw[k] = d.weights ? ((const T*)d.weights)[k] : T(0);
            }
            // a tile whose halo lies inside the array needs no boundary rule: parent index = logical + ring offset
            const bool inside = x0 - R >= 0 && x0 + MT_TX + R <= n0 && y0 - R >= 0 && y0 + MT_TY + R <= n1;
#pragma unroll 2
            for (int ly = warp; ly < H; ly += MT_THREADS / 32) {
                const int gy = y0 - R + ly;
                T* trow = tile + ly * W;
                if (inside) {   // all loads of the row in flight before the first store (W <= 136: five per lane)
                    const T* __restrict__ srow = src + (long long)(gy + d.soff[1]) * d.sstr[1] + (x0 - R + d.soff[0]);
                    T v[5];
#pragma unroll
                    for (int q = 0; q < 5; q++) v[q] = lane + 32 * q < W ? __ldg(srow + lane + 32 * q) : T(0);
#pragma unroll
                    for (int q = 0; q < 5; q++)
                        if (lane + 32 * q < W) trow[lane + 32 * q] = v[q];
                } else {
                    const long long qy = (gy >= -R && gy < n1 + R) ? mt_index(d, 1, gy) : -1;   // rows a tap of this tile can reach
                    for (int lx = lane; lx < W; lx += 32) {
                        const int gx = x0 - R + lx;
                        T v = pv;
                        if (qy >= 0 && gx >= -R && gx < n0 + R) {
                            const long long qx = mt_index(d, 0, gx);
                            if (qx >= 0) v = __ldg(src + qy * d.sstr[1] + qx);
                        }
                        trow[lx] = v;
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int rg = 0; rg < MT_RG; rg++) {
                const T* c = &tile[(warp + 8 * rg + R) * W + lane + R];
                T acc[4];
                switch (d.reducer) {
                    case SB200_MAX: mt_fold<T, SB200_MAX>(c, toff, tw, L, acc); break;
                    case SB200_MIN: mt_fold<T, SB200_MIN>(c, toff, tw, L, acc); break;
                    case SB200_KERNELDOT: mt_fold<T, SB200_KERNELDOT>(c, toff, tw, L, acc); break;
                    default: mt_fold<T, SB200_SUM>(c, toff, tw, L, acc); break;   // SUM, MEAN, DIFFUSION
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    T g = acc[i];
                    if (d.reducer == SB200_MEAN) g = div_rn(g, (T)L);
                    if (d.reducer == SB200_DIFFUSION) {
                        const T ctr = c[32 * i];
                        g = add_rn(ctr, mul_rn((T)d.alpha, sub_rn(g, mul_rn((T)L, ctr))));
                    }
                    // f(h1, h2, ...) = c1 g1 + c2 g2 + ...: left to right, every operation rounded separately
                    const T term = p.t[j].has_coef ? mul_rn((T)p.t[j].coef, g) : g;
                    res[rg][i] = j == 0 ? term : add_rn(res[rg][i], term);
                }
            }
        }
#pragma unroll
        for (int rg = 0; rg < MT_RG; rg++) {
            const int y = y0 + warp + 8 * rg;
            if (y >= n1) continue;
            T* __restrict__ drow = (T*)p.dst + (long long)(y + p.doff1) * p.dstr1 + p.doff0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int x = x0 + lane + 32 * i;
                if (x < n0) drow[x] = res[rg][i];
            }
        }
    }
}

// plans[j] = the validated sweep of argument j (device tables inside). dry: only answer whether the combination is this kernel's
// (SB200_OK) or not (-1); the caller refreshes the Halo rings of the sources between the question and the launch.
int try_multi_tile2d(const Plan* const* plans, const sb200_term* terms, int nterms, void* dst, cudaStream_t st, bool dry) {
    const sb200_desc& d0 = plans[0]->d;
    if (d0.ndim != 2 || (d0.eltype != SB200_F32 && d0.eltype != SB200_F64)) return -1;
    // Opt-in (SB200_MULTI_SINGLE_PASS=1): measured r02ae / r02af on 16384^2 Float32, mean(Window(1)) + 0.5 sum(VonNeumann(1)): 2.10 ms per
    // call against 1.83 ms for the sweep-per-argument path — the run-time tap loops and the per-tile staging cost 123 instructions per
    // cell (ncu: issue slots 55 % busy, DRAM 20 %), more than the four extra array transits they save at 0.9 of the roofline each. The
    // kernel needs no scratch parent, which is its use today; a multi-ring version of gather_stream_kernel is what would win.
    if (!(getenv("SB200_MULTI_SINGLE_PASS") && atoi(getenv("SB200_MULTI_SINGLE_PASS")) == 1)) return -1;
    MtParams p;
    p.nterms = nterms;
    p.n0 = d0.size[0]; p.n1 = d0.size[1];
    p.dstr1 = d0.dst_ext[0];
    p.doff0 = d0.dst_off[0]; p.doff1 = d0.dst_off[1];
    p.dst = dst;
    for (int j = 0; j < nterms; j++) {
        const sb200_desc& d = plans[j]->d;
        const int red = d.reducer;
        if (red != SB200_SUM && red != SB200_MEAN && red != SB200_MIN && red != SB200_MAX && red != SB200_KERNELDOT && red != SB200_DIFFUSION) return -1;
        if (d.radius > MT_MAXR || d.noffsets < 1 || d.noffsets > MT_MAXL) return -1;
        if (plans[j]->dd.sstr[0] != 1 || d.size[0] >= (1LL << 30) || d.size[1] >= (1LL << 30)) return -1;
        if (d.flags & (SB200_FLAG_FORCE_GENERIC | SB200_FLAG_STEP_MASK)) return -1;
        for (int a = 0; a < 2; a++)
            if (d.src_off[a] > 0 && d.src_off[a] < d.radius) return -1;   // a ring thinner than the radius: not a Halo layout
        p.t[j].d = plans[j]->dd;
        p.t[j].src = terms[j].src_parent;
        p.t[j].coef = terms[j].coef;
        p.t[j].has_coef = terms[j].has_coef;
    }
    if (dry || p.n0 == 0 || p.n1 == 0) return SB200_OK;
    const long long tiles = ((p.n0 + MT_TX - 1) / MT_TX) * ((p.n1 + MT_TY - 1) / MT_TY);
    const unsigned blocks = (unsigned)std::min<long long>(tiles, (long long)num_sms() * 16);
    if (d0.eltype == SB200_F32) multi_tile2d_kernel<float><<<blocks, MT_THREADS, 0, st>>>(p);
    else multi_tile2d_kernel<double><<<blocks, MT_THREADS, 0, st>>>(p);
    SB_LAUNCH_CHECK();
    set_kernel_name("multi_tile2d_kernel");
    return SB200_OK;
}

}  // namespace sb
