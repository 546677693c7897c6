// small2d.cu — radius-1 2-D gathers on grids that fit the L2 (BASELINE configs[0] at the README's size: mean, Window(1), Float64
// 1000 x 1000, Remove(0); /root/reference/README.md:142-203 is the only configuration the reference publishes numbers for).
//
// Replaces gatherstencil_kernel! (src/gatherstencil.jl:105-109) + the neighbour read path (src/array.jl:91-138) for small arrays.
// The streaming kernels (stream2d.cuh) are built for grids far larger than the caches: 296 persistent CTAs, a TMA ring, mbarrier
// hand-offs — on a 1000 x 1000 grid every CTA streams ~3400 cells and the sweep lasts 6.4 us (tools/mean1000_probe.py). The question
// this file answers is whether that is pipeline start-up. Here nothing is staged: a thread owns 16 bytes of one row (two Float64 / four Float32 cells), reads its
// three source rows straight from L1 / L2 (one 128-bit load per row plus the two neighbour cells), folds the taps in the
// reference's offset order (bit-identical to the streaming kernels) and stores 16 bytes; ~2000 small CTAs cover the
// grid in under two waves. Threads whose cells touch the array edge resolve every neighbour through the boundary rule
// (Remove padval / Wrap / Reflect, or the Halo ring read straight through).
// Window(1), Moore(1), VonNeumann(1) x sum / mean / minimum / maximum x Float32 / Float64, whole-array sweeps up to
// SB200_SMALL2D_MAX_CELLS cells — an EXPERIMENT that is off by default (0): it measured no faster than the streaming kernel, see
// try_small2d; everything stays with the streaming kernels unless the variable is set.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"

namespace sb {

constexpr int SM2_BX = 64, SM2_BY = 4;

__host__ __device__ constexpr bool sm2_has(int shape, int dx, int dy) {
    const int manh = (dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy);
    return shape == SB200_WINDOW ? true : shape == SB200_MOORE ? manh != 0 : manh == 1;   // VonNeumann(1)
}
__host__ __device__ constexpr int sm2_count(int shape) { return shape == SB200_WINDOW ? 9 : shape == SB200_MOORE ? 8 : 4; }

template <typename T> struct Sm2Params {
    const T* src;
    T* dst;
    long long spitch, dpitch;
    int W, H;
    int soff0, soff1, doff0, doff1;
    int bc0, bc1;
    int vec;   // rows and bases 16-byte aligned: 128-bit loads / stores
    T pad;
};

template <typename T, int RED> __device__ __forceinline__ T sm2_op(T acc, T v) {
    if (RED == SB200_MAX) return jl_max(acc, v);
    if (RED == SB200_MIN) return jl_min(acc, v);
    return add_rn(acc, v);
}

template <typename T, int SHAPE, int RED>
__global__ void __launch_bounds__(SM2_BX * SM2_BY) small2d_kernel(const __grid_constant__ Sm2Params<T> p) {
    constexpr int VX = 16 / (int)sizeof(T);
    const int x = (blockIdx.x * SM2_BX + threadIdx.x) * VX;
    const int y = blockIdx.y * SM2_BY + threadIdx.y;
    if (x >= p.W || y >= p.H) return;
    T v[3][VX + 2];
    // a thread is "inner" when its VX + 2 columns and three rows exist as they are (ring reads count as inner)
    const bool inner_x = (p.soff0 > 0 || x >= 1) && (p.soff0 > 0 ? x + VX <= p.W : x + VX + 1 <= p.W);
    const bool inner_y = p.soff1 > 0 || (y >= 1 && y + 1 < p.H);
    if (inner_x && inner_y) {
        const T* __restrict__ s = p.src + (long long)(y - 1 + p.soff1) * p.spitch + (x + p.soff0);
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const T* __restrict__ row = s + r * p.spitch;
            v[r][0] = __ldg(row - 1);
            if (p.vec) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(row));
                memcpy(&v[r][1], &q, 16);
            } else {
#pragma unroll
                for (int i = 0; i < VX; i++) v[r][1 + i] = __ldg(row + i);
            }
            v[r][VX + 1] = __ldg(row + VX);
        }
    } else {   // edge threads: every neighbour through the boundary rule (src/array.jl:101-138); cells beyond W are never stored
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const long long jy = y - 1 + r;
            const long long qy = p.soff1 > 0 ? jy + p.soff1 : bounded(jy, p.H, p.bc1);
#pragma unroll
            for (int i = 0; i < VX + 2; i++) {
                const long long jx = x - 1 + i;
                T val = p.pad;
                if (jx <= p.W) {   // column W is the right neighbour of the last cell
                    const long long qx = p.soff0 > 0 ? jx + p.soff0 : bounded(jx, p.W, p.bc0);
                    if (qx >= 0 && qy >= 0) val = __ldg(p.src + qy * p.spitch + qx);
                }
                v[r][i] = val;
            }
        }
    }
    T out[VX];
#pragma unroll
    for (int c = 0; c < VX; c++) {
        T acc = T(0);
        bool first = true;
#pragma unroll
        for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
                if (!sm2_has(SHAPE, dx, dy)) continue;
                const T t = v[dy + 1][c + 1 + dx];
                acc = first ? t : sm2_op<T, RED>(acc, t);
                first = false;
            }
        out[c] = RED == SB200_MEAN ? div_rn(acc, (T)sm2_count(SHAPE)) : acc;
    }
    T* __restrict__ d = p.dst + (long long)(y + p.doff1) * p.dpitch + (x + p.doff0);
    if (p.vec && x + VX <= p.W) {
        uint4 q;
        memcpy(&q, out, 16);
        *reinterpret_cast<uint4*>(d) = q;
    } else {
#pragma unroll
        for (int c = 0; c < VX; c++)
            if (x + c < p.W) d[c] = out[c];
    }
}

template <typename T, int SHAPE> static int sm2_launch_red(const Sm2Params<T>& p, int red, dim3 grid, cudaStream_t st) {
    const dim3 block(SM2_BX, SM2_BY);
    switch (red) {
    case SB200_SUM: small2d_kernel<T, SHAPE, SB200_SUM><<<grid, block, 0, st>>>(p); break;
    case SB200_MEAN: small2d_kernel<T, SHAPE, SB200_MEAN><<<grid, block, 0, st>>>(p); break;
    case SB200_MIN: small2d_kernel<T, SHAPE, SB200_MIN><<<grid, block, 0, st>>>(p); break;
    case SB200_MAX: small2d_kernel<T, SHAPE, SB200_MAX><<<grid, block, 0, st>>>(p); break;
    default: return -1;
    }
    return SB200_OK;
}

template <typename T> static int sm2_try(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const sb200_desc& d = pl.d;
    constexpr int VX = 16 / (int)sizeof(T);
    Sm2Params<T> p;
    p.src = (const T*)src; p.dst = (T*)dst;
    p.spitch = d.src_ext[0]; p.dpitch = d.dst_ext[0];
    p.W = (int)d.size[0]; p.H = (int)d.size[1];
    p.soff0 = d.src_off[0]; p.soff1 = d.src_off[1]; p.doff0 = d.dst_off[0]; p.doff1 = d.dst_off[1];
    p.bc0 = d.boundary[0]; p.bc1 = d.boundary[1];
    memcpy(&p.pad, &d.padval_bits, sizeof(T));
    const size_t es = sizeof(T);
    p.vec = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0) && (p.spitch * es) % 16 == 0 && (p.dpitch * es) % 16 == 0 &&
            (p.soff0 * es) % 16 == 0 && (p.doff0 * es) % 16 == 0;
    const dim3 grid((unsigned)((p.W + SM2_BX * VX - 1) / (SM2_BX * VX)), (unsigned)((p.H + SM2_BY - 1) / SM2_BY));
    int rc = -1;
    switch (pl.shape_tag) {
    case SB200_WINDOW: rc = sm2_launch_red<T, SB200_WINDOW>(p, d.reducer, grid, st); break;
    case SB200_MOORE: rc = sm2_launch_red<T, SB200_MOORE>(p, d.reducer, grid, st); break;
    case SB200_VONNEUMANN: rc = sm2_launch_red<T, SB200_VONNEUMANN>(p, d.reducer, grid, st); break;
    default: break;
    }
    return rc;
}

int try_small2d(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const sb200_desc& d = pl.d;
    if (d.ndim != 2 || d.radius != 1 || pl.shape_ndim != 2) return -1;
    if (d.eltype != SB200_F32 && d.eltype != SB200_F64) return -1;
    if (d.reducer != SB200_SUM && d.reducer != SB200_MEAN && d.reducer != SB200_MIN && d.reducer != SB200_MAX) return -1;
    if (pl.shape_tag != SB200_WINDOW && pl.shape_tag != SB200_MOORE && pl.shape_tag != SB200_VONNEUMANN) return -1;
    if (d.flags & (SB200_FLAG_NO_TMA | SB200_FLAG_FORCE_GENERIC | SB200_FLAG_STEP_MASK)) return -1;
    // OPT-IN (default 0 = off). Measured r02ag on the 1000 x 1000 Float64 mean: 6.71 us per sweep in a replayed CUDA graph against 6.44 us
    // for the streaming kernel, 8.2 us for both as plain back-to-back launches — the floor of a 16 MB sweep on this part is launch +
    // first-wave latency, not the kernel's structure, so the direct kernel buys nothing and the streaming kernel stays the default.
    const char* e = getenv("SB200_SMALL2D_MAX_CELLS");   // read per call, but only for sweeps that passed every test above
    const long long max_cells = e ? atoll(e) : 0;
    if (d.size[0] * d.size[1] > max_cells || d.size[0] < 2 || d.size[1] < 2) return -1;
    if (pl.dd.lo[0] != 0 || pl.dd.lo[1] != 0 || pl.dd.n[0] != d.size[0] || pl.dd.n[1] != d.size[1]) return -1;   // whole-array sweeps
    for (int a = 0; a < 2; a++)
        if (d.src_off[a] == 0 && d.boundary[a] == SB200_USE) return -1;
    if (g_mirror.ptr) return -1;
    const int rc = d.eltype == SB200_F32 ? sm2_try<float>(pl, src, dst, st) : sm2_try<double>(pl, src, dst, st);
    if (rc != SB200_OK) return rc;
    SB_LAUNCH_CHECK();
    set_kernel_name("small2d_kernel");
    return SB200_OK;
}

}  // namespace sb
