// generic.cu — shape-agnostic kernels: any offset table, 1..3 dimensions, every boundary/padding/eltype/reducer.
//
//  gather_generic  : gatherstencil_kernel!  (src/gatherstencil.jl:105-109) with the read path of
//                    src/array.jl:91-138 resolved per neighbour. One thread per output cell, axis 0 fastest so
//                    a warp reads/writes contiguous cells; neighbour re-reads are served by L1/L2.
//  halo_kernel     : update_boundary!       (src/array.jl:195-239)
//  scatter_generic : scatterstencil!        (src/scatterstencil.jl:49-112) as a deterministic per-destination
//                    fold in the reference's (pass, column, row, k) order — no atomics.
// The specialised kernels (life.cu, tile2d.cu, diffusion3d.cu, scatter_fast.cu) take over the headline
// configurations; this file is the complete-coverage CUDA path (it is NOT a CPU fallback).
#include <type_traits>
#include "common.cuh"

namespace sb {

template <typename T, int RED> struct OutOf { using type = T; };
template <> struct OutOf<uint8_t, SB200_MEAN> { using type = double; };
template <> struct OutOf<int32_t, SB200_MEAN> { using type = double; };
template <> struct OutOf<int64_t, SB200_MEAN> { using type = double; };

template <typename T> __device__ __forceinline__ T pad_of(unsigned long long bits) {
    T v;
    memcpy(&v, &bits, sizeof(T));
    return v;
}

// IS_BOOL: sum/mean of Bool accumulate in Int64 (Base.reduce_first(+, ::Bool) = Int(x)).
template <typename T, int RED, bool IS_BOOL>
__global__ void __launch_bounds__(256) gather_generic(DevDesc p, const T* __restrict__ src, void* __restrict__ dstv) {
    const long long total = p.n[0] * p.n[1] * p.n[2];
    const T pv = pad_of<T>(p.padbits);
    const T* __restrict__ w = (const T*)p.weights;
    for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < total;
         id += (long long)gridDim.x * blockDim.x) {
        long long I[3];
        long long rem = id;
        I[0] = p.lo[0] + rem % p.n[0]; rem /= p.n[0];
        I[1] = p.lo[1] + rem % p.n[1]; rem /= p.n[1];
        I[2] = p.lo[2] + rem;
        long long cidx = 0, didx = 0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (a < p.ndim) {
                cidx += (I[a] + p.soff[a]) * p.sstr[a];
                didx += (I[a] + p.doff[a]) * p.dstr[a];
            }
        }
        const T c = src[cidx];
        // streaming left fold in offset order
        T acc = T(0);
        long long iacc = 0;
        int cnt = 0;
        for (int k = 0; k < p.L; k++) {
            long long idx = 0;
            bool oob = false;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                if (a < p.ndim) {
                    long long j = I[a] + __ldg(&p.offs[3 * k + a]);
                    long long q;
                    if (p.soff[a] > 0) q = j + p.soff[a];                       // ring read (Halo)
                    else { q = bounded(j, p.size[a], p.bc[a]); if (q < 0) { oob = true; q = 0; } }
                    idx += q * p.sstr[a];
                }
            }
            const T v = oob ? pv : __ldg(&src[idx]);
            if (RED == SB200_SUM || RED == SB200_MEAN || RED == SB200_DIFFUSION) {
                if (IS_BOOL) iacc += (long long)v;
                else acc = (k == 0) ? v : add_rn(acc, v);
            } else if (RED == SB200_MAX) {
                acc = (k == 0) ? v : jl_max(acc, v);
            } else if (RED == SB200_MIN) {
                acc = (k == 0) ? v : jl_min(acc, v);
            } else if (RED == SB200_KERNELDOT) {
                acc = add_rn(acc, mul_rn(v, w[k]));
            } else if (RED == SB200_LIFE) {
                cnt += (v != T(0));
            }
        }
        if (RED == SB200_SUM) {
            if (IS_BOOL) ((long long*)dstv)[didx] = iacc;
            else ((T*)dstv)[didx] = acc;
        } else if (RED == SB200_MEAN) {
            if constexpr (sizeof(T) == 4 && !IS_BOOL && std::is_floating_point<T>::value) {
                ((float*)dstv)[didx] = div_rn((float)acc, (float)p.L);
            } else if constexpr (std::is_floating_point<T>::value) {
                ((double*)dstv)[didx] = div_rn((double)acc, (double)p.L);
            } else {
                const double s = IS_BOOL ? (double)iacc : (double)acc;
                ((double*)dstv)[didx] = div_rn(s, (double)p.L);
            }
        } else if (RED == SB200_MAX || RED == SB200_MIN || RED == SB200_KERNELDOT) {
            ((T*)dstv)[didx] = acc;
        } else if (RED == SB200_LIFE) {
            const unsigned m = (c != T(0)) ? p.survive : p.born;
            ((T*)dstv)[didx] = (T)((m >> cnt) & 1u);
        } else if (RED == SB200_DIFFUSION) {
            if constexpr (std::is_floating_point<T>::value) {
                const T lc = mul_rn((T)p.L, c);
                const T u = sub_rn(acc, lc);
                const T v = mul_rn((T)p.alpha, u);
                ((T*)dstv)[didx] = add_rn(c, v);
            }
        }
    }
}

template <typename T, bool IS_BOOL>
static int launch_gather_t(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const DevDesc& p = pl.dd;
    const long long total = p.n[0] * p.n[1] * p.n[2];
    if (total == 0) return SB200_OK;
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 32;
    if (blocks > cap) blocks = cap;
#define SB_GG(RED)                                                                                      \
    case RED:                                                                                           \
        gather_generic<T, RED, IS_BOOL><<<(unsigned)blocks, threads, 0, st>>>(p, (const T*)src, dst);    \
        break;
    switch (p.reducer) {
        SB_GG(SB200_SUM)
        SB_GG(SB200_MEAN)
        SB_GG(SB200_MIN)
        SB_GG(SB200_MAX)
        SB_GG(SB200_LIFE)
    case SB200_KERNELDOT:
        if constexpr (sizeof(T) >= 4) { gather_generic<T, SB200_KERNELDOT, IS_BOOL><<<(unsigned)blocks, threads, 0, st>>>(p, (const T*)src, dst); break; }
        else { set_error("kernelproduct is not supported for 1-byte element types"); return SB200_EUNSUPPORTED; }
    case SB200_DIFFUSION:
        if constexpr (std::is_floating_point<T>::value) { gather_generic<T, SB200_DIFFUSION, IS_BOOL><<<(unsigned)blocks, threads, 0, st>>>(p, (const T*)src, dst); break; }
        else { set_error("diffusion needs a floating-point element type"); return SB200_EUNSUPPORTED; }
    default:
        set_error("unsupported reducer %d", p.reducer);
        return SB200_EUNSUPPORTED;
    }
#undef SB_GG
    SB_LAUNCH_CHECK();
    set_kernel_name("gather_generic");
    return SB200_OK;
}

int launch_generic_gather(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    switch (pl.d.eltype) {
    case SB200_BOOL: return launch_gather_t<uint8_t, true>(pl, src, dst, st);
    case SB200_U8: return launch_gather_t<uint8_t, false>(pl, src, dst, st);
    case SB200_I32: return launch_gather_t<int32_t, false>(pl, src, dst, st);
    case SB200_I64: return launch_gather_t<int64_t, false>(pl, src, dst, st);
    case SB200_F32: return launch_gather_t<float, false>(pl, src, dst, st);
    case SB200_F64: return launch_gather_t<double, false>(pl, src, dst, st);
    }
    set_error("unsupported eltype %d", pl.d.eltype);
    return SB200_EUNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------ halo
// blockIdx.y = slab (axis*2 + side). Each slab is `off[axis]` thick and spans the full parent on the other
// axes, corners included, exactly like the 2N broadcasts of src/array.jl:207-226. Values come from the inner
// region only, so overlapping slabs write identical values.
template <typename T>
__global__ void __launch_bounds__(256) halo_kernel(DevDesc p, T* __restrict__ par) {
    const int axis = blockIdx.y >> 1, side = blockIdx.y & 1;
    if (axis >= p.ndim || p.soff[axis] == 0) return;
    long long ext[3] = {1, 1, 1};
    for (int a = 0; a < p.ndim; a++) ext[a] = p.sext[a];
    const long long thick = side == 0 ? p.soff[axis] : ext[axis] - p.soff[axis] - p.size[axis];
    if (thick <= 0) return;
    const long long base = side == 0 ? 0 : p.soff[axis] + p.size[axis];
    long long dims[3] = {ext[0], ext[1], ext[2]};
    dims[axis] = thick;
    const long long total = dims[0] * dims[1] * dims[2];
    const T pv = pad_of<T>(p.padbits);
    for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < total;
         id += (long long)gridDim.x * blockDim.x) {
        long long P[3], rem = id;
        P[0] = rem % dims[0]; rem /= dims[0];
        P[1] = rem % dims[1]; rem /= dims[1];
        P[2] = rem;
        P[axis] += base;
        bool use = false, remv = false;
        long long sidx = 0, didx = 0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (a < p.ndim) {
                long long i = P[a] - p.soff[a];
                if (i < 0 || i >= p.size[a]) {
                    if (p.bc[a] == SB200_USE) use = true;
                    else if (p.bc[a] == SB200_REMOVE) remv = true;
                    else i = bounded(i, p.size[a], p.bc[a]);
                }
                sidx += (i + p.soff[a]) * p.sstr[a];
                didx += P[a] * p.sstr[a];
            }
        }
        if (use) continue;
        par[didx] = remv ? pv : par[sidx];
    }
}

int launch_update_halo(const Plan& pl, void* parent, cudaStream_t st) {
    const DevDesc& p = pl.dd;
    bool any = false;
    long long maxslab = 0;
    for (int a = 0; a < p.ndim; a++) {
        if (p.soff[a] > 0 && p.bc[a] != SB200_USE) any = true;
        long long slab = p.soff[a];
        for (int b = 0; b < p.ndim; b++) if (b != a) slab *= p.sext[b];
        if (slab > maxslab) maxslab = slab;
    }
    if (!any) return SB200_OK;  // Conditional / Use: nothing to do (src/array.jl:199-200)
    long long bx = (maxslab + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, 2 * p.ndim);
    switch (sb::elsize(pl.d.eltype)) {
    case 1: halo_kernel<uint8_t><<<grid, 256, 0, st>>>(p, (uint8_t*)parent); break;
    case 4: halo_kernel<uint32_t><<<grid, 256, 0, st>>>(p, (uint32_t*)parent); break;
    case 8: halo_kernel<uint64_t><<<grid, 256, 0, st>>>(p, (uint64_t*)parent); break;
    default: set_error("unsupported eltype %d", pl.d.eltype); return SB200_EUNSUPPORTED;
    }
    SB_LAUNCH_CHECK();
    set_kernel_name("halo_kernel");
    return SB200_OK;
}

// ------------------------------------------------------------------------------------------------ scatter
// Per destination cell: fold the contributions in the reference's serial order. For a destination column
// nj the contributing source columns nj - o1_k fall into distinct passes mod1(j, 2R+1); within a pass rows
// ascend, within a cell k ascends. `order` holds, per residue (nj mod (2R+1)), the k's sorted that way
// (pass, then o0 descending == source row ascending, then k). Cells whose sources can wrap or reflect onto
// them (within R of an edge under Wrap/Reflect) enumerate every pre-image and sort by the explicit key
// (pass, source column, source row, k).
template <typename T> __device__ __forceinline__ T scatter_fold(T acc, T val, int op) {
    if (op == SB200_OP_ADD) return add_rn(acc, val);
    if (op == SB200_OP_MAX) return jl_max(acc, val);
    return jl_min(acc, val);
}

__device__ __forceinline__ int preimages(long long n, long long s, int bc, int R, long long out[3]) {
    int c = 0;
    out[c++] = n;
    if (bc == SB200_WRAP) {
        if (n - s >= -R) out[c++] = n - s;
        if (n + s <= s - 1 + R) out[c++] = n + s;
    } else if (bc == SB200_REFLECT) {
        if (n >= 1 && -n >= -R) out[c++] = -n;
        if (n <= s - 2 && 2 * (s - 1) - n <= s - 1 + R) out[c++] = 2 * (s - 1) - n;
    }
    return c;
}

template <typename T>
__global__ void __launch_bounds__(256) scatter_generic(DevDesc p, const int* __restrict__ order,
                                                       const T* __restrict__ src, T* __restrict__ dst) {
    const long long ny = p.size[0], nx = p.size[1];
    const long long total = p.n[0] * p.n[1];  // destination rectangle [lo, lo + n) (whole array by default)
    const int S = 2 * p.R + 1;
    const T* __restrict__ w = (const T*)p.weights;
    const bool zero = p.flags & SB200_FLAG_ZERO_DEST;
    for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < total;
         id += (long long)gridDim.x * blockDim.x) {
        const long long ni = p.lo[0] + id % p.n[0], nj = p.lo[1] + id / p.n[0];
        T* cell = &dst[(ni + p.doff[0]) * p.dstr[0] + (nj + p.doff[1]) * p.dstr[1]];
        T acc = zero ? T(0) : *cell;
        // cells that a wrapped (raw -j -> s-j, raw s-1+j -> j-1) or reflected (raw -j -> j, raw s-1+j -> s-1-j)
        // target can land on, j = 1..R
        const bool edge0 = (p.bc[0] == SB200_WRAP && (ni < p.R || ni >= ny - p.R)) ||
                           (p.bc[0] == SB200_REFLECT && (ni <= p.R || ni >= ny - 1 - p.R));
        const bool edge1 = (p.bc[1] == SB200_WRAP && (nj < p.R || nj >= nx - p.R)) ||
                           (p.bc[1] == SB200_REFLECT && (nj <= p.R || nj >= nx - 1 - p.R));
        if (!edge0 && !edge1) {
            const int* ord = order + (int)(nj % S) * p.L;
            for (int q = 0; q < p.L; q++) {
                const int k = __ldg(&ord[q]);
                const long long si = ni - __ldg(&p.offs[3 * k]), sj = nj - __ldg(&p.offs[3 * k + 1]);
                if (si < 0 || si >= ny || sj < 0 || sj >= nx) continue;  // that source does not exist
                T val = w[k];
                if (p.scatter_rule == SB200_SCATTER_CENTER_WEIGHTS)
                    val = mul_rn(__ldg(&src[(si + p.soff[0]) * p.sstr[0] + (sj + p.soff[1]) * p.sstr[1]]), val);
                acc = scatter_fold(acc, val, p.scatter_op);
            }
        } else {
            long long pre0[3], pre1[3];
            const int c0 = preimages(ni, ny, p.bc[0], p.R, pre0), c1 = preimages(nj, nx, p.bc[1], p.R, pre1);
            unsigned long long last = 0;  // keys are >= 1<<57 (pass >= 1)
            for (;;) {
                unsigned long long best = ~0ULL;
                long long bsi = 0, bsj = 0;
                int bk = -1;
                for (int k = 0; k < p.L; k++) {
                    const int o0 = __ldg(&p.offs[3 * k]), o1 = __ldg(&p.offs[3 * k + 1]);
                    for (int a = 0; a < c0; a++) {
                        const long long si = pre0[a] - o0;
                        if (si < 0 || si >= ny) continue;
                        for (int b = 0; b < c1; b++) {
                            const long long sj = pre1[b] - o1;
                            if (sj < 0 || sj >= nx) continue;
                            const unsigned long long key = ((unsigned long long)(sj % S + 1) << 57) |
                                                           ((unsigned long long)sj << 34) |
                                                           ((unsigned long long)si << 10) | (unsigned)k;
                            if (key > last && key < best) { best = key; bsi = si; bsj = sj; bk = k; }
                        }
                    }
                }
                if (bk < 0) break;
                last = best;
                T val = w[bk];
                if (p.scatter_rule == SB200_SCATTER_CENTER_WEIGHTS)
                    val = mul_rn(__ldg(&src[(bsi + p.soff[0]) * p.sstr[0] + (bsj + p.soff[1]) * p.sstr[1]]), val);
                acc = scatter_fold(acc, val, p.scatter_op);
            }
        }
        *cell = acc;
    }
}

int launch_generic_scatter_rect(const Plan& pl, const void* src, void* dst, cudaStream_t st, long long lo0, long long hi0,
                                long long lo1, long long hi1) {
    DevDesc p = pl.dd;
    p.lo[0] = lo0; p.n[0] = hi0 - lo0; p.lo[1] = lo1; p.n[1] = hi1 - lo1;
    if (p.n[0] <= 0 || p.n[1] <= 0) return SB200_OK;
    const long long total = p.n[0] * p.n[1];
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 32;
    if (blocks > cap) blocks = cap;
    switch (pl.d.eltype) {
    case SB200_I32: scatter_generic<int32_t><<<(unsigned)blocks, 256, 0, st>>>(p, pl.scatter_order_dev, (const int32_t*)src, (int32_t*)dst); break;
    case SB200_I64: scatter_generic<int64_t><<<(unsigned)blocks, 256, 0, st>>>(p, pl.scatter_order_dev, (const int64_t*)src, (int64_t*)dst); break;
    case SB200_F32: scatter_generic<float><<<(unsigned)blocks, 256, 0, st>>>(p, pl.scatter_order_dev, (const float*)src, (float*)dst); break;
    case SB200_F64: scatter_generic<double><<<(unsigned)blocks, 256, 0, st>>>(p, pl.scatter_order_dev, (const double*)src, (double*)dst); break;
    default:
        set_error("scatterstencil! supports Int32/Int64/Float32/Float64, got eltype %d", pl.d.eltype);
        return SB200_EUNSUPPORTED;
    }
    SB_LAUNCH_CHECK();
    set_kernel_name("scatter_generic");
    return SB200_OK;
}

int launch_generic_scatter(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    return launch_generic_scatter_rect(pl, src, dst, st, 0, pl.dd.size[0], 0, pl.dd.size[1]);
}

}  // namespace sb
