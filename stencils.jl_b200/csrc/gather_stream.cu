// gather_stream.cu — 2-D gather for ANY offset table (Positional / NamedStencil / Rectangle / Annulus / slashes /
// Cardinal / Ordinal / larger named shapes ...) as a TMA-fed row-streaming kernel with a run-time tap table.
//
// Replaces gatherstencil_kernel! (src/gatherstencil.jl:105-109) + the neighbour read path (src/array.jl:91-138)
// wherever stream2d.cuh has no compile-time instantiation. Same data movement as stream2d / scatter_stream: a CTA
// owns a strip of GS_BXB bytes of the contiguous axis and streams along axis 1; one producer thread issues
// cp.async.bulk (UBLKCP) copies of one source column segment (+ halo) per stage into a ring of shared-memory
// stages and resolves Wrap / Reflect / ring rows and the Wrap halo of axis 0 by choosing source addresses; the
// 2R+1 source columns a destination column needs stay resident in the ring, so every cell is read from HBM once.
// Lane l of a warp owns cells l, l+32, ... of the warp's 512 bytes, so the shared-memory read of a tap with any
// offset is conflict-free and every global store is a coalesced 128-byte access. The fold visits the taps in
// table order (the reference's offset order): bit-identical to the StaticArrays left fold.
// The R cells next to each end of axis 0 under Remove / Reflect go to gather_generic as two thin bands, so no warp
// of this kernel runs an edge path (a slower edge warp would pace its whole CTA).
// Two producers feed the same ring: bulk copies (TMA engine) when rows are 16-byte aligned and axis 0 is unpadded, and
// element-granular cp.async (LDGSTS, completion counted on the same mbarriers) for everything else — Halo padding
// (the ring of axis 0 is read straight through: parent index = logical + R, src/array.jl:367-377), odd widths, pitches
// that are not multiples of 16 bytes. The consumers never notice: their lane-strided scalar accesses have no alignment.
// Algorithmic traffic: sizeof(T) read + sizeof(T) written per cell.
#include <algorithm>
#include <type_traits>
#include "common.cuh"
#include "tma.cuh"

namespace sb {

constexpr int GS_WARPS = 8;                   // consumer warps
constexpr int GS_PW = 1;                      // producer warps (more did not help the cp.async mode and cost the bulk mode 4 %)
constexpr int GS_BXB = GS_WARPS * 512;        // strip width in bytes
constexpr int GS_LEFT = 128;                  // margin: global and shared addresses of the main copy agree mod 128
constexpr int GS_STAGE = GS_LEFT + GS_BXB + 128;
constexpr int GS_NS = 13;                     // ring slots: 2R+1 resident columns + prefetch (three CTAs per SM)
constexpr int GS_HDR = 512;                   // mbarriers
constexpr int GS_TAB = 256;                   // taps
constexpr int GS_SMEM = GS_HDR + GS_TAB * 16 + GS_NS * GS_STAGE;
constexpr int GS_MAXR = 4;

// One tap, 16 bytes in shared memory: byte shift inside the column segment (+o0 cells), ring distance of its source
// column from the oldest resident one (R + o1 stages), the row offset itself, the kernelproduct weight.
template <typename T> struct GsTap { short boff, d1, o1, pad_; T w; };

template <typename T> struct GsParams {
    const T* src;
    T* dst;
    long long spitch, dpitch;   // elements per column (axis-1 stride)
    int W, H;                   // logical size: W along the contiguous axis
    int soff0, soff1, doff0, doff1;
    int cpasync;                // 1: element-granular cp.async producer (unaligned / padded axis 0), 0: bulk copies
    int bc0, bc1;
    T pad, alpha;
    int x_lo, x_hi;             // destination cells handled here along axis 0 (edge bands are gather_generic's)
    int y_lo, rows;             // destination columns [y_lo, y_lo + rows)
    int nstrips, nruns;
    int R, L;
    const int* offs;            // [L][3]
    const T* weights;           // [L] (KERNELDOT) or null
};

template <typename T> __device__ __forceinline__ long long gs_map_row(const GsParams<T>& p, int r) {
    if (p.soff1 > 0) return (long long)r + p.soff1;
    if (r >= 0 && r < p.H) return r;
    if (p.bc1 == SB200_WRAP) return r < 0 ? r + p.H : r - p.H;
    if (p.bc1 == SB200_REFLECT) return r < 0 ? -r : 2 * (p.H - 1) - r;
    return -1;
}

template <typename T, int RED> __device__ __forceinline__ T gs_fold(T acc, T v, T w) {
    if (RED == SB200_MAX) return jl_max(acc, v);
    if (RED == SB200_MIN) return jl_min(acc, v);
    if (RED == SB200_KERNELDOT) return add_rn(acc, mul_rn(v, w));
    return add_rn(acc, v);
}
template <typename T, int RED> __device__ __forceinline__ T gs_first(T v, T w) {
    if (RED == SB200_KERNELDOT) return add_rn(T(0), mul_rn(v, w));  // acc = zero(T); acc += v*w
    return v;
}

template <typename T, int RED>
__global__ void __launch_bounds__((GS_WARPS + GS_PW) * 32, 3) gather_stream_kernel(const __grid_constant__ GsParams<T> p) {
    constexpr int VX = 16 / (int)sizeof(T);
    constexpr int EW = 512 / (int)sizeof(T);   // elements per warp
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + GS_NS;
    unsigned char* tabraw = smem + GS_HDR;
    unsigned char* ring = smem + GS_HDR + GS_TAB * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = p.R, L = p.L;
    if (threadIdx.x == 0) {
        // full: one arrive.expect_tx (bulk copies) or one arrive-on-completion per producer lane (cp.async)
        for (int s = 0; s < GS_NS; s++) { mbar_init(&full[s], p.cpasync ? 32 : 1); mbar_init(&empty[s], GS_WARPS); }
        mbar_fence_init();
    }
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
        GsTap<T>* t = reinterpret_cast<GsTap<T>*>(tabraw + k * 16);
        t->boff = (short)(p.offs[3 * k] * (int)sizeof(T));
        t->d1 = (short)(R + p.offs[3 * k + 1]);
        t->o1 = (short)p.offs[3 * k + 1];
        t->pad_ = 0;
        t->w = p.weights ? p.weights[k] : T(0);
    }
    __syncthreads();
    const int ntasks = p.nstrips * p.nruns;
    const int Wb = p.W * (int)sizeof(T);
    const int HLB = ((R * (int)sizeof(T) + 15) / 16) * 16;
    constexpr int EWS = GS_BXB / (int)sizeof(T);   // cells per strip
    const bool rows_can_pad = p.soff1 == 0 && p.bc1 == SB200_REMOVE;
    unsigned kb = 0;  // ring position of stage 0 of the current task
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int strip = task % p.nstrips, run = task / p.nstrips;
        const int x0b = strip * GS_BXB;
        const int wbytes = min(GS_BXB, Wb - x0b);
        const int y0 = p.y_lo + (int)((long long)p.rows * run / p.nruns);
        const int y1 = p.y_lo + (int)((long long)p.rows * (run + 1) / p.nruns);
        const int nout = y1 - y0;
        const int nst = nout + 2 * R;  // stage i holds source column y0 - R + i
        if (warp >= GS_WARPS && p.cpasync) {
            // ---------------- producer, element-granular: lane l copies cells l, l+32, ... of every segment ----------------
            const int xs = strip * EWS, wc = min(EWS, p.W - xs);            // strip cells [xs, xs + wc)
            const bool ring0 = p.soff0 > 0;                                  // axis 0 has a ring: every neighbour is in the parent
            const int lA = (xs > 0 || ring0) ? R : 0;
            const int rA = ring0 ? R : min(R, p.W - (xs + wc));
            const bool wrap0 = !ring0 && p.bc0 == SB200_WRAP;
            for (int i = warp - GS_WARPS; i < nst; i += GS_PW) {
                const unsigned k = kb + i;
                const int slot = k % GS_NS;
                mbar_wait_producer(&empty[slot], ((k / GS_NS) & 1) ^ 1);
                T* srow = reinterpret_cast<T*>(ring + slot * GS_STAGE + GS_LEFT);   // strip cell 0
                const long long prow = gs_map_row(p, y0 - R + i);
                if (prow >= 0) {
                    const T* g = p.src + prow * p.spitch + p.soff0;                  // logical cell 0 of the row
                    for (int e = lane - lA; e < wc + rA; e += 32) cp_async_elem(srow + e, g + xs + e);
                    if (wrap0 && lA == 0)
                        for (int e = lane; e < R; e += 32) cp_async_elem(srow - R + e, g + p.W - R + e);
                    if (wrap0 && rA < R)
                        for (int e = lane; e < R - rA; e += 32) cp_async_elem(srow + wc + rA + e, g + e);
                    cp_async_arrive_noinc(&full[slot]);
                } else {
                    mbar_arrive(&full[slot]);
                }
            }
            kb += nst;
            continue;
        }
        if (warp >= GS_WARPS) {
            // ---------------- producer, bulk copies ----------------
            if (warp == GS_WARPS && lane == 0) {
                const bool l_in = x0b > 0, r_in = x0b + wbytes < Wb;
                const bool l_wrap = !l_in && p.bc0 == SB200_WRAP, r_wrap = !r_in && p.bc0 == SB200_WRAP;
                const int r_in_bytes = r_in ? min(HLB, Wb - (x0b + wbytes)) : 0;
                const int mstart = x0b - (l_in ? HLB : 0), mdst = GS_LEFT - (l_in ? HLB : 0);
                const unsigned mlen = wbytes + (l_in ? HLB : 0) + r_in_bytes;
                const int wrap_bytes = min(HLB, Wb);
                const unsigned rowbytes = mlen + (l_wrap ? wrap_bytes : 0) + (r_wrap ? wrap_bytes : 0);
                for (int i = 0; i < nst; i++) {
                    const unsigned k = kb + i;
                    const int slot = k % GS_NS;
                    mbar_wait_producer(&empty[slot], ((k / GS_NS) & 1) ^ 1);
                    unsigned char* srow = ring + slot * GS_STAGE;
                    const long long prow = gs_map_row(p, y0 - R + i);
                    mbar_arrive_expect_tx(&full[slot], prow >= 0 ? rowbytes : 0u);
                    if (prow >= 0) {
                        const unsigned char* g = reinterpret_cast<const unsigned char*>(p.src + prow * p.spitch);
                        bulk_g2s(srow + mdst, g + mstart, mlen, &full[slot]);
                        if (l_wrap) bulk_g2s(srow + GS_LEFT - wrap_bytes, g + Wb - wrap_bytes, wrap_bytes, &full[slot]);
                        if (r_wrap) bulk_g2s(srow + GS_LEFT + wbytes, g, wrap_bytes, &full[slot]);
                    }
                }
            }
            kb += nst;
            continue;
        }
        // ---------------- consumers ----------------
        const int e0 = warp * EW + lane;                        // first element of this lane inside the strip
        const int gx0 = x0b / (int)sizeof(T) + e0;              // its global index along axis 0
        bool inr[VX];
#pragma unroll
        for (int v = 0; v < VX; v++) {
            const int x = gx0 + v * 32;
            inr[v] = x >= p.x_lo && x < p.x_hi;   // cells outside the range are computed from in-stage bytes, never stored
        }
        for (int i = 0; i < 2 * R; i++) {         // source columns y0-R .. y0+R-1
            const unsigned k = kb + i;
            mbar_wait(&full[k % GS_NS], (k / GS_NS) & 1);
        }
        const unsigned char* mine = ring + GS_LEFT + e0 * (int)sizeof(T);   // this lane's first cell in slot 0
        unsigned slot_t = kb % GS_NS;                                      // slot of stage t (source column y - R)
        unsigned slot_n = (kb + 2 * R) % GS_NS;                            // slot of stage t + 2R (source column y + R)
        unsigned par_n = ((kb + 2 * R) / GS_NS) & 1;
        T* drow = p.dst + (long long)(y0 + p.doff1) * p.dpitch + p.doff0 + gx0;
        for (int t = 0; t < nout; t++, drow += p.dpitch) {
            const int y = y0 + t;
            mbar_wait(&full[slot_n], par_n);
            T acc[VX];
            if (!rows_can_pad || (y - R >= 0 && y + R < p.H)) {
                // every source column of every tap is in the ring: no per-tap tests, four taps of loads in flight
                {
                    const GsTap<T> tp = *reinterpret_cast<const GsTap<T>*>(tabraw);
                    unsigned sl = slot_t + tp.d1;
                    sl = sl >= GS_NS ? sl - GS_NS : sl;
                    const T* srow = reinterpret_cast<const T*>(mine + sl * GS_STAGE + tp.boff);
#pragma unroll
                    for (int v = 0; v < VX; v++) acc[v] = gs_first<T, RED>(srow[v * 32], tp.w);
                }
#pragma unroll 4
                for (int q = 1; q < L; q++) {
                    const GsTap<T> tp = *reinterpret_cast<const GsTap<T>*>(tabraw + q * 16);
                    unsigned sl = slot_t + tp.d1;
                    sl = sl >= GS_NS ? sl - GS_NS : sl;
                    const T* srow = reinterpret_cast<const T*>(mine + sl * GS_STAGE + tp.boff);
#pragma unroll
                    for (int v = 0; v < VX; v++) acc[v] = gs_fold<T, RED>(acc[v], srow[v * 32], tp.w);
                }
            } else {
                // Remove on axis 1 near the array ends: columns outside the array read padval
                for (int q = 0; q < L; q++) {
                    const GsTap<T> tp = *reinterpret_cast<const GsTap<T>*>(tabraw + q * 16);
                    const int r = y + tp.o1;
                    unsigned sl = slot_t + tp.d1;
                    sl = sl >= GS_NS ? sl - GS_NS : sl;
                    const T* srow = reinterpret_cast<const T*>(mine + sl * GS_STAGE + tp.boff);
                    const bool oob = r < 0 || r >= p.H;
#pragma unroll
                    for (int v = 0; v < VX; v++) {
                        const T x = oob ? p.pad : srow[v * 32];
                        acc[v] = q == 0 ? gs_first<T, RED>(x, tp.w) : gs_fold<T, RED>(acc[v], x, tp.w);
                    }
                }
            }
            if constexpr (RED == SB200_MEAN) {
#pragma unroll
                for (int v = 0; v < VX; v++) acc[v] = div_rn(acc[v], (T)L);
            }
            if constexpr (RED == SB200_DIFFUSION) {
                unsigned sl = slot_t + R;
                sl = sl >= GS_NS ? sl - GS_NS : sl;
                const T* crow = reinterpret_cast<const T*>(mine + sl * GS_STAGE);
#pragma unroll
                for (int v = 0; v < VX; v++) {
                    const T c = crow[v * 32];
                    acc[v] = add_rn(c, mul_rn(p.alpha, sub_rn(acc[v], mul_rn((T)L, c))));
                }
            }
#pragma unroll
            for (int v = 0; v < VX; v++)
                if (inr[v]) drow[v * 32] = acc[v];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot_t]);  // source column y-R is done
            if (++slot_t == GS_NS) slot_t = 0;
            if (++slot_n == GS_NS) { slot_n = 0; par_n ^= 1; }
        }
        // release the 2R trailing stages
        __syncwarp();
        if (lane == 0)
            for (int i = nout; i < nst; i++) mbar_arrive(&empty[(kb + i) % GS_NS]);
        kb += nst;
    }
}

template <typename T, int RED> static int gs_launch(GsParams<T>& p, cudaStream_t st) {
    static thread_local int cfg_dev = -1, ctas_per_sm = 0;
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (dev != cfg_dev) {
        SB_CUDA(cudaFuncSetAttribute(gather_stream_kernel<T, RED>, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_stream_kernel<T, RED>, (GS_WARPS + GS_PW) * 32, GS_SMEM) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        ctas_per_sm = per_sm;
        cfg_dev = dev;
    }
    const long long ctas = (long long)ctas_per_sm * num_sms();
    long long nruns = std::max<long long>(1, 4 * ctas / p.nstrips);
    nruns = std::min<long long>(nruns, std::max(1, p.rows / (4 * (2 * p.R + 1))));
    p.nruns = (int)nruns;
    const long long grid = std::min<long long>(ctas, (long long)p.nstrips * p.nruns);
    gather_stream_kernel<T, RED><<<(unsigned)grid, (GS_WARPS + GS_PW) * 32, GS_SMEM, st>>>(p);
    SB_LAUNCH_CHECK();
    return SB200_OK;
}

template <typename T> static int gs_try(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const sb200_desc& d = pl.d;
    const DevDesc& dd = pl.dd;
    const int R = d.radius, L = d.noffsets;
    if (R < 1 || R > GS_MAXR || L < 1 || L > GS_TAB) return -1;
    for (int k = 0; k < L; k++)
        if (d.offsets_host[3 * k + 2] != 0) return -1;
    if (d.src_off[0] == 0 && d.boundary[0] == SB200_USE) return -1;
    if (d.src_off[1] == 0 && d.boundary[1] == SB200_USE) return -1;
    const long long Wb = d.size[0] * (long long)sizeof(T);
    if (Wb < 64 || d.size[0] <= 4 * R) return -1;
    // bulk copies need 16-byte aligned rows, an unpadded axis 0 and a last strip at least as wide as a wrap halo;
    // otherwise the element-granular producer
    const bool narrow_last = Wb > GS_BXB && (Wb % GS_BXB) != 0 && (Wb % GS_BXB) < 64;
    const bool aligned = d.src_off[0] == 0 && Wb % 16 == 0 && (d.src_ext[0] * sizeof(T)) % 16 == 0 && ((uintptr_t)src & 15) == 0 && !narrow_last;
    if (((uintptr_t)src | (uintptr_t)dst) % sizeof(T)) return -1;
    if (d.size[0] > (1LL << 28) || d.size[1] > (1LL << 30) || R >= d.size[1]) return -1;
    if (dd.lo[0] != 0 || dd.n[0] != d.size[0]) return -1;                       // regions only along axis 1
    if (dd.n[1] == 0) return SB200_OK;
    GsParams<T> p;
    p.src = (const T*)src; p.dst = (T*)dst;
    p.spitch = d.src_ext[0]; p.dpitch = d.dst_ext[0];
    p.W = (int)d.size[0]; p.H = (int)d.size[1];
    p.soff0 = d.src_off[0]; p.soff1 = d.src_off[1]; p.doff0 = d.dst_off[0]; p.doff1 = d.dst_off[1];
    p.cpasync = aligned ? 0 : 1;
    p.bc0 = d.boundary[0]; p.bc1 = d.boundary[1];
    memcpy(&p.pad, &d.padval_bits, sizeof(T));
    p.alpha = (T)d.alpha;
    const int band = (d.boundary[0] == SB200_WRAP || d.src_off[0] > 0) ? 0 : R;
    p.x_lo = band; p.x_hi = p.W - band;
    p.y_lo = (int)dd.lo[1]; p.rows = (int)dd.n[1];
    p.nstrips = (int)((Wb + GS_BXB - 1) / GS_BXB);
    p.R = R; p.L = L;
    p.offs = dd.offs; p.weights = d.reducer == SB200_KERNELDOT ? (const T*)dd.weights : nullptr;
    int rc = -1;
    switch (d.reducer) {
    case SB200_SUM: rc = gs_launch<T, SB200_SUM>(p, st); break;
    case SB200_MIN: rc = gs_launch<T, SB200_MIN>(p, st); break;
    case SB200_MAX: rc = gs_launch<T, SB200_MAX>(p, st); break;
    case SB200_KERNELDOT: rc = gs_launch<T, SB200_KERNELDOT>(p, st); break;
    case SB200_MEAN:
        if constexpr (std::is_floating_point<T>::value) rc = gs_launch<T, SB200_MEAN>(p, st);
        break;
    case SB200_DIFFUSION:
        if constexpr (std::is_floating_point<T>::value) rc = gs_launch<T, SB200_DIFFUSION>(p, st);
        break;
    default: break;
    }
    if (rc != SB200_OK) return rc;
    if (band > 0) {  // the two edge bands of axis 0, every boundary rule resolved per neighbour
        Plan edge = pl;
        edge.dd.lo[0] = 0; edge.dd.n[0] = band;
        if ((rc = launch_generic_gather(edge, src, dst, st))) return rc;
        edge.dd.lo[0] = p.W - band;
        if ((rc = launch_generic_gather(edge, src, dst, st))) return rc;
    }
    set_kernel_name("gather_stream_kernel");
    return SB200_OK;
}

int try_gather_stream(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const sb200_desc& d = pl.d;
    if ((d.flags & SB200_FLAG_NO_TMA) || d.ndim != 2 || d.out_eltype != d.eltype) return -1;
    switch (d.eltype) {
    case SB200_F32: return gs_try<float>(pl, src, dst, st);
    case SB200_F64: return gs_try<double>(pl, src, dst, st);
    case SB200_I32: return gs_try<int32_t>(pl, src, dst, st);
    case SB200_I64: return gs_try<int64_t>(pl, src, dst, st);
    default: return -1;
    }
}

}  // namespace sb
