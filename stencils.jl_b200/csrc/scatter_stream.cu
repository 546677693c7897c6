// scatter_stream.cu — scatterstencil! (src/scatterstencil.jl:36-112) as a TMA-fed row-streaming kernel.
//
// Same decomposition as scatter_fast.cu (read per destination cell, the reference's 2R+1 passes are the fold
// order (pass of the source column, source row, k), which for a whole destination column nj depends only on
// nj mod (2R+1)), but the data moves the way stream2d does: a CTA owns a strip of SS_BXB bytes of the contiguous
// axis and streams along axis 1 (Julia columns); a producer thread issues cp.async.bulk (UBLKCP) copies of one
// source column segment (+ halo) and one destination column segment per stage into a ring of shared-memory stages;
// the 2R+1 source columns a destination column folds stay resident in the ring, so every source cell is read from
// HBM once. Lane l of a warp owns cells l, l+32, ... of the warp's 512 bytes: every shared-memory read of a tap
// (any offset) and every global store is a conflict-free, fully coalesced 128-byte access.
// No atomics; the result is the reference's serial result bit for bit.
// Algorithmic traffic: source read + dest read + dest write (dest read dropped with SB200_FLAG_ZERO_DEST).
#include <algorithm>
#include "common.cuh"
#include "tma.cuh"

namespace sb {

constexpr int SS_WARPS = 8;
constexpr int SS_BXB = SS_WARPS * 512;        // strip width in bytes
constexpr int SS_LEFT = 128;                  // margin: global and shared addresses of the source copy agree mod 128
constexpr int SS_SROWB = SS_LEFT + SS_BXB + 128;
constexpr int SS_STAGE = SS_SROWB + SS_BXB;   // source segment | destination segment
constexpr int SS_NS = 12;                     // ring slots
constexpr int SS_HDR = 256;                   // mbarriers
constexpr int SS_TAB = 448;                  // (2R+1)*L entries of the ordered tap table (7 x 64)
constexpr int SS_SMEM = SS_HDR + SS_TAB * 16 + SS_NS * SS_STAGE;

// One ordered tap, 16 bytes in shared memory: byte shift of the source cell inside its column segment (-o0 cells),
// ring distance of its source column from the oldest resident one (R - o1 stages), the offset itself, the weight.
template <typename T> struct SsTap { short boff, d1, o0, o1; T w; };

template <typename T> struct SsParams {
    const T* src;
    T* dst;
    long long spitch, dpitch;   // elements per column (axis-1 stride)
    int W, H;                   // logical size: W along the contiguous axis
    int soff0, soff1, doff0, doff1;
    int x_lo, x_hi;             // destination cells handled here along axis 0 (edge bands are someone else's)
    int y_lo, rows;             // destination columns [y_lo, y_lo + rows)
    int nstrips, nruns;
    int R, L;
    int zero_dest;
    const int* order;           // [2R+1][L]
    const int* offs;            // [L][3]
    const T* weights;           // [L]
};

template <typename T, int OP> __device__ __forceinline__ T ss_fold(T acc, T val) {
    if (OP == SB200_OP_ADD) return add_rn(acc, val);
    if (OP == SB200_OP_MAX) return jl_max(acc, val);
    return jl_min(acc, val);
}

template <typename T, int OP, bool MULC>
__global__ void __launch_bounds__((SS_WARPS + 1) * 32, 2) scatter_stream_kernel(const __grid_constant__ SsParams<T> p) {
    constexpr int VX = 16 / (int)sizeof(T);
    constexpr int EW = 512 / (int)sizeof(T);   // elements per warp
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + SS_NS;
    unsigned char* tabraw = smem + SS_HDR;
    unsigned char* ring = smem + SS_HDR + SS_TAB * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = p.R, L = p.L, S = 2 * R + 1;
    if (threadIdx.x == 0) {
        for (int s = 0; s < SS_NS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], SS_WARPS); }
        mbar_fence_init();
    }
    // ordered tap table: entry (c, q) = q-th tap folded into a destination column with residue c
    for (int e = threadIdx.x; e < S * L; e += blockDim.x) {
        const int k = p.order[e];
        SsTap<T>* t = reinterpret_cast<SsTap<T>*>(tabraw + e * 16);
        t->o0 = (short)p.offs[3 * k];
        t->o1 = (short)p.offs[3 * k + 1];
        t->boff = (short)(-p.offs[3 * k] * (int)sizeof(T));
        t->d1 = (short)(R - p.offs[3 * k + 1]);
        t->w = p.weights[k];
    }
    __syncthreads();
    const int ntasks = p.nstrips * p.nruns;
    const int Wb = p.W * (int)sizeof(T);
    const int HLB = ((R * (int)sizeof(T) + 15) / 16) * 16;
    unsigned kb = 0;  // ring position of stage 0 of the current task
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int strip = task % p.nstrips, run = task / p.nstrips;
        const int x0b = strip * SS_BXB;
        const int wbytes = min(SS_BXB, Wb - x0b);
        const int y0 = p.y_lo + (int)((long long)p.rows * run / p.nruns);
        const int y1 = p.y_lo + (int)((long long)p.rows * (run + 1) / p.nruns);
        const int nout = y1 - y0;
        const int nst = nout + 2 * R;  // stage i: source column y0-R+i and (i >= 2R) destination column y0+i-2R
        if (warp == SS_WARPS) {
            // ---------------- producer ----------------
            if (lane == 0) {
                const bool l_in = x0b > 0, r_in = x0b + wbytes < Wb;
                const int r_in_bytes = r_in ? min(HLB, Wb - (x0b + wbytes)) : 0;
                const int mstart = x0b - (l_in ? HLB : 0), mdst = SS_LEFT - (l_in ? HLB : 0);
                const unsigned mlen = wbytes + (l_in ? HLB : 0) + r_in_bytes;
                for (int i = 0; i < nst; i++) {
                    const unsigned k = kb + i;
                    const int slot = k % SS_NS;
                    mbar_wait_producer(&empty[slot], ((k / SS_NS) & 1) ^ 1);
                    unsigned char* sbase = ring + slot * SS_STAGE;
                    const int sj = y0 - R + i;
                    const bool has_src = sj >= 0 && sj < p.H;
                    const bool has_dst = i >= 2 * R && !p.zero_dest;
                    mbar_arrive_expect_tx(&full[slot], (has_src ? mlen : 0u) + (has_dst ? (unsigned)wbytes : 0u));
                    if (has_src) {
                        const unsigned char* g = reinterpret_cast<const unsigned char*>(p.src + (long long)(sj + p.soff1) * p.spitch + p.soff0);
                        bulk_g2s(sbase + mdst, g + mstart, mlen, &full[slot]);
                    }
                    if (has_dst) {
                        const unsigned char* g = reinterpret_cast<const unsigned char*>(p.dst + (long long)(y0 + i - 2 * R + p.doff1) * p.dpitch + p.doff0);
                        bulk_g2s(sbase + SS_SROWB, g + x0b, wbytes, &full[slot]);
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
        bool plain = true;
#pragma unroll
        for (int v = 0; v < VX; v++) {
            const int x = gx0 + v * 32;
            inr[v] = x >= p.x_lo && x < p.x_hi;
            plain = plain && (!inr[v] || (x - R >= 0 && x + R < p.W));  // cells outside the range are computed but never stored
        }
        const bool warp_plain = __all_sync(0xffffffffu, plain);
        // wait for the first 2R stages (source columns y0-R .. y0+R-1)
        for (int i = 0; i < 2 * R; i++) {
            const unsigned k = kb + i;
            mbar_wait(&full[k % SS_NS], (k / SS_NS) & 1);
        }
        const unsigned char* mine = ring + SS_LEFT + e0 * (int)sizeof(T);   // this lane's first cell in slot 0
        unsigned slot_t = kb % SS_NS;                                      // slot of stage t (source column y - R)
        unsigned slot_d = (kb + 2 * R) % SS_NS;                            // slot of stage t + 2R (destination column y)
        unsigned par_d = ((kb + 2 * R) / SS_NS) & 1;
        for (int t = 0; t < nout; t++) {
            const int y = y0 + t;
            mbar_wait(&full[slot_d], par_d);
            T acc[VX];
#pragma unroll
            for (int v = 0; v < VX; v++)
                acc[v] = p.zero_dest ? T(0) : *reinterpret_cast<const T*>(ring + slot_d * SS_STAGE + SS_SROWB + (e0 + v * 32) * (int)sizeof(T));
            const unsigned char* tab = tabraw + (y % S) * L * 16;
            if (warp_plain && y - R >= 0 && y + R < p.H) {
                // every source of every tap exists: no per-tap tests, loads of four taps in flight
#pragma unroll 4
                for (int q = 0; q < L; q++) {
                    const SsTap<T> tp = *reinterpret_cast<const SsTap<T>*>(tab + q * 16);
                    unsigned sl = slot_t + tp.d1;
                    sl = sl >= SS_NS ? sl - SS_NS : sl;
                    const T* srow = reinterpret_cast<const T*>(mine + sl * SS_STAGE + tp.boff);
#pragma unroll
                    for (int v = 0; v < VX; v++) {
                        const T val = MULC ? mul_rn(srow[v * 32], tp.w) : tp.w;
                        acc[v] = ss_fold<T, OP>(acc[v], val);
                    }
                }
            } else {
                for (int q = 0; q < L; q++) {
                    const SsTap<T> tp = *reinterpret_cast<const SsTap<T>*>(tab + q * 16);
                    const int sj = y - tp.o1;
                    if (sj < 0 || sj >= p.H) continue;  // that source column does not exist
                    unsigned sl = slot_t + tp.d1;
                    sl = sl >= SS_NS ? sl - SS_NS : sl;
                    const T* srow = reinterpret_cast<const T*>(mine + sl * SS_STAGE + tp.boff);
#pragma unroll
                    for (int v = 0; v < VX; v++) {
                        const int si = gx0 + v * 32 - tp.o0;
                        if (inr[v] && si >= 0 && si < p.W) {
                            const T val = MULC ? mul_rn(srow[v * 32], tp.w) : tp.w;
                            acc[v] = ss_fold<T, OP>(acc[v], val);
                        }
                    }
                }
            }
            T* drow = p.dst + (long long)(y + p.doff1) * p.dpitch + p.doff0 + gx0;
#pragma unroll
            for (int v = 0; v < VX; v++)
                if (inr[v]) drow[v * 32] = acc[v];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot_t]);  // source column y-R is done
            if (++slot_t == SS_NS) slot_t = 0;
            if (++slot_d == SS_NS) { slot_d = 0; par_d ^= 1; }
        }
        // release the 2R trailing stages
        __syncwarp();
        if (lane == 0)
            for (int i = nout; i < nst; i++) mbar_arrive(&empty[(kb + i) % SS_NS]);
        kb += nst;
    }
}

template <typename T, int OP, bool MULC> static int ss_launch(SsParams<T>& p, cudaStream_t st) {
    static thread_local int cfg_dev = -1, ctas_per_sm = 0;
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (dev != cfg_dev) {
        SB_CUDA(cudaFuncSetAttribute(scatter_stream_kernel<T, OP, MULC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SS_SMEM));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scatter_stream_kernel<T, OP, MULC>, (SS_WARPS + 1) * 32, SS_SMEM) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        ctas_per_sm = per_sm;
        cfg_dev = dev;
    }
    const long long ctas = (long long)ctas_per_sm * num_sms();
    long long nruns = std::max<long long>(1, ctas / p.nstrips);
    nruns = std::min<long long>(nruns, std::max(1, p.rows / (4 * (2 * p.R + 1))));
    p.nruns = (int)nruns;
    const long long grid = std::min<long long>(ctas, (long long)p.nstrips * p.nruns);
    scatter_stream_kernel<T, OP, MULC><<<(unsigned)grid, (SS_WARPS + 1) * 32, SS_SMEM, st>>>(p);
    SB_LAUNCH_CHECK();
    return SB200_OK;
}

template <typename T> static int ss_dispatch(SsParams<T>& p, int op, bool mulc, cudaStream_t st) {
    switch (op) {
    case SB200_OP_ADD: return mulc ? ss_launch<T, SB200_OP_ADD, true>(p, st) : ss_launch<T, SB200_OP_ADD, false>(p, st);
    case SB200_OP_MAX: return mulc ? ss_launch<T, SB200_OP_MAX, true>(p, st) : ss_launch<T, SB200_OP_MAX, false>(p, st);
    case SB200_OP_MIN: return mulc ? ss_launch<T, SB200_OP_MIN, true>(p, st) : ss_launch<T, SB200_OP_MIN, false>(p, st);
    default: return -1;
    }
}

// Interior destination cells [x_lo,x_hi) x [y_lo,y_hi) of a static-order scatter; -1 when the layout is not streamable.
template <typename T>
static int ss_try(const Plan& pl, const void* src, void* dst, cudaStream_t st, int x_lo, int x_hi, int y_lo, int y_hi) {
    const DevDesc& d = pl.dd;
    const int R = d.R, L = d.L;
    if ((2 * R + 1) * L > SS_TAB || 2 * R + 1 > SS_NS - 4) return -1;
    const long long Wb = d.size[0] * (long long)sizeof(T);
    if (Wb % 16 || Wb < 16) return -1;
    if ((d.sstr[1] * sizeof(T)) % 16 || (d.dstr[1] * sizeof(T)) % 16 || (d.soff[0] * sizeof(T)) % 16 || (d.doff[0] * sizeof(T)) % 16) return -1;
    if (((uintptr_t)src | (uintptr_t)dst) & 15) return -1;
    if (d.size[0] > (1LL << 28) || d.size[1] > (1LL << 30) || R >= d.size[0]) return -1;
    SsParams<T> p;
    p.src = (const T*)src; p.dst = (T*)dst;
    p.spitch = d.sstr[1]; p.dpitch = d.dstr[1];
    p.W = (int)d.size[0]; p.H = (int)d.size[1];
    p.soff0 = d.soff[0]; p.soff1 = d.soff[1]; p.doff0 = d.doff[0]; p.doff1 = d.doff[1];
    p.x_lo = x_lo; p.x_hi = x_hi; p.y_lo = y_lo; p.rows = y_hi - y_lo;
    p.nstrips = (int)((Wb + SS_BXB - 1) / SS_BXB);
    p.R = R; p.L = L;
    p.zero_dest = (d.flags & SB200_FLAG_ZERO_DEST) ? 1 : 0;
    p.order = pl.scatter_order_dev; p.offs = d.offs; p.weights = (const T*)d.weights;
    return ss_dispatch<T>(p, d.scatter_op, d.scatter_rule == SB200_SCATTER_CENTER_WEIGHTS, st);
}

int try_scatter_stream(const Plan& pl, const void* src, void* dst, cudaStream_t st, int x_lo, int x_hi, int y_lo, int y_hi) {
    if (pl.d.flags & SB200_FLAG_NO_TMA) return -1;
    switch (pl.d.eltype) {
    case SB200_F32: return ss_try<float>(pl, src, dst, st, x_lo, x_hi, y_lo, y_hi);
    case SB200_F64: return ss_try<double>(pl, src, dst, st, x_lo, x_hi, y_lo, y_hi);
    case SB200_I32: return ss_try<int32_t>(pl, src, dst, st, x_lo, x_hi, y_lo, y_hi);
    case SB200_I64: return ss_try<int64_t>(pl, src, dst, st, x_lo, x_hi, y_lo, y_hi);
    default: return -1;
    }
}

}  // namespace sb
