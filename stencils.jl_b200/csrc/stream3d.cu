// stream3d.cu — 3-D VonNeumann(1) gathers (diffusion / sum / mean / min / max) for Float32/Float64, 2.5-D streaming.
//
// Replaces gatherstencil_kernel! (src/gatherstencil.jl:105-109) for VonNeumann{1,3} (src/stencils/vonneumman.jl:5-15,
// offsets (0,0,-1),(0,-1,0),(-1,0,0),(1,0,0),(0,1,0),(0,0,1)). A CTA owns an (x,y) tile of S3_TXB bytes x S3_TY
// rows and marches along z. The 32 lanes of a producer warp issue cp.async.bulk (UBLKCP) copies of one z-plane
// of the tile (+1 halo row above/below, +16 B halo left/right) per stage into a ring of shared-memory stages;
// every boundary (Wrap / Reflect on y and z, ghost planes, the Wrap column halo) is resolved by the producer
// choosing source addresses. A consumer thread owns 16 bytes of x by S3_RT rows and keeps, per cell, the centre value
// of the previous plane and the partial fold (((zm + ym) + xm) + xp) + yp of the previous plane in registers; when
// plane z+1 arrives it adds zp and stores plane z. The fold order is the reference's offset order, every
// operation rounded separately (bit-identical to the Julia left fold).
// Algorithmic traffic: sizeof(T) read + sizeof(T) written per cell.
#include <algorithm>
#include "common.cuh"
#include "tma.cuh"

namespace sb {

constexpr int S3_WX = 2, S3_WY = 4;             // consumer warps across x and y
constexpr int S3_WARPS = S3_WX * S3_WY;
#ifndef SB200_S3_PRODUCERS
#define SB200_S3_PRODUCERS 2
#endif
// producer warps (the rows of a stage dealt round-robin). One producer warp per CTA was the bottleneck: every bulk copy costs
// ~14 issue slots of lane-by-lane serialisation (ELECT / R2UR / UBLKCP / BRA.U.ANY), 16-18 copies per plane. Measured on
// 1024^3 Float32 diffusion (r01k): 1 producer 637, 2 producers 763 Gcell-updates/s (0.79 -> 0.94 of the HBM roofline).
constexpr int S3_PRODUCERS = SB200_S3_PRODUCERS;
constexpr int S3_THREADS = (S3_WARPS + S3_PRODUCERS) * 32;
constexpr int S3_TXB = S3_WX * 512;             // tile width in bytes
constexpr int S3_RT = 4;                        // rows per thread
constexpr int S3_TY = S3_WY * S3_RT;            // tile height in rows
constexpr int S3_LEFT = 128;                    // margin (halo at its end): global and shared addresses agree mod 128
constexpr int S3_ROWB = S3_LEFT + S3_TXB + 128;  // shared-memory row: margin | tile | margin
constexpr int S3_STAGE = (S3_TY + 2) * S3_ROWB;
constexpr int S3_STAGES = 4;
constexpr int S3_SMEM = 128 + S3_STAGES * S3_STAGE;

template <typename T> struct S3Params {
    const T* src;
    T* dst;
    long long sp1, sp2, dp1, dp2;  // source / dest pitches (elements) of axes 1 and 2
    int X, Y, Z;                   // logical size
    int so1, so2, do0, do1, do2;   // ring / ghost offsets (source axis 0 is unpadded)
    int bc0, bc1, bc2;
    T pad, alpha;
    int z_lo, zn;                  // output planes [z_lo, z_lo + zn)
    int ntx, nty, nzruns;
    T* mirror;                     // fused ghost push: planes [m_lo, m_hi) are also stored here (plane m_lo first), or null
    int m_lo, m_hi;
    int ty;                        // rows per tile (<= S3_TY): chosen so that the tiles fill whole waves of CTAs
};

__device__ __forceinline__ long long s3_map(int r, int n, int off, int bc) {
    if (off > 0) return (long long)r + off;
    if (r >= 0 && r < n) return r;
    if (bc == SB200_WRAP) return r < 0 ? r + n : r - n;
    if (bc == SB200_REFLECT) return r < 0 ? -r : 2 * (n - 1) - r;
    return -1;
}

template <typename T, int RED> __device__ __forceinline__ T s3_op(T a, T b) {
    if (RED == SB200_MAX) return jl_max(a, b);
    if (RED == SB200_MIN) return jl_min(a, b);
    return add_rn(a, b);
}

template <typename T> struct S3Vec;
template <> struct S3Vec<float> { using type = float4; };
template <> struct S3Vec<double> { using type = double2; };

// MIRROR: the fused ghost-plane push (S3Params::mirror) is compiled in only for the boundary sweeps of slab runs.
template <typename T, int RED, bool MIRROR>
__global__ void __launch_bounds__(S3_THREADS) stream3d_kernel(const __grid_constant__ S3Params<T> p) {
    constexpr int VX = 16 / (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + S3_STAGES;
    unsigned char* ring = smem + 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S3_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], S3_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    const int ntiles = p.ntx * p.nty;
    const int ntasks = ntiles * p.nzruns;
    const int Xb = p.X * (int)sizeof(T);
    unsigned k = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int tile = task % ntiles, zrun = task / ntiles;
        const int x0b = (tile % p.ntx) * S3_TXB, y0 = (tile / p.ntx) * p.ty;
        const int wbytes = min(S3_TXB, Xb - x0b);
        const int z0 = p.z_lo + (int)((long long)p.zn * zrun / p.nzruns);
        const int z1 = p.z_lo + (int)((long long)p.zn * (zrun + 1) / p.nzruns);
        const int nsrc = z1 - z0 + 2;  // source planes z0-1 .. z1
        if (warp >= S3_WARPS) {
            // ---------------- producer warps: lane j of producer w copies row P*j+w of the plane ----------------
            const int pw = warp - S3_WARPS;
            // One bulk copy per row covers the tile plus the halo cells that are ordinary neighbours in the row; only
            // the wrapped halo of an array-edge tile needs its own (16-byte) copy.
            const bool l_in = x0b > 0, r_in = x0b + wbytes < Xb;
            const bool l_wrap = !l_in && p.bc0 == SB200_WRAP, r_wrap = !r_in && p.bc0 == SB200_WRAP;
            const int mstart = x0b - (l_in ? 16 : 0);
            const unsigned mlen = wbytes + (l_in ? 16 : 0) + (r_in ? 16 : 0);
            const int mdst = S3_LEFT - (l_in ? 16 : 0);
            const unsigned rowbytes = mlen + (l_wrap ? 16 : 0) + (r_wrap ? 16 : 0);
            // Lane j owns shared-memory row j = logical row y0-1+j (same mapping for every plane). Rows below the
            // halo row of a ragged last tile are never read.
            long long yrow = -1, yany = -1;   // yany: row `lane` of the stage (every producer counts all rows for expect_tx)
            if (lane < p.ty + 2) {
                const int y = y0 - 1 + lane;
                if (y <= p.Y) yany = s3_map(y, p.Y, p.so1, p.bc1);
            }
            const unsigned nrows = __popc(__ballot_sync(0xffffffffu, yany >= 0));
            const int srow_i = S3_PRODUCERS * lane + pw;
            if (srow_i < p.ty + 2) {
                const int y = y0 - 1 + srow_i;
                if (y <= p.Y) yrow = s3_map(y, p.Y, p.so1, p.bc1);
            }
            for (int i = 0; i < nsrc; i++, k++) {
                const int slot = k % S3_STAGES;
                const long long zpl = s3_map(z0 - 1 + i, p.Z, p.so2, p.bc2);
                if (lane == 0) {
                    mbar_wait_producer(&empty[slot], ((k / S3_STAGES) & 1) ^ 1);
                    if (pw == 0) mbar_arrive_expect_tx(&full[slot], zpl >= 0 ? nrows * rowbytes : 0u);
                }
                __syncwarp();
                if (zpl >= 0 && yrow >= 0) {
                    const unsigned char* g = reinterpret_cast<const unsigned char*>(p.src + zpl * p.sp2 + yrow * p.sp1);
                    unsigned char* srow = ring + slot * S3_STAGE + srow_i * S3_ROWB;
                    bulk_g2s(srow + mdst, g + mstart, mlen, &full[slot]);
                    if (l_wrap) bulk_g2s(srow + S3_LEFT - 16, g + Xb - 16, 16, &full[slot]);
                    if (r_wrap) bulk_g2s(srow + S3_LEFT + wbytes, g, 16, &full[slot]);
                }
            }
            continue;
        }
        // ---------------- consumers ----------------
        const int wx = warp % S3_WX, wy = warp / S3_WX;
        const int xtb = (wx * 32 + lane) * 16;         // byte offset inside the tile
        const int ry0 = wy * S3_RT;                    // first tile row of this thread
        const bool xact = xtb < wbytes;
        const int gx = (x0b + xtb) / (int)sizeof(T);
        const bool edge_l = xact && p.bc0 != SB200_WRAP && gx == 0;
        const bool edge_r = xact && p.bc0 != SB200_WRAP && gx + VX == p.X;
        const bool pad1 = p.so1 == 0 && p.bc1 == SB200_REMOVE;   // OOB rows read padval
        const bool pad2 = p.so2 == 0 && p.bc2 == SB200_REMOVE;   // OOB planes read padval
        T cprev[S3_RT][VX], part[S3_RT][VX];
#pragma unroll
        for (int r = 0; r < S3_RT; r++)
#pragma unroll
            for (int v = 0; v < VX; v++) { cprev[r][v] = T(0); part[r][v] = T(0); }
        T* __restrict__ dbase = p.dst + (long long)(y0 + ry0 + p.do1) * p.dp1 + p.do0 + gx;
        for (int i = 0; i < nsrc; i++, k++) {
            const int slot = k % S3_STAGES;
            const int z = z0 - 1 + i;                  // logical plane held by this stage
            const bool zpad = pad2 && (z < 0 || z >= p.Z);
            mbar_wait(&full[slot], (k / S3_STAGES) & 1);
            const unsigned char* sb_ = ring + slot * S3_STAGE + S3_LEFT + xtb;
            // rows ry0-1 .. ry0+RT of the tile  (shared-memory row index = tile row + 1)
            T rowv[S3_RT + 2][VX];
            T xl[S3_RT], xr[S3_RT];
#pragma unroll
            for (int r = 0; r < S3_RT + 2; r++) {
                const int y = y0 + ry0 - 1 + r;
                const bool ypad = zpad || (pad1 && (y < 0 || y >= p.Y));
                if (ypad) {
#pragma unroll
                    for (int v = 0; v < VX; v++) rowv[r][v] = p.pad;
                } else {
                    const typename S3Vec<T>::type q = *reinterpret_cast<const typename S3Vec<T>::type*>(sb_ + (ry0 + r) * S3_ROWB);
                    if constexpr (VX == 4) { rowv[r][0] = q.x; rowv[r][1] = q.y; rowv[r][2] = q.z; rowv[r][3] = q.w; }
                    else { rowv[r][0] = q.x; rowv[r][1] = q.y; }
                }
                if (r >= 1 && r <= S3_RT) {
                    if (ypad) { xl[r - 1] = p.pad; xr[r - 1] = p.pad; }
                    else {
                        // x neighbours across the 16-byte vectors come from the adjacent lanes (a scalar shared-memory
                        // read with a 16-byte lane stride would be a 4-way bank conflict); only the warp's two end
                        // lanes read the halo cells. ypad is warp-uniform, so every lane takes part in the shuffles.
                        const unsigned char* t = sb_ + (ry0 + r) * S3_ROWB;
                        T l_ = __shfl_up_sync(0xffffffffu, rowv[r][VX - 1], 1);
                        T r_ = __shfl_down_sync(0xffffffffu, rowv[r][0], 1);
                        if (lane == 0) l_ = *reinterpret_cast<const T*>(t - sizeof(T));
                        if (lane == 31) r_ = *reinterpret_cast<const T*>(t + 16);
                        xl[r - 1] = l_;
                        xr[r - 1] = r_;
                        if (edge_l) xl[r - 1] = p.bc0 == SB200_REFLECT ? rowv[r][1] : p.pad;
                        if (edge_r) xr[r - 1] = p.bc0 == SB200_REFLECT ? rowv[r][VX - 2] : p.pad;
                    }
                }
            }
            const int zo = z - 1;  // output plane completed by this stage
            const bool store = i >= 2;
#pragma unroll
            for (int r = 0; r < S3_RT; r++) {
                T out[VX];
#pragma unroll
                for (int v = 0; v < VX; v++) {
                    const T c = rowv[r + 1][v];
                    // finish plane z-1:  s = part + zp ;  zp = centre of this plane
                    const T s = s3_op<T, RED>(part[r][v], c);
                    const T cc = cprev[r][v];
                    if (RED == SB200_DIFFUSION) out[v] = add_rn(cc, mul_rn(p.alpha, sub_rn(s, mul_rn((T)6, cc))));
                    else if (RED == SB200_MEAN) out[v] = div_rn(s, (T)6);
                    else out[v] = s;
                    // start plane z:  (((zm + ym) + xm) + xp) + yp   with zm = centre of plane z-1
                    const T xm = v == 0 ? xl[r] : rowv[r + 1][v == 0 ? 0 : v - 1];
                    const T xp = v == VX - 1 ? xr[r] : rowv[r + 1][v == VX - 1 ? v : v + 1];
                    T a = s3_op<T, RED>(cc, rowv[r][v]);
                    a = s3_op<T, RED>(a, xm);
                    a = s3_op<T, RED>(a, xp);
                    a = s3_op<T, RED>(a, rowv[r + 2][v]);
                    part[r][v] = a;
                    cprev[r][v] = c;
                }
                const int y = y0 + ry0 + r;
                if (store && xact && y < p.Y && ry0 + r < p.ty) {
                    T* d = dbase + (long long)(zo + p.do2) * p.dp2 + (long long)r * p.dp1;
                    if constexpr (VX == 4) *reinterpret_cast<float4*>(d) = make_float4(out[0], out[1], out[2], out[3]);
                    else *reinterpret_cast<double2*>(d) = make_double2(out[0], out[1]);
                    if (MIRROR && zo >= p.m_lo && zo < p.m_hi) {  // boundary planes cross NVLink as they are produced
                        T* m = p.mirror + (long long)(zo - p.m_lo) * p.dp2 + (long long)(y0 + ry0 + r) * p.dp1 + gx;
                        if constexpr (VX == 4) *reinterpret_cast<float4*>(m) = make_float4(out[0], out[1], out[2], out[3]);
                        else *reinterpret_cast<double2*>(m) = make_double2(out[0], out[1]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
    }
}

template <typename T, int RED, bool MIRROR> static int s3_launch_m(S3Params<T>& p, cudaStream_t st) {
    static thread_local int cfg_dev = -1, ctas_per_sm = 0;
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (dev != cfg_dev) {
        SB_CUDA(cudaFuncSetAttribute(stream3d_kernel<T, RED, MIRROR>, cudaFuncAttributeMaxDynamicSharedMemorySize, S3_SMEM));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stream3d_kernel<T, RED, MIRROR>, S3_THREADS, S3_SMEM) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        ctas_per_sm = per_sm;
        cfg_dev = dev;
    }
    const long long ctas = (long long)ctas_per_sm * num_sms();
    // Tile height and z-runs: a task loads (ty + 2) rows x (zn / nz + 2) planes; pick the pair that minimises
    // waves x rows loaded per task (the idle tail of a partial last wave against re-read halo rows / planes).
    int best_ty = S3_TY, best = 1;
    double best_cost = 1e300;
    for (int ty = S3_TY; ty >= S3_TY / 2; ty--) {
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
    const long long ntiles = (long long)p.ntx * p.nty;
    const long long grid = std::min<long long>(ctas, ntiles * p.nzruns);
    stream3d_kernel<T, RED, MIRROR><<<(unsigned)grid, S3_THREADS, S3_SMEM, st>>>(p);
    SB_LAUNCH_CHECK();
    return SB200_OK;
}
template <typename T, int RED> static int s3_launch(S3Params<T>& p, cudaStream_t st) {
    return p.mirror ? s3_launch_m<T, RED, true>(p, st) : s3_launch_m<T, RED, false>(p, st);
}

template <typename T> static int s3_try(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const sb200_desc& d = pl.d;
    if (d.src_off[0] != 0) return -1;
    if ((d.size[0] * sizeof(T)) % 16 || (d.src_ext[0] * sizeof(T)) % 16 || (d.dst_ext[0] * sizeof(T)) % 16 ||
        (d.dst_off[0] * sizeof(T)) % 16)
        return -1;
    if (((uintptr_t)src | (uintptr_t)dst) & 15) return -1;
    if (d.size[0] * (long long)sizeof(T) < 32 || d.size[0] > (1 << 28) || d.size[1] > (1 << 28) || d.size[2] > (1 << 28)) return -1;
    if (pl.dd.lo[0] != 0 || pl.dd.n[0] != d.size[0] || pl.dd.lo[1] != 0 || pl.dd.n[1] != d.size[1]) return -1;  // z regions only
    for (int a = 0; a < 3; a++)
        if (d.src_off[a] == 0 && d.boundary[a] == SB200_USE) return -1;
    if (pl.dd.n[2] == 0) return SB200_OK;
    S3Params<T> p;
    p.src = (const T*)src; p.dst = (T*)dst;
    p.sp1 = d.src_ext[0]; p.sp2 = d.src_ext[0] * d.src_ext[1];
    p.dp1 = d.dst_ext[0]; p.dp2 = d.dst_ext[0] * d.dst_ext[1];
    p.X = (int)d.size[0]; p.Y = (int)d.size[1]; p.Z = (int)d.size[2];
    p.so1 = d.src_off[1]; p.so2 = d.src_off[2];
    p.do0 = d.dst_off[0]; p.do1 = d.dst_off[1]; p.do2 = d.dst_off[2];
    p.bc0 = d.boundary[0]; p.bc1 = d.boundary[1]; p.bc2 = d.boundary[2];
    memcpy(&p.pad, &d.padval_bits, sizeof(T));
    p.alpha = (T)d.alpha;
    p.z_lo = (int)pl.dd.lo[2]; p.zn = (int)pl.dd.n[2];
    p.mirror = nullptr; p.m_lo = p.m_hi = 0;
    if (g_mirror.ptr && d.dst_off[0] == 0 && d.dst_off[1] == 0 && d.dst_ext[0] == d.size[0] && d.dst_ext[1] == d.size[1] &&
        (d.reducer == SB200_DIFFUSION || d.reducer == SB200_SUM || d.reducer == SB200_MEAN || d.reducer == SB200_MAX || d.reducer == SB200_MIN)) {
        p.mirror = (T*)g_mirror.ptr; p.m_lo = (int)g_mirror.lo; p.m_hi = (int)g_mirror.hi;
        g_mirror.honoured = true;
    }
    const long long Xb = d.size[0] * (long long)sizeof(T);
    p.ntx = (int)((Xb + S3_TXB - 1) / S3_TXB);
    p.nty = (int)((d.size[1] + S3_TY - 1) / S3_TY);
    switch (d.reducer) {
    case SB200_DIFFUSION: return s3_launch<T, SB200_DIFFUSION>(p, st);
    case SB200_SUM: return s3_launch<T, SB200_SUM>(p, st);
    case SB200_MEAN: return s3_launch<T, SB200_MEAN>(p, st);
    case SB200_MAX: return s3_launch<T, SB200_MAX>(p, st);
    case SB200_MIN: return s3_launch<T, SB200_MIN>(p, st);
    default: return -1;
    }
}

int try_diffusion3d(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const sb200_desc& d = pl.d;
    if (d.flags & SB200_FLAG_NO_TMA) return -1;
    if (SB200_FLAG_GENS_OF(d.flags) == 2 && d.reducer == SB200_DIFFUSION) return try_diffusion3d_double(pl, src, dst, st);
    if (d.ndim != 3 || pl.shape_tag != SB200_VONNEUMANN || pl.shape_ndim != 3 || d.radius != 1 || d.noffsets != 6) return -1;
    int rc = -1;
    if (d.eltype == SB200_F32) rc = s3_try<float>(pl, src, dst, st);
    else if (d.eltype == SB200_F64) rc = s3_try<double>(pl, src, dst, st);
    if (rc == SB200_OK) set_kernel_name("stream3d_kernel");
    return rc;
}

}  // namespace sb
