// life.cu — Game of Life (Moore(1) neighbour count + born/survive table) on 1-byte cells, SWAR.
//
// Reference path being replaced: gatherstencil_kernel! (src/gatherstencil.jl:105-109) applied with a
// Moore{1,2} stencil (src/stencils/moore.jl) to a UInt8/Bool grid, one work-item per cell, 8 scattered loads
// each. Here one thread owns 16 consecutive cells (one 128-bit load / store per row) and walks down a run of
// rows keeping the horizontal 3-sums of the previous two rows in registers, so every source row is loaded
// once per run. Cells stay packed four to a 32-bit word: byte-wise sums never exceed 9, so plain integer
// adds cannot carry between bytes. Left/right neighbours come from funnel shifts (ptxas fuses shift+add into
// LEA.HI); the words of the neighbouring threads come from warp shuffles; only lanes 0 and 31 touch memory
// for their outer neighbour (one byte). The next row is prefetched while the current one is evaluated.
// Work is split into equal runs of rows per warp over the linearised (column group, row) space, so all
// resident warps finish together. Algorithmic traffic: 1 B read + 1 B written per cell.
//
// Handles: eltype Bool/UInt8, 2-D, any of Remove/Wrap/Reflect on the fly (Conditional) or a ring / ghost rows
// (src_off > 0) on axis 1; axis 0 must be unpadded with 16-byte-aligned rows. Everything else is declined
// and goes to gather_generic.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "tma.cuh"
#include "life_params.cuh"

namespace sb {


struct Row {
    unsigned h0, h1, h2, h3;  // horizontal 3-sums (left + centre + right), per byte
    unsigned c0, c1, c2, c3;  // centre cells as 0/1
};
struct Raw {
    uint4 v;
    unsigned bl, br;  // the cells left of byte 0 and right of byte 15 (raw bytes)
};

// Per-thread constants of the column this thread owns (hoisted out of the row loops).
struct Col {
    long long x_off;     // byte offset of the thread's 16 cells in a row
    int l_edge, r_edge;  // array-edge columns under Wrap / Reflect: byte offset in the row of the outer neighbour
    bool active;         // column exists (inactive lanes shadow the last column and do not store)
    bool first, last;    // first / last column of the array
    bool edge_mem;       // Wrap / Reflect on axis 0: the array-edge neighbour is a cell of the same row
    bool edge_warp;      // warp-uniform: this column group contains the first or the last column
    unsigned pad;        // outer neighbour when it is the Remove padval
};

// tp = row pointer + x_off (per thread); rowp = row pointer (warp-uniform). Every thread loads its 16 cells and
// its two outer neighbour bytes (same cache lines, L1 hits), so the steady state needs no cross-lane traffic.
__device__ __forceinline__ Raw load_raw(const uint8_t* __restrict__ tp, const uint8_t* __restrict__ rowp, const Col& c) {
    Raw r;
    r.v = __ldg(reinterpret_cast<const uint4*>(tp));
    r.bl = c.pad; r.br = c.pad;
    if (!c.first) r.bl = __ldg(tp - 1);
    if (!c.last) r.br = __ldg(tp + 16);
    if (c.edge_warp && c.edge_mem) {  // warp-uniform branch
        if (c.first) r.bl = __ldg(rowp + c.l_edge);
        if (c.last) r.br = __ldg(rowp + c.r_edge);
    }
    return r;
}

template <bool CELLS01>
__device__ __forceinline__ Row finish_row(const Raw& r, const Col& c) {
    unsigned w0 = r.v.x, w1 = r.v.y, w2 = r.v.z, w3 = r.v.w, bl = r.bl, br = r.br;
    if (!CELLS01) {
        w0 = nz_bytes(w0); w1 = nz_bytes(w1); w2 = nz_bytes(w2); w3 = nz_bytes(w3);
        bl = min(bl, 1u); br = min(br, 1u);
    }
    Row o;
    o.c0 = w0; o.c1 = w1; o.c2 = w2; o.c3 = w3;
    o.h0 = ((w0 << 8) + bl) + w0 + __funnelshift_r(w0, w1, 8);
    o.h1 = __funnelshift_l(w0, w1, 8) + w1 + __funnelshift_r(w1, w2, 8);
    o.h2 = __funnelshift_l(w1, w2, 8) + w2 + __funnelshift_r(w2, w3, 8);
    o.h3 = __funnelshift_l(w2, w3, 8) + w3 + __funnelshift_r(w3, br, 8);
    return o;
}

// A source row addressed through the boundary condition of axis 1 (used for the first and last row of a run).
template <bool CELLS01>
__device__ __forceinline__ Row mapped_row(const LifeParams& p, int r, const Col& c) {
    long long prow;
    if (p.soff1 > 0) {
        prow = (long long)r + p.soff1;  // ring / ghost rows: read straight through
    } else if (r < 0 || r >= p.H) {
        if (p.bc1 == SB200_WRAP) prow = r < 0 ? r + p.H : r - p.H;
        else if (p.bc1 == SB200_REFLECT) prow = r < 0 ? -r : 2 * (p.H - 1) - r;
        else {  // Remove: the whole row, and what lies left and right of it, is padval
            const unsigned pv = p.pad01 * 0x01010101u;
            Row o;
            o.c0 = o.c1 = o.c2 = o.c3 = pv;
            o.h0 = o.h1 = o.h2 = o.h3 = pv * 3u;
            return o;
        }
    } else {
        prow = r;
    }
    const uint8_t* rowp = p.src + prow * p.spitch;
    return finish_row<CELLS01>(load_raw(rowp + c.x_off, rowp, c), c);
}

// CONWAY: B3/S23 via the (n | c) == 3 identity; otherwise the general born/survive table.
template <bool CONWAY>
__device__ __forceinline__ unsigned rule(unsigned t, unsigned c, unsigned born, unsigned survive) {
    const unsigned n = t - c;  // neighbour count without the centre, 0..8 per byte
    if (CONWAY) return eqz_small((n | c) ^ 0x03030303u);
    unsigned out = 0;
#pragma unroll
    for (int s = 0; s <= 8; s++) {
        const bool b = (born >> s) & 1u, sv = (survive >> s) & 1u;
        if (b | sv) {  // warp-uniform
            const unsigned eq = eqz_small(n ^ (0x01010101u * s));
            const unsigned sel = (b ? (c ^ 0x01010101u) : 0u) | (sv ? c : 0u);
            out |= eq & sel;
        }
    }
    return out;
}

// out(row of b) from the rows above (a), at (b) and below (n); tp = dest row pointer + x_off.
template <bool CONWAY>
__device__ __forceinline__ void emit(const LifeParams& p, const Col& c, uint8_t* __restrict__ tp, const Row& a,
                                     const Row& b, const Row& n) {
    uint4 o;
    o.x = rule<CONWAY>(a.h0 + b.h0 + n.h0, b.c0, p.born, p.survive);
    o.y = rule<CONWAY>(a.h1 + b.h1 + n.h1, b.c1, p.born, p.survive);
    o.z = rule<CONWAY>(a.h2 + b.h2 + n.h2, b.c2, p.born, p.survive);
    o.w = rule<CONWAY>(a.h3 + b.h3 + n.h3, b.c3, p.born, p.survive);
    if (c.active) *reinterpret_cast<uint4*>(tp) = o;
}

template <bool CELLS01, bool CONWAY>
__global__ void __launch_bounds__(256) life_swar_kernel(const LifeParams p) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // Neighbouring warps own neighbouring column groups of the SAME run of rows, so at any moment the machine
    // streams whole rows (contiguous DRAM pages); runs are equal, so all resident warps finish together.
    const int run = (int)(warp / p.colgroups);
    for (int cg = (int)(warp % p.colgroups); cg < p.colgroups && run < p.nruns; cg += p.colgroups) {
        const int y0 = p.y_lo + (int)((long long)p.rows * run / p.nruns);
        const int y1 = p.y_lo + (int)((long long)p.rows * (run + 1) / p.nruns);
        if (y0 >= y1) break;

        Col c;
        const int col_raw = cg * 32 + lane;
        c.active = col_raw < p.ncols;
        const int col = c.active ? col_raw : p.ncols - 1;
        c.first = col == 0;
        c.last = col == p.ncols - 1;
        c.x_off = (long long)col * 16;
        c.edge_warp = cg == 0 || cg == p.colgroups - 1;
        c.edge_mem = p.bc0 != SB200_REMOVE;
        c.l_edge = p.bc0 == SB200_WRAP ? p.W - 1 : 1;
        c.r_edge = p.bc0 == SB200_WRAP ? 0 : p.W - 2;
        c.pad = p.pad01;

        // Rows A, B, C rotate through the roles (above, centre, below) so the steady state moves no registers.
        Row A = mapped_row<CELLS01>(p, y0 - 1, c);
        Row B = mapped_row<CELLS01>(p, y0, c);
        Row C;
        const uint8_t* __restrict__ sp = p.src + (long long)(y0 + 1 + p.soff1) * p.spitch;  // row y+1 (uniform)
        const uint8_t* __restrict__ st = sp + c.x_off;                                       // same, this thread
        uint8_t* __restrict__ dt = p.dst + (long long)(y0 + p.doff1) * p.dpitch + c.x_off;  // dest row y, this thread
        int n = y1 - 1 - y0;  // rows y0+1 .. y1-1 are ordinary in-run rows
        if (n > 0) {
            Raw r0 = load_raw(st, sp, c), r1, r2;
            // steady state: three rows per trip, the load of row y+2 is in flight while row y is evaluated
            while (n > 3) {
                sp += p.spitch; st += p.spitch;
                r1 = load_raw(st, sp, c);
                C = finish_row<CELLS01>(r0, c);
                emit<CONWAY>(p, c, dt, A, B, C);
                dt += p.dpitch;
                sp += p.spitch; st += p.spitch;
                r2 = load_raw(st, sp, c);
                A = finish_row<CELLS01>(r1, c);
                emit<CONWAY>(p, c, dt, B, C, A);
                dt += p.dpitch;
                sp += p.spitch; st += p.spitch;
                r0 = load_raw(st, sp, c);
                B = finish_row<CELLS01>(r2, c);
                emit<CONWAY>(p, c, dt, C, A, B);
                dt += p.dpitch;
                n -= 3;
            }
            // 1..3 ordinary rows left; r0 holds the first of them
            while (true) {
                const bool more = n > 1;
                if (more) { sp += p.spitch; st += p.spitch; r1 = load_raw(st, sp, c); }
                C = finish_row<CELLS01>(r0, c);
                emit<CONWAY>(p, c, dt, A, B, C);
                dt += p.dpitch;
                A = B; B = C; r0 = r1;
                if (!more) break;
                n--;
            }
        }
        C = mapped_row<CELLS01>(p, y1, c);
        emit<CONWAY>(p, c, dt, A, B, C);
    }
}

template <bool CELLS01, bool CONWAY> static int resident_blocks() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, life_swar_kernel<CELLS01, CONWAY>, 256, 0) != cudaSuccess || per_sm < 1)
            per_sm = 4;
        cached = per_sm * num_sms();
        cached_dev = dev;
    }
    return cached;
}

template <bool CELLS01, bool CONWAY> static void launch(LifeParams& p, cudaStream_t st) {
    const long long warps = (long long)resident_blocks<CELLS01, CONWAY>() * 8;
    long long nruns = std::max<long long>(1, warps / p.colgroups);
    nruns = std::min<long long>(nruns, std::max(1, p.rows / 8));  // at least 8 rows per run
    p.nruns = (int)nruns;
    const long long blocks = (nruns * p.colgroups + 7) / 8;
    life_swar_kernel<CELLS01, CONWAY><<<(unsigned)blocks, 256, 0, st>>>(p);
}

// ------------------------------------------------------------------------------------------------------------
// TMA-fed variant. A CTA owns a strip of LT_WARPS*512 bytes of columns and streams down a run of rows. One
// producer thread issues bulk copies (cp.async.bulk -> UBLKCP) of LT_CH source rows per stage into a ring of
// LT_STAGES shared-memory stages and resolves the row boundary (Wrap / Reflect / ghost rows) and the Wrap
// column halo purely by choosing source addresses; full/empty mbarriers hand stages back and forth. The
// consumer warps read their 16 cells + 2 neighbour bytes from shared memory, keep the rolling 3-row state in
// registers and store results with 128-bit coalesced stores. Bytes in flight are set by the ring
// (LT_STAGES * LT_CH * 4 KiB per CTA), not by registers.
constexpr int LT_WARPS = 8;                    // consumer warps per CTA
constexpr int LT_BXB = LT_WARPS * 32 * 16;     // strip width in bytes (4096)
constexpr int LT_ROWB = LT_BXB + 32;           // shared-memory row: 16 B left halo | strip | 16 B right halo
constexpr int LT_CH = 6;                       // source rows per stage (multiple of 3: the row rotation closes)
constexpr int LT_STAGES = 4;
constexpr int LT_SMEM = 128 + LT_STAGES * LT_CH * LT_ROWB;


// MIRROR: the fused ghost-row push (LifeParams::mirror) is compiled in only for the boundary sweeps of slab runs.
template <bool CELLS01, bool CONWAY, bool MIRROR>
__global__ void __launch_bounds__((LT_WARPS + 1) * 32) life_tma_kernel(const LifeTmaParams q) {
    extern __shared__ __align__(128) uint8_t smem[];
    const LifeParams& p = q.lp;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + LT_STAGES;
    uint8_t* ring = smem + 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < LT_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], LT_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    const int ntasks = q.nstrips * q.nruns;
    unsigned k = 0;  // stage-use counter, advances identically in the producer and in every consumer
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int strip = task % q.nstrips, run = task / q.nstrips;
        const int x0 = strip * LT_BXB;
        const int wbytes = min(LT_BXB, p.W - x0);
        const int y0 = p.y_lo + (int)((long long)p.rows * run / q.nruns);
        const int y1 = p.y_lo + (int)((long long)p.rows * (run + 1) / q.nruns);
        const int nsrc = y1 - y0 + 2;  // source rows y0-1 .. y1
        const int nchunks = (nsrc + LT_CH - 1) / LT_CH;
        if (warp == LT_WARPS) {
            // ---------------- producer ----------------
            if (lane == 0) {
                const bool lh = x0 > 0 || p.bc0 == SB200_WRAP;              // left halo comes from memory
                const bool rh = x0 + wbytes < p.W || p.bc0 == SB200_WRAP;   // right halo comes from memory
                const int lx = x0 > 0 ? x0 - 16 : p.W - 16;
                const int rx = x0 + wbytes < p.W ? x0 + wbytes : 0;
                for (int c = 0; c < nchunks; c++, k++) {
                    const int slot = k % LT_STAGES;
                    mbar_wait_producer(&empty[slot], ((k / LT_STAGES) & 1) ^ 1);
                    uint8_t* sbase = ring + slot * (LT_CH * LT_ROWB);
                    unsigned bytes = 0;
                    long long prow[LT_CH];
#pragma unroll
                    for (int j = 0; j < LT_CH; j++) {
                        const int i = c * LT_CH + j;
                        prow[j] = i < nsrc ? life_map_row(p, y0 - 1 + i) : -1;
                        if (prow[j] >= 0) bytes += wbytes + (lh ? 16 : 0) + (rh ? 16 : 0);
                    }
                    mbar_arrive_expect_tx(&full[slot], bytes);
#pragma unroll
                    for (int j = 0; j < LT_CH; j++) {
                        if (prow[j] < 0) continue;
                        const uint8_t* g = p.src + prow[j] * p.spitch;
                        uint8_t* srow = sbase + j * LT_ROWB;
                        bulk_g2s(srow + 16, g + x0, wbytes, &full[slot]);
                        if (lh) bulk_g2s(srow, g + lx, 16, &full[slot]);
                        if (rh) bulk_g2s(srow + 16 + wbytes, g + rx, 16, &full[slot]);
                    }
                }
            } else {
                k += nchunks;
            }
            continue;
        }
        // ---------------- consumers ----------------
        const int xt = (warp * 32 + lane) * 16;  // byte offset of this thread's cells inside the strip
        const bool active = xt < wbytes;
        const bool first = x0 + xt == 0, last = x0 + xt + 16 == p.W;
        const bool lfix = first && p.bc0 != SB200_WRAP, rfix = last && p.bc0 != SB200_WRAP;
        const bool reflect = p.bc0 == SB200_REFLECT;
        uint8_t* __restrict__ dt = p.dst + (long long)(y0 + p.doff1) * p.dpitch + x0 + xt;
        long long moff = (long long)(y0 - p.m_lo) * p.dpitch + x0 + xt;  // offset of output row y0 in the mirror
        int yout = y0;
        const unsigned padrow = p.pad01 * 0x01010101u;
        Row A, B, C;
        A.h0 = A.h1 = A.h2 = A.h3 = B.h0 = B.h1 = B.h2 = B.h3 = 0;
        B.c0 = B.c1 = B.c2 = B.c3 = 0;
        // One source row: build its Row from shared memory (or the Remove pad row).
        auto take = [&](const uint8_t* srow, bool is_pad, Row& o) {
            if (is_pad) {
                o.c0 = o.c1 = o.c2 = o.c3 = padrow;
                o.h0 = o.h1 = o.h2 = o.h3 = padrow * 3u;
                return;
            }
            const uint8_t* t = srow + 16 + xt;
            const uint4 v = *reinterpret_cast<const uint4*>(t);
            unsigned w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w, bl = t[-1], br = t[16];
            if (lfix) bl = reflect ? (w0 >> 8) & 0xFFu : p.pad01;
            if (rfix) br = reflect ? (w3 >> 16) & 0xFFu : p.pad01;
            if (!CELLS01) {
                w0 = nz_bytes(w0); w1 = nz_bytes(w1); w2 = nz_bytes(w2); w3 = nz_bytes(w3);
                bl = min(bl, 1u); br = min(br, 1u);
            }
            o.c0 = w0; o.c1 = w1; o.c2 = w2; o.c3 = w3;
            o.h0 = ((w0 << 8) + bl) + w0 + __funnelshift_r(w0, w1, 8);
            o.h1 = __funnelshift_l(w0, w1, 8) + w1 + __funnelshift_r(w1, w2, 8);
            o.h2 = __funnelshift_l(w1, w2, 8) + w2 + __funnelshift_r(w2, w3, 8);
            o.h3 = __funnelshift_l(w2, w3, 8) + w3 + __funnelshift_r(w3, br, 8);
        };
        auto put = [&](int i, const Row& a, const Row& b, const Row& n) {  // i = index of the newest source row
            if (i >= 2 && i < nsrc) {
                uint4 o;
                o.x = rule<CONWAY>(a.h0 + b.h0 + n.h0, b.c0, p.born, p.survive);
                o.y = rule<CONWAY>(a.h1 + b.h1 + n.h1, b.c1, p.born, p.survive);
                o.z = rule<CONWAY>(a.h2 + b.h2 + n.h2, b.c2, p.born, p.survive);
                o.w = rule<CONWAY>(a.h3 + b.h3 + n.h3, b.c3, p.born, p.survive);
                if (active) *reinterpret_cast<uint4*>(dt) = o;
                if (MIRROR) {
                    if (active && yout >= p.m_lo && yout < p.m_hi)   // boundary rows cross NVLink as they are produced
                        *reinterpret_cast<uint4*>(p.mirror + moff) = o;
                    moff += p.dpitch; yout++;
                }
                dt += p.dpitch;
            }
        };
        const bool may_pad = p.soff1 == 0 && p.bc1 == SB200_REMOVE;
        for (int c = 0; c < nchunks; c++, k++) {
            const int slot = k % LT_STAGES;
            mbar_wait(&full[slot], (k / LT_STAGES) & 1);
            const uint8_t* sbase = ring + slot * (LT_CH * LT_ROWB);
            const int i0 = c * LT_CH;
            // pad rows can only be the first row of the first chunk or the last source row
            const bool pad_first = may_pad && c == 0 && y0 - 1 < 0;
            const int pad_last_i = (may_pad && y1 >= p.H) ? nsrc - 1 : -1;
#pragma unroll
            for (int j = 0; j < LT_CH; j += 3) {
                take(sbase + (j + 0) * LT_ROWB, (j == 0 && pad_first) || i0 + j + 0 == pad_last_i, C);
                put(i0 + j + 0, A, B, C);
                take(sbase + (j + 1) * LT_ROWB, i0 + j + 1 == pad_last_i, A);
                put(i0 + j + 1, B, C, A);
                take(sbase + (j + 2) * LT_ROWB, i0 + j + 2 == pad_last_i, B);
                put(i0 + j + 2, C, A, B);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
    }
}

template <bool CELLS01, bool CONWAY, bool MIRROR> static int launch_tma_m(const LifeParams& p, cudaStream_t st) {
    static thread_local int cfg_dev = -1, ctas_per_sm = 0;
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (dev != cfg_dev) {
        SB_CUDA(cudaFuncSetAttribute(life_tma_kernel<CELLS01, CONWAY, MIRROR>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, life_tma_kernel<CELLS01, CONWAY, MIRROR>, (LT_WARPS + 1) * 32, LT_SMEM) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        ctas_per_sm = per_sm;
        cfg_dev = dev;
    }
    LifeTmaParams q;
    q.lp = p;
    q.nstrips = (p.W + LT_BXB - 1) / LT_BXB;
    const long long ctas = (long long)ctas_per_sm * num_sms();
    long long nruns = std::max<long long>(1, ctas / q.nstrips);
    nruns = std::min<long long>(nruns, std::max(1, p.rows / 16));  // at least 16 rows per run
    q.nruns = (int)nruns;
    const long long grid = std::min<long long>(ctas, (long long)q.nstrips * q.nruns);
    life_tma_kernel<CELLS01, CONWAY, MIRROR><<<(unsigned)grid, (LT_WARPS + 1) * 32, LT_SMEM, st>>>(q);
    return SB200_OK;
}
template <bool CELLS01, bool CONWAY> static int launch_tma(const LifeParams& p, cudaStream_t st) {
    return p.mirror ? launch_tma_m<CELLS01, CONWAY, true>(p, st) : launch_tma_m<CELLS01, CONWAY, false>(p, st);
}

// ------------------------------------------------------------------------------------------------------------
// Two generations per sweep (SB200_FLAG_DOUBLE_STEP): dest = step(step(src)) with ONE read and ONE write of the grid.
// Same TMA ring as life_tma_kernel. A warp's 32 lanes hold 512 consecutive cells of the first-generation row; the
// left / right neighbour bytes of that intermediate row come from the adjacent lanes by warp shuffle, so only lanes
// 1..30 own final cells: warps overlap by two lanes and a strip is LT2_WARPS * 480 bytes wide (6 % recomputation).
// The intermediate generation lives only in registers (three rolling rows per thread), never in memory.
// Supported where evolving the halo equals the boundary rule: Wrap on axis 0, Wrap / Reflect (or a region that stays
// inside the parent) on axis 1, no ring. Algorithmic traffic: 1 B read + 1 B written per cell per TWO generations.
constexpr int LT2_WARPS = 8;
constexpr int LT2_OUTB = LT2_WARPS * 480;      // final cells per strip row
constexpr int LT2_D0 = 96;                     // data start inside a shared-memory row: global x0-32 is 96 mod 128
constexpr int LT2_ROWB = 4096;                 // 96 + 32 + 3840 + 32 = 4000, padded
constexpr int LT2_CH = 6;
constexpr int LT2_STAGES = 3;
constexpr int LT2_SMEM = 128 + LT2_STAGES * LT2_CH * LT2_ROWB;

__device__ __forceinline__ Row row_of_words(unsigned w0, unsigned w1, unsigned w2, unsigned w3, unsigned bl, unsigned br) {
    Row o;
    o.c0 = w0; o.c1 = w1; o.c2 = w2; o.c3 = w3;
    o.h0 = ((w0 << 8) + bl) + w0 + __funnelshift_r(w0, w1, 8);
    o.h1 = __funnelshift_l(w0, w1, 8) + w1 + __funnelshift_r(w1, w2, 8);
    o.h2 = __funnelshift_l(w1, w2, 8) + w2 + __funnelshift_r(w2, w3, 8);
    o.h3 = __funnelshift_l(w2, w3, 8) + w3 + __funnelshift_r(w3, br, 8);
    return o;
}

template <bool CELLS01, bool CONWAY>
__global__ void __launch_bounds__((LT2_WARPS + 1) * 32, 3) life_tma2_kernel(const LifeTmaParams q) {
    extern __shared__ __align__(128) uint8_t smem[];
    const LifeParams& p = q.lp;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + LT2_STAGES;
    uint8_t* ring = smem + 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < LT2_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], LT2_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    const int ntasks = q.nstrips * q.nruns;
    unsigned k = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int strip = task % q.nstrips, run = task / q.nstrips;
        const int x0 = strip * q.outb;
        const int wout = min(q.outb, p.W - x0);
        const int y0 = p.y_lo + (int)((long long)p.rows * run / q.nruns);
        const int y1 = p.y_lo + (int)((long long)p.rows * (run + 1) / q.nruns);
        const int nsrc = y1 - y0 + 4;  // source rows y0-2 .. y1+1
        const int nchunks = (nsrc + LT2_CH - 1) / LT2_CH;
        if (warp == LT2_WARPS) {
            // ---------------- producer: shared-memory row byte b <-> global column x0 - 32 + b (mod W) ----------------
            if (lane == 0) {
                const int lin = x0 >= 32 ? 32 : 0;                       // left halo bytes that are ordinary cells
                const int rin = min(32, p.W - (x0 + wout));              // right halo bytes that are ordinary cells
                const unsigned mlen = lin + wout + rin;                  // one contiguous copy
                const unsigned rowbytes = 64 + wout;
                for (int c = 0; c < nchunks; c++, k++) {
                    const int slot = k % LT2_STAGES;
                    mbar_wait_producer(&empty[slot], ((k / LT2_STAGES) & 1) ^ 1);
                    uint8_t* sbase = ring + slot * (LT2_CH * LT2_ROWB);
                    const int nrows = min(LT2_CH, nsrc - c * LT2_CH);
                    mbar_arrive_expect_tx(&full[slot], nrows * rowbytes);
                    for (int j = 0; j < nrows; j++) {
                        const uint8_t* g = p.src + life_map_row(p, y0 - 2 + c * LT2_CH + j) * p.spitch;
                        uint8_t* srow = sbase + j * LT2_ROWB + LT2_D0;
                        bulk_g2s(srow + 32 - lin, g + x0 - lin, mlen, &full[slot]);
                        if (!lin) bulk_g2s(srow, g + p.W - 32, 32, &full[slot]);                        // wrapped left halo
                        if (rin < 32) bulk_g2s(srow + 32 + wout + rin, g, 32 - rin, &full[slot]);       // wrapped right halo
                    }
                }
            } else {
                k += nchunks;
            }
            continue;
        }
        // ---------------- consumers ----------------
        const int cell0 = warp * 480 + (lane - 1) * 16;          // first final cell of this lane inside the strip (may be < 0)
        const bool active = lane >= 1 && lane <= 30 && cell0 < wout;
        if (warp * 480 >= wout) {
            // no final cell in this warp (strips are W / nstrips wide, not always 8 x 480): keep the ring protocol only
            for (int c = 0; c < nchunks; c++, k++) {
                const int slot = k % LT2_STAGES;
                mbar_wait(&full[slot], (k / LT2_STAGES) & 1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
            }
            continue;
        }
        uint8_t* __restrict__ dt = p.dst + (long long)(y0 + p.doff1) * p.dpitch + x0 + cell0;
        const int soff = LT2_D0 + 32 + cell0;                    // offset of the lane's 16 cells in a shared-memory row
        Row S0, S1, S2, T0, T1, T2;
        S0 = S1 = S2 = T0 = T1 = T2 = Row{0, 0, 0, 0, 0, 0, 0, 0};
        auto take = [&](const uint8_t* srow, Row& o) {
            const uint8_t* t = srow + soff;
            const uint4 v = *reinterpret_cast<const uint4*>(t);
            unsigned w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w, bl = t[-1], br = t[16];
            if (!CELLS01) {
                w0 = nz_bytes(w0); w1 = nz_bytes(w1); w2 = nz_bytes(w2); w3 = nz_bytes(w3);
                bl = min(bl, 1u); br = min(br, 1u);
            }
            o = row_of_words(w0, w1, w2, w3, bl, br);
        };
        // first generation of the row of `b` from source rows a, b, n; neighbour bytes from the adjacent lanes
        auto gen1 = [&](const Row& a, const Row& b, const Row& n, Row& o) {
            const unsigned x0_ = rule<CONWAY>(a.h0 + b.h0 + n.h0, b.c0, p.born, p.survive);
            const unsigned x1_ = rule<CONWAY>(a.h1 + b.h1 + n.h1, b.c1, p.born, p.survive);
            const unsigned x2_ = rule<CONWAY>(a.h2 + b.h2 + n.h2, b.c2, p.born, p.survive);
            const unsigned x3_ = rule<CONWAY>(a.h3 + b.h3 + n.h3, b.c3, p.born, p.survive);
            const unsigned bl = __shfl_up_sync(0xffffffffu, x3_, 1) >> 24;
            const unsigned br = __shfl_down_sync(0xffffffffu, x0_, 1) & 0xFFu;
            o = row_of_words(x0_, x1_, x2_, x3_, bl, br);
        };
        auto gen2 = [&](int i, const Row& a, const Row& b, const Row& n) {  // i = stream index of the newest source row
            if (i >= 4 && i < nsrc) {
                uint4 o;
                o.x = rule<CONWAY>(a.h0 + b.h0 + n.h0, b.c0, p.born, p.survive);
                o.y = rule<CONWAY>(a.h1 + b.h1 + n.h1, b.c1, p.born, p.survive);
                o.z = rule<CONWAY>(a.h2 + b.h2 + n.h2, b.c2, p.born, p.survive);
                o.w = rule<CONWAY>(a.h3 + b.h3 + n.h3, b.c3, p.born, p.survive);
                if (active) *reinterpret_cast<uint4*>(dt) = o;
                dt += p.dpitch;
            }
        };
        for (int c = 0; c < nchunks; c++, k++) {
            const int slot = k % LT2_STAGES;
            mbar_wait(&full[slot], (k / LT2_STAGES) & 1);
            const uint8_t* sbase = ring + slot * (LT2_CH * LT2_ROWB);
            const int i0 = c * LT2_CH;
            // stream index i: S_i -> slot i%3;  T_{i-1} = gen1(S_{i-2}, S_{i-1}, S_i) -> slot (i-1)%3;  out(i-2) = gen2(T_{i-3}, T_{i-2}, T_{i-1})
#pragma unroll
            for (int j = 0; j < LT2_CH; j += 3) {
                take(sbase + (j + 0) * LT2_ROWB, S0);   // i % 3 == 0
                gen1(S1, S2, S0, T2);                   // T_{i-1}, (i-1) % 3 == 2
                gen2(i0 + j + 0, T0, T1, T2);
                take(sbase + (j + 1) * LT2_ROWB, S1);   // i % 3 == 1
                gen1(S2, S0, S1, T0);
                gen2(i0 + j + 1, T1, T2, T0);
                take(sbase + (j + 2) * LT2_ROWB, S2);   // i % 3 == 2
                gen1(S0, S1, S2, T1);
                gen2(i0 + j + 2, T2, T0, T1);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Can this sweep run two generations per launch?
bool life2_accepts(const sb200_desc& d, const Plan& pl) { return life_multi_accepts(d, pl, 2); }

bool life_multi_accepts(const sb200_desc& d, const Plan& pl, int gens) {
    if (gens > 8 || gens < ((d.flags & SB200_FLAG_SRC_BITS) ? 1 : 2)) return false;   // a packed source also runs single generations
    if (gens >= 3 && (d.born_mask != (1u << 3) || d.survive_mask != ((1u << 2) | (1u << 3)) || d.size[0] % 32 || getenv("SB200_NO_BITSLICE")))
        return false;   // three and more generations: the bit-sliced B3/S23 kernel only
    if (gens > 4 && !SB200_LB_ONE_HALO_LANE) return false;   // five to eight generations need the one-halo-lane layout
    if (d.flags & (SB200_FLAG_SRC_BITS | SB200_FLAG_DST_BITS)) {   // packed state: the bit-sliced kernel, rows of whole 16-byte groups
        if (!SB200_LB_ONE_HALO_LANE || d.born_mask != (1u << 3) || d.survive_mask != ((1u << 2) | (1u << 3)) || getenv("SB200_NO_BITSLICE")) return false;
        if (d.size[0] % 128) return false;
        if ((d.flags & SB200_FLAG_SRC_BITS) && d.src_ext[0] % 128) return false;
        if ((d.flags & SB200_FLAG_DST_BITS) && d.dst_ext[0] % 128) return false;
    }
    if (d.reducer != SB200_LIFE || d.ndim != 2 || (d.eltype != SB200_BOOL && d.eltype != SB200_U8)) return false;
    if (pl.shape_tag != SB200_MOORE || pl.shape_ndim != 2 || d.radius != 1 || d.noffsets != 8) return false;
    if (d.flags & (SB200_FLAG_NO_TMA | SB200_FLAG_FORCE_GENERIC)) return false;
    for (int a = 0; a < 2; a++)
        if (d.src_off[a] != 0 || d.dst_off[a] != 0) return false;
    if (d.size[0] % 16 || d.src_ext[0] % 16 || d.dst_ext[0] % 16 || d.size[0] < 512 || d.size[0] > (1LL << 30) || d.size[1] > (1LL << 30)) return false;
    if (d.boundary[0] != SB200_WRAP) return false;
    if (pl.dd.lo[0] != 0 || pl.dd.n[0] != d.size[0] || pl.dd.n[1] < 16) return false;
    const long long lo = pl.dd.lo[1], hi = lo + pl.dd.n[1];
    const bool inside = lo >= gens && hi + gens <= d.size[1];  // never leaves the parent: the boundary rule of axis 1 is not exercised
    if (!inside && d.boundary[1] != SB200_WRAP && d.boundary[1] != SB200_REFLECT) return false;
    if (d.size[1] < 2 * gens) return false;
    return true;
}

template <bool CELLS01, bool CONWAY> static int launch_tma2(const LifeParams& p, cudaStream_t st) {
    static thread_local int cfg_dev = -1, ctas_per_sm = 0;
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (dev != cfg_dev) {
        SB_CUDA(cudaFuncSetAttribute(life_tma2_kernel<CELLS01, CONWAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT2_SMEM));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, life_tma2_kernel<CELLS01, CONWAY>, (LT2_WARPS + 1) * 32, LT2_SMEM) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        ctas_per_sm = per_sm;
        cfg_dev = dev;
    }
    LifeTmaParams q;
    q.lp = p;
    q.nstrips = (p.W + LT2_OUTB - 1) / LT2_OUTB;
    q.outb = std::min(LT2_OUTB, ((p.W + q.nstrips - 1) / q.nstrips + 127) / 128 * 128);  // equal strips
    q.nstrips = (p.W + q.outb - 1) / q.outb;
    const long long ctas = (long long)ctas_per_sm * num_sms();
    long long nruns = std::max<long long>(1, 2 * ctas / q.nstrips);  // at most two tasks per CTA: no straggler third task
    nruns = std::min<long long>(nruns, std::max(1, p.rows / 32));  // at least 32 rows per run (4 are re-read)
    q.nruns = (int)nruns;
    const long long grid = std::min<long long>(ctas, (long long)q.nstrips * q.nruns);
    life_tma2_kernel<CELLS01, CONWAY><<<(unsigned)grid, (LT2_WARPS + 1) * 32, LT2_SMEM, st>>>(q);
    return SB200_OK;
}

int launch_life_bit_u8(int gens, bool cells01, const LifeParams& p, cudaStream_t st);        // life_bit_u8.cu
int launch_life_bit_to_bits(int gens, bool cells01, const LifeParams& p, cudaStream_t st);   // life_bit_pk_a.cu
int launch_life_bit_from_bits(int gens, bool out_bits, const LifeParams& p, cudaStream_t st); // life_bit_pk_b.cu

int try_life_swar(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const sb200_desc& d = pl.d;
    if (d.reducer != SB200_LIFE || d.ndim != 2) return -1;
    if (d.eltype != SB200_BOOL && d.eltype != SB200_U8) return -1;
    if (pl.shape_tag != SB200_MOORE || pl.shape_ndim != 2 || d.radius != 1 || d.noffsets != 8) return -1;
    if (d.src_off[0] != 0 || d.dst_off[0] != 0) return -1;
    if (d.size[0] % 16 || d.src_ext[0] % 16 || d.dst_ext[0] % 16) return -1;
    if (((uintptr_t)src | (uintptr_t)dst) & 15) return -1;
    if (d.size[0] > (1LL << 30) || d.size[1] > (1LL << 30) || d.size[0] < 16) return -1;
    if (pl.dd.lo[0] != 0 || pl.dd.n[0] != d.size[0]) return -1;  // regions only along axis 1
    if (d.src_off[1] == 0 && d.boundary[1] == SB200_USE) return -1;
    if (pl.dd.n[1] == 0) return SB200_OK;

    LifeParams p;
    p.src = (const uint8_t*)src; p.dst = (uint8_t*)dst;
    p.spitch = d.src_ext[0]; p.dpitch = d.dst_ext[0];
    p.W = (int)d.size[0]; p.H = (int)d.size[1]; p.ncols = p.W / 16;
    p.colgroups = (p.ncols + 31) / 32;
    p.soff1 = d.src_off[1]; p.doff1 = d.dst_off[1];
    p.bc0 = d.boundary[0]; p.bc1 = d.boundary[1];
    p.pad01 = (d.padval_bits & 0xFF) != 0;
    p.y_lo = (int)pl.dd.lo[1]; p.rows = (int)pl.dd.n[1];
    p.born = d.born_mask; p.survive = d.survive_mask;
    const bool conway = d.born_mask == (1u << 3) && d.survive_mask == ((1u << 2) | (1u << 3));
    // Bool cells are 0/1 by type; UInt8 cells are 0/1 when the caller says so (sb200_iterate does for every
    // step after the first, because the source is then this kernel's own output).
    const bool cells01 = d.eltype == SB200_BOOL || (d.flags & SB200_FLAG_CELLS_01);
    const int gens = SB200_FLAG_GENS_OF(d.flags);
    const bool src_bits = (d.flags & SB200_FLAG_SRC_BITS) != 0, dst_bits = (d.flags & SB200_FLAG_DST_BITS) != 0;
    if (dst_bits && !src_bits && gens < 2) { set_error("SB200_FLAG_DST_BITS on a byte source needs SB200_FLAG_GENS(2 .. 8)"); return SB200_EUNSUPPORTED; }
    if (gens > 2 || src_bits || dst_bits) {   // the bit-sliced kernel (5 .. 8 generations and the packed formats need its one-halo-lane layout, the default build)
        if (!life_multi_accepts(d, pl, gens)) { set_error("%d generations per sweep%s: layout / boundary / rule / build not supported", gens, src_bits || dst_bits ? " on packed state" : ""); return SB200_EUNSUPPORTED; }
        p.mirror = nullptr; p.m_lo = p.m_hi = 0;
        if (src_bits) p.spitch = d.src_ext[0] / 8;   // bytes per packed row
        if (dst_bits) p.dpitch = d.dst_ext[0] / 8;
        const int rc = src_bits ? launch_life_bit_from_bits(gens, dst_bits, p, st)
                                : dst_bits ? launch_life_bit_to_bits(gens, cells01, p, st) : launch_life_bit_u8(gens, cells01, p, st);
        if (rc) return rc;
        SB_LAUNCH_CHECK();
        static thread_local char name[64];
        snprintf(name, sizeof(name), "life_bit_kernel<%d,%s%s>", gens, src_bits ? "bits" : cells01 ? "cells01" : "u8", dst_bits ? "->bits" : src_bits ? "->u8" : "");
        set_kernel_name(name);
        return SB200_OK;
    }
    if (gens == 2) {
        if (!life2_accepts(d, pl)) { set_error("two generations per sweep: layout / boundary not supported"); return SB200_EUNSUPPORTED; }
        p.mirror = nullptr; p.m_lo = p.m_hi = 0;
        int rc;
        if (conway && p.W % 32 == 0 && !getenv("SB200_NO_BITSLICE")) {
            // B3/S23: the bit-sliced kernel (one bit per cell inside the kernel)
            rc = launch_life_bit_u8(2, cells01, p, st);
            if (rc) return rc;
            SB_LAUNCH_CHECK();
            set_kernel_name(cells01 ? "life_bit_kernel<2,cells01>" : "life_bit_kernel<2,u8>");
            return SB200_OK;
        }
        if (cells01) rc = conway ? launch_tma2<true, true>(p, st) : launch_tma2<true, false>(p, st);
        else rc = conway ? launch_tma2<false, true>(p, st) : launch_tma2<false, false>(p, st);
        if (rc) return rc;
        SB_LAUNCH_CHECK();
        set_kernel_name(conway ? (cells01 ? "life_tma2_kernel<cells01,conway>" : "life_tma2_kernel<u8,conway>")
                               : (cells01 ? "life_tma2_kernel<cells01,table>" : "life_tma2_kernel<u8,table>"));
        return SB200_OK;
    }
    const bool use_tma = !(d.flags & SB200_FLAG_NO_TMA) && p.W >= 512 && p.rows >= 16;
    p.mirror = nullptr; p.m_lo = p.m_hi = 0;
    if (use_tma && g_mirror.ptr && d.dst_ext[0] == d.size[0]) {
        p.mirror = (uint8_t*)g_mirror.ptr; p.m_lo = (int)g_mirror.lo; p.m_hi = (int)g_mirror.hi;
        g_mirror.honoured = true;
    }
    if (use_tma) {
        int rc;
        if (cells01) rc = conway ? launch_tma<true, true>(p, st) : launch_tma<true, false>(p, st);
        else rc = conway ? launch_tma<false, true>(p, st) : launch_tma<false, false>(p, st);
        if (rc) return rc;
        SB_LAUNCH_CHECK();
        set_kernel_name(conway ? (cells01 ? "life_tma_kernel<cells01,conway>" : "life_tma_kernel<u8,conway>")
                               : (cells01 ? "life_tma_kernel<cells01,table>" : "life_tma_kernel<u8,table>"));
        return SB200_OK;
    }
    if (cells01) { if (conway) launch<true, true>(p, st); else launch<true, false>(p, st); }
    else { if (conway) launch<false, true>(p, st); else launch<false, false>(p, st); }
    SB_LAUNCH_CHECK();
    set_kernel_name(conway ? (cells01 ? "life_swar_kernel<cells01,conway>" : "life_swar_kernel<u8,conway>")
                           : (cells01 ? "life_swar_kernel<cells01,table>" : "life_swar_kernel<u8,table>"));
    return SB200_OK;
}

}  // namespace sb
