// life.cu — Game of Life (Moore(1) neighbour count + born/survive table) on 1-byte cells, SWAR.
//
// Reference path being replaced: gatherstencil_kernel! (src/gatherstencil.jl:105-109) applied with a
// Moore{1,2} stencil (src/stencils/moore.jl) to a UInt8/Bool grid, one work-item per cell, 8 scattered loads
// each. Here one thread owns 16 consecutive cells (one 128-bit load / store per row) and walks down a band
// of rows keeping the horizontal 3-sums of the previous two rows in registers, so every source row is
// loaded once per band. Cells stay packed four to a 32-bit word: byte-wise sums never exceed 9, so plain
// integer adds cannot carry between bytes. Left/right neighbours come from funnel shifts; the words of the
// neighbouring threads come from warp shuffles. Algorithmic traffic: 1 B read + 1 B written per cell.
//
// Handles: eltype Bool/UInt8, 2-D, any of Remove/Wrap/Reflect on the fly (Conditional) or a ring/ghost rows
// (src_off > 0) on axis 1; axis 0 must be unpadded with 16-byte-aligned rows. Everything else is declined
// and goes to gather_generic.
#include "common.cuh"

namespace sb {

struct LifeParams {
    const uint8_t* src;
    uint8_t* dst;
    long long spitch, dpitch;  // bytes per row of the parents
    int W, H;                  // logical size (axis 0 = W contiguous)
    int ncols;                 // W / 16
    int soff1, doff1;          // ring / ghost rows on axis 1
    int bc0, bc1;              // boundary per axis
    unsigned pad01;            // Remove: (padval != 0)
    int y_lo, y_hi;            // output rows [y_lo, y_hi)
    int RY;                    // rows per band
    int nbands;
    unsigned born, survive;
};

// 0/1 per byte: byte != 0
__device__ __forceinline__ unsigned nz_bytes(unsigned w) {
    return ((((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) >> 7) & 0x01010101u;
}
// 0/1 per byte: byte == 0, valid for bytes <= 0x70
__device__ __forceinline__ unsigned eqz_small(unsigned x) {
    return (((x + 0x0F0F0F0Fu) >> 4) & 0x01010101u) ^ 0x01010101u;
}

struct Row {
    unsigned h[4];  // horizontal 3-sums (left + centre + right), per byte
    unsigned c[4];  // centre cells as 0/1
};

template <bool IS_BOOL>
__device__ __forceinline__ void load_row(const LifeParams& p, int r, int col, bool active, int lane, Row& out) {
    // map the source row
    bool oob_row = false;
    long long prow;
    if (p.soff1 > 0) {
        prow = (long long)r + p.soff1;
    } else if (r < 0 || r >= p.H) {
        if (p.bc1 == SB200_WRAP) prow = r < 0 ? r + p.H : r - p.H;
        else if (p.bc1 == SB200_REFLECT) prow = r < 0 ? -r : 2 * (p.H - 1) - r;
        else { prow = 0; oob_row = true; }
    } else {
        prow = r;
    }
    unsigned w0 = 0, w1 = 0, w2 = 0, w3 = 0, wl = 0, wr = 0;
    if (oob_row) {  // Remove: the whole row (and what lies left/right of it) is padval
        const unsigned pv = p.pad01 * 0x01010101u;
        out.c[0] = out.c[1] = out.c[2] = out.c[3] = pv;
        out.h[0] = out.h[1] = out.h[2] = out.h[3] = pv * 3u;
        return;
    }
    const uint8_t* rowp = p.src + prow * p.spitch;
    const bool first = col == 0, last = col == p.ncols - 1;
    if (active) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(rowp) + col);
        w0 = v.x; w1 = v.y; w2 = v.z; w3 = v.w;
        // words owned by another warp (or across the array edge)
        if (lane == 0 && !first) wl = __ldg(reinterpret_cast<const unsigned*>(rowp) + col * 4 - 1);
        if ((lane == 31 || last) && !last) wr = __ldg(reinterpret_cast<const unsigned*>(rowp) + col * 4 + 4);
        if (first && p.bc0 == SB200_WRAP) wl = __ldg(reinterpret_cast<const unsigned*>(rowp) + p.ncols * 4 - 1);
        if (last && p.bc0 == SB200_WRAP) wr = __ldg(reinterpret_cast<const unsigned*>(rowp));
    }
    if (!IS_BOOL) { w0 = nz_bytes(w0); w1 = nz_bytes(w1); w2 = nz_bytes(w2); w3 = nz_bytes(w3); wl = nz_bytes(wl); wr = nz_bytes(wr); }
    const unsigned sl = __shfl_up_sync(0xffffffffu, w3, 1);
    const unsigned sr = __shfl_down_sync(0xffffffffu, w0, 1);
    if (lane != 0) wl = sl;
    if (lane != 31 && !last) wr = sr;
    if (first && p.bc0 != SB200_WRAP) wl = (p.bc0 == SB200_REFLECT) ? (w0 << 16) & 0xFF000000u : p.pad01 << 24;
    if (last && p.bc0 != SB200_WRAP) wr = (p.bc0 == SB200_REFLECT) ? (w3 >> 16) & 0xFFu : p.pad01;
    out.c[0] = w0; out.c[1] = w1; out.c[2] = w2; out.c[3] = w3;
    out.h[0] = __funnelshift_l(wl, w0, 8) + w0 + __funnelshift_r(w0, w1, 8);
    out.h[1] = __funnelshift_l(w0, w1, 8) + w1 + __funnelshift_r(w1, w2, 8);
    out.h[2] = __funnelshift_l(w1, w2, 8) + w2 + __funnelshift_r(w2, w3, 8);
    out.h[3] = __funnelshift_l(w2, w3, 8) + w3 + __funnelshift_r(w3, wr, 8);
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

template <bool IS_BOOL, bool CONWAY>
__global__ void __launch_bounds__(256) life_swar_kernel(LifeParams p) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int colgroups = (p.ncols + 31) >> 5;
    const long long ntasks = (long long)colgroups * p.nbands;
    for (long long task = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); task < ntasks;
         task += (long long)gridDim.x * warps_per_block) {
        const int cg = (int)(task % colgroups), band = (int)(task / colgroups);
        const int col = cg * 32 + lane;
        const bool active = col < p.ncols;
        const int y0 = p.y_lo + band * p.RY;
        const int y1 = min(y0 + p.RY, p.y_hi);
        Row prev, cur, next;
        load_row<IS_BOOL>(p, y0 - 1, col, active, lane, prev);
        load_row<IS_BOOL>(p, y0, col, active, lane, cur);
#pragma unroll 2
        for (int y = y0; y < y1; y++) {
            load_row<IS_BOOL>(p, y + 1, col, active, lane, next);
            uint4 o;
            o.x = rule<CONWAY>(prev.h[0] + cur.h[0] + next.h[0], cur.c[0], p.born, p.survive);
            o.y = rule<CONWAY>(prev.h[1] + cur.h[1] + next.h[1], cur.c[1], p.born, p.survive);
            o.z = rule<CONWAY>(prev.h[2] + cur.h[2] + next.h[2], cur.c[2], p.born, p.survive);
            o.w = rule<CONWAY>(prev.h[3] + cur.h[3] + next.h[3], cur.c[3], p.born, p.survive);
            if (active)
                *(reinterpret_cast<uint4*>(p.dst + (long long)(y + p.doff1) * p.dpitch) + col) = o;
            prev = cur;
            cur = next;
        }
    }
}

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
    p.soff1 = d.src_off[1]; p.doff1 = d.dst_off[1];
    p.bc0 = d.boundary[0]; p.bc1 = d.boundary[1];
    p.pad01 = (d.padval_bits & 0xFF) != 0;
    p.y_lo = (int)pl.dd.lo[1]; p.y_hi = (int)(pl.dd.lo[1] + pl.dd.n[1]);
    p.born = d.born_mask; p.survive = d.survive_mask;
    const bool conway = d.born_mask == (1u << 3) && d.survive_mask == ((1u << 2) | (1u << 3));
    const bool is_bool = d.eltype == SB200_BOOL;

    // Band height: trade halo re-reads (2 extra rows per band) against the tail of the last wave.
    const int colgroups = (p.ncols + 31) / 32;
    const int rows = p.y_hi - p.y_lo;
    const long long resident = (long long)num_sms() * 48;  // warps in flight at ~40 registers/thread
    int best_ry = 16;
    double best_cost = 1e300;
    for (int ry : {64, 48, 32, 24, 16, 8}) {
        const long long tasks = (long long)colgroups * ((rows + ry - 1) / ry);
        const long long waves = (tasks + resident - 1) / resident;
        const double cost = (double)waves * (ry + 2);
        if (cost < best_cost) { best_cost = cost; best_ry = ry; }
    }
    p.RY = best_ry;
    p.nbands = (rows + p.RY - 1) / p.RY;
    const long long ntasks = (long long)colgroups * p.nbands;
    const int wpb = 8;
    long long blocks = (ntasks + wpb - 1) / wpb;
    blocks = std::min<long long>(blocks, (long long)num_sms() * 8);
    if (is_bool) {
        if (conway) life_swar_kernel<true, true><<<(unsigned)blocks, wpb * 32, 0, st>>>(p);
        else life_swar_kernel<true, false><<<(unsigned)blocks, wpb * 32, 0, st>>>(p);
    } else {
        if (conway) life_swar_kernel<false, true><<<(unsigned)blocks, wpb * 32, 0, st>>>(p);
        else life_swar_kernel<false, false><<<(unsigned)blocks, wpb * 32, 0, st>>>(p);
    }
    SB_LAUNCH_CHECK();
    set_kernel_name(conway ? "life_swar_kernel<conway>" : "life_swar_kernel<table>");
    return SB200_OK;
}

}  // namespace sb
