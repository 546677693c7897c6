// life_params.cuh — parameter blocks and small helpers shared by the Life kernels (life.cu, life_bit.cuh)
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace sb {

// ONE halo lane per side of a warp instead of G - 1 (default since r02a; -DSB200_LB_ONE_HALO_LANE=0 restores the round-1 layout).
// A wrong edge bit enters an end lane at one cell per generation, so after G <= 32 generations only the end lanes themselves
// hold wrong cells (tools/model_life_bit_lanes.py emulates the scheme: lanes 1 .. 30 are exact for G = 2 .. 31). With G = 4 that
// is 30 instead of 26 useful lanes per warp and smaller halos, and it is what makes 8 generations per launch worthwhile (18
// useful lanes with G - 1 halo lanes). Measured r02a, 16384^2: 8508 (G - 1 halo lanes, 4 generations) -> 9843 (one halo lane,
// one task per CTA) -> 10773 Gcell-updates/s (8 generations per launch); every Life test bit-identical to the CPU restatement.
#ifndef SB200_LB_ONE_HALO_LANE
#define SB200_LB_ONE_HALO_LANE 1
#endif

struct LifeParams {
    const uint8_t* src;
    uint8_t* dst;
    long long spitch, dpitch;  // bytes per row of the parents
    int W, H;                  // logical size (axis 0 = W contiguous)
    int ncols;                 // W / 16
    int colgroups;             // ceil(ncols / 32)
    int soff1, doff1;          // ring / ghost rows on axis 1
    uint8_t* mirror;           // fused ghost push: rows [m_lo, m_hi) are also stored here (row m_lo first), or null
    int m_lo, m_hi;
    int bc0, bc1;              // boundary per axis
    unsigned pad01;            // Remove: (padval != 0)
    int y_lo, rows;            // output rows [y_lo, y_lo + rows)
    int nruns;                 // the rows are split into nruns equal runs; a warp owns one (column group, run)
    unsigned born, survive;
};

// 0/1 per byte: byte != 0
__device__ __forceinline__ unsigned nz_bytes(unsigned w) {
    return ((((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) >> 7) & 0x01010101u;
}
// 0/1 per byte: byte == 0, valid for bytes <= 0x10
__device__ __forceinline__ unsigned eqz_small(unsigned x) { return ((0x10101010u - x) >> 4) & 0x01010101u; }

struct LifeTmaParams {
    LifeParams lp;
    int nstrips, nruns;
    int outb;   // life_tma2_kernel: final cells per strip row (<= LT2_OUTB, a multiple of 128: equal strips)
};

// source row r (logical) -> parent row, or -1 for a Remove pad row
__device__ __forceinline__ long long life_map_row(const LifeParams& p, int r) {
    if (p.soff1 > 0) return (long long)r + p.soff1;
    if (r >= 0 && r < p.H) return r;
    if (p.bc1 == SB200_WRAP) return r < 0 ? r + p.H : r - p.H;
    if (p.bc1 == SB200_REFLECT) return r < 0 ? -r : 2 * (p.H - 1) - r;
    return -1;
}

}  // namespace sb
