// life_bit.cuh — the bit-sliced B3/S23 kernel (G generations per launch, cells packed one bit each inside the kernel) and its launcher.
// Instantiated in life_bit_u8.cu (byte in, byte out), life_bit_pk.cu (packed source and / or dest); dispatched from life.cu.
#pragma once
#include <algorithm>
#include <cstdlib>
#include "life_params.cuh"

namespace sb {

// Bit-sliced Conway kernel: G generations per launch with the cells of a row packed ONE BIT each inside the kernel.
// Memory keeps the reference's one byte per cell; a lane loads 32 cells (two 128-bit shared-memory loads), packs them
// into one 32-bit word with four multiplies (w * 0x10204080 gathers the four 0/1 bytes of a word into its top nibble),
// and every logic instruction then advances 32 cells: the horizontal 3-sum of a row is two LOP3 (xor3, majority) on the
// word and its two one-bit shifts, the 3-row total a 4-bit carry-save sum (8 LOP3/LOP), B3/S23 four more — about 20
// instructions per 32 cells and generation against ~18 per FOUR cells in the byte-SWAR kernels. The shifted-in edge bits
// of an intermediate generation come from the adjacent lanes by warp shuffle, so lanes G-1 .. 32-G of a warp own final
// cells (warps overlap by 2(G-1) lanes). Same TMA ring, same boundary support as life_tma2_kernel; B3/S23 only.
constexpr int LB_WARPS = 6;
constexpr int LB_CH = 3;
constexpr int LB_STAGES = 4;
constexpr int LB_ROWB = 6144;
constexpr int LB_SMEM = 128 + LB_STAGES * LB_CH * LB_ROWB;
template <int G> struct LbCfg {
    static constexpr int HLN = SB200_LB_ONE_HALO_LANE ? 1 : G - 1;   // halo lanes per side of a warp
    static constexpr int VALID = 32 - 2 * HLN;               // lanes of a warp that own final cells
    static constexpr int WO = VALID * 32;                     // final cells per warp row
    static constexpr int CAP = LB_WARPS * WO;                 // final cells per strip row (multiple of 128)
    static constexpr int HL = HLN * 32 + 16;                  // halo bytes per side of a shared-memory row
    static constexpr int D0 = (128 - HL % 128) % 128;         // data start in a row: global x0 - HL is D0 mod 128
    static_assert(D0 + 2 * HL + CAP <= LB_ROWB, "row does not fit");
};

struct BRow { unsigned c, s0, s1; };  // 32 cells, and their horizontal 3-sums (left + centre + right) as two bit planes

__device__ __forceinline__ unsigned lop3_xor3(unsigned a, unsigned b, unsigned c) {
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ unsigned lop3_maj(unsigned a, unsigned b, unsigned c) {
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// `prev` / `next` = the words of the 32 cells to the left / right (only their top / bottom bit is used): the one-bit shifts with
// the neighbour's edge bit shifted in are single funnel shifts (SHF.L.W / SHF.R.W on two registers).
// SB200_LB_IMAD_SHIFT=1 computes them on the FMA pipe instead (the kernel is bound by the ALU pipe — LOP3 / SHF issue every other
// cycle per scheduler — while the FMA pipe idles at 7 %): L = w * 2 + hi32(prev * 2), R = hi32(w * 2^31) + next * 2^31, four IMAD /
// IMAD.HI with the multipliers read from constant memory (an immediate power of two is strength-reduced back to SHF / LEA).
#ifndef SB200_LB_IMAD_SHIFT
#define SB200_LB_IMAD_SHIFT 0
#endif
static __constant__ unsigned lb_mul_consts[7] = {2u, 0x80000000u, 16u, 1u << 28, 1u << 24, 1u << 20, 1u << 16};
struct LbMul { unsigned two, half, sixteen, un[4]; };
#ifndef SB200_LB_IMAD_UNPACK
#define SB200_LB_IMAD_UNPACK 0   // nibble k of a word as hi32((bits << (28 - 4k)) * 16): IMAD + IMAD.HI instead of SHF + LOP3
#endif
#ifndef SB200_LB_IMAD_PACK
#define SB200_LB_IMAD_PACK 0     // acc = acc * 16 + hi32(product * 16): IMAD.HI + IMAD instead of SHF.L.W
#endif
__device__ __forceinline__ BRow brow(unsigned w, unsigned prev, unsigned next, const LbMul& m) {
#if SB200_LB_IMAD_SHIFT
    unsigned e, f, L, R;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(e) : "r"(prev), "r"(m.two));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(L) : "r"(w), "r"(m.two), "r"(e));
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(f) : "r"(next), "r"(m.half));
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(R) : "r"(w), "r"(m.half), "r"(f));
#else
    const unsigned L = __funnelshift_l(prev, w, 1), R = __funnelshift_r(w, next, 1);
#endif
    return BRow{w, lop3_xor3(L, w, R), lop3_maj(L, w, R)};
}
// B3/S23 from the rows above, at and below: T = 3x3 total including the centre; alive' = (T == 3) | (centre & T == 4)
// T = t0 + 2 (u1 + c0) + 4 c1 with t0 / c0 = sum / carry of the three low bit planes and u1 / c1 of the three high ones, so
//   T == 3  <=>  t0 & (u1 ^ c0) & ~c1          (bit 1 set without a carry into bit 2, no c1)
//   T == 4  <=>  ~t0 & ~(u1 ^ c0) & (u1 ^ c1)  (u1 == c0: bit 1 clear, carry = u1; exactly one of carry and c1)
// so alive' = t0 ? (u1 ^ c0) & ~c1 : ~(u1 ^ c0) & (u1 ^ c1) & centre: two tables over (u1, c0, c1), the centre AND and one select —
// EIGHT logic ops for the 3-row total and the rule (round 1's ripple form took twelve, round 2's first table form nine): every
// instruction counts, the packed launches are bound by instruction issue. tests/test_kernel_models.py checks the identity over
// every input combination with the tables read out of this source.
template <int IMM> __device__ __forceinline__ unsigned lop3_imm(unsigned a, unsigned b, unsigned c) {
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(IMM));
    return r;
}
__device__ __forceinline__ unsigned conway_bits(const BRow& a, const BRow& b, const BRow& n) {
    const unsigned t0 = lop3_xor3(a.s0, b.s0, n.s0), c0 = lop3_maj(a.s0, b.s0, n.s0);
    const unsigned u1 = lop3_xor3(a.s1, b.s1, n.s1), c1 = lop3_maj(a.s1, b.s1, n.s1);
    const unsigned a3 = lop3_imm<0x14>(u1, c0, c1);            // (u1 ^ c0) & ~c1: T == 3 once t0 is set
    const unsigned b4 = lop3_imm<0x42>(u1, c0, c1) & b.c;      // ~(u1 ^ c0) & (u1 ^ c1): T == 4 once t0 is clear; and the centre is alive
    return lop3_imm<0xCA>(t0, a3, b4);                         // t0 ? a3 : b4
}
// The product w * 0x10204080 holds the four 0/1 bytes of w in its top nibble; SHF.L.W (acc:product) << 4 appends exactly that
// nibble to the accumulator (words 7 .. 0, so that cell 0 ends in bit 0): two instructions per word, no mask.
template <bool CELLS01> __device__ __forceinline__ unsigned pack32(const uint4& lo, const uint4& hi, const LbMul& m) {
    unsigned w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    unsigned acc = 0;
#pragma unroll
    for (int k = 7; k >= 0; k--) {
        const unsigned v = CELLS01 ? w[k] : nz_bytes(w[k]);
#if SB200_LB_IMAD_PACK
        unsigned nib;
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(nib) : "r"(v * 0x10204080u), "r"(m.sixteen));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(acc) : "r"(acc), "r"(m.sixteen), "r"(nib));
#else
        acc = __funnelshift_l(v * 0x10204080u, acc, 4);
#endif
    }
    return acc;
}
__device__ __forceinline__ unsigned unpack4(unsigned bits, int k, const LbMul& m) {  // cells 4k .. 4k+3 as four 0/1 bytes
#if SB200_LB_IMAD_UNPACK
    unsigned nib;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(nib) : "r"(bits * m.un[k]), "r"(m.sixteen));   // (bits << (28 - 4k)) >> 28
    return (nib * 0x00204081u) & 0x01010101u;
#else
    return (((bits >> (4 * k)) & 0xFu) * 0x00204081u) & 0x01010101u;
#endif
}

// Storage formats of a launch. IN: LB_U8 = one byte per cell, any value (alive = non-zero); LB_U8_01 = bytes known to be 0 / 1;
// LB_BITS = one BIT per cell (SB200_FLAG_SRC_BITS: row r of the parent starts at byte r * pitch, cell c is bit c % 32 of its 32-bit
// word c / 32 — the kernel's own register layout, so a lane loads its 32 cells with one LDS.32 and the pack disappears). OUT_BITS
// stores the lane's word as it is (SB200_FLAG_DST_BITS) instead of unpacking it into 32 bytes. sb200_iterate and the slab plans keep
// the state packed between the first and the last launch of a run: half of the instructions of a byte-to-byte launch are pack and
// unpack, and the launches in between move W * H / 8 bytes each way instead of W * H.
enum { LB_U8 = 0, LB_U8_01 = 1, LB_BITS = 2 };
constexpr int LB_HLB = 16;   // packed rows: halo bytes per side of a strip (128 cells: bulk copies move multiples of 16 bytes)
constexpr int LB_ROWB_PK = 768;   // shared-memory row of a packed source: 16 + 5760 / 8 + 16 = 752 bytes
template <int IN> constexpr int lb_rowb() { return IN == LB_BITS ? LB_ROWB_PK : LB_ROWB; }
template <int IN> constexpr int lb_smem() { return 128 + LB_STAGES * LB_CH * lb_rowb<IN>(); }
// Resident CTAs per SM the kernel is compiled for. Byte launches: 3 (74 KB of ring each, <= 96 registers). Packed -> packed launches
// have a 9 KB ring, so the registers decide: 9 G state registers + temporaries fit 4 CTAs (<= 72 registers) up to G = 6 and 5 CTAs
// (<= 56) up to G = 4; more resident warps hide the latency of the dependent LOP3 chains. SB200_LB_PK_CTAS=0 keeps 3 everywhere.
#ifndef SB200_LB_PRED_EMIT
#define SB200_LB_PRED_EMIT 0   // packed dest: the output row as a predicated store instead of a branch around the store. Measured r02ar:
                               // slower (six generations 72.4 -> 76.5 us: the last generation is computed for rows that store nothing). Off.
#endif
#ifndef SB200_LB_PK_CTAS
#define SB200_LB_PK_CTAS 1
#endif
template <int G, int IN, bool OUT_BITS> constexpr int lb_min_ctas() {
    return (SB200_LB_PK_CTAS && IN == LB_BITS && OUT_BITS) ? (G <= 4 ? 5 : G <= 6 + (SB200_LB_PK_CTAS > 1 ? SB200_LB_PK_CTAS - 1 : 0) ? 4 : 3) : 3;
}

template <int G, int IN, bool OUT_BITS>
__global__ void __launch_bounds__((LB_WARPS + 1) * 32, lb_min_ctas<G, IN, OUT_BITS>()) life_bit_kernel(const LifeTmaParams q) {
    using C = LbCfg<G>;
    constexpr bool CELLS01 = IN == LB_U8_01;
    constexpr int ROWB = lb_rowb<IN>();
    static_assert((IN != LB_BITS && !OUT_BITS) || SB200_LB_ONE_HALO_LANE, "the packed formats use the one-halo-lane layout");
    extern __shared__ __align__(128) uint8_t smem[];
    const LifeParams& p = q.lp;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + LB_STAGES;
    uint8_t* ring = smem + 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < LB_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], LB_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    const int ntasks = q.nstrips * q.nruns;
    const LbMul mul{lb_mul_consts[0], lb_mul_consts[1], lb_mul_consts[2], {lb_mul_consts[3], lb_mul_consts[4], lb_mul_consts[5], lb_mul_consts[6]}};
    unsigned k = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int strip = task % q.nstrips, run = task / q.nstrips;
        const int x0 = strip * q.outb;
        const int wout = min(q.outb, p.W - x0);
        const int y0 = p.y_lo + (int)((long long)p.rows * run / q.nruns);
        const int y1 = p.y_lo + (int)((long long)p.rows * (run + 1) / q.nruns);
        const int nsrc = y1 - y0 + 2 * G;  // source rows y0-G .. y1+G-1
        const int nchunks = (nsrc + LB_CH - 1) / LB_CH;
        if (warp == LB_WARPS) {
            // ---------------- producer: shared-memory row byte b <-> global column x0 - HL + b (mod W) ----------------
            // (packed source: byte b <-> byte x0 / 8 - LB_HLB + b of the packed row, mod W / 8)
            if (lane == 0) {
                constexpr int HLx = IN == LB_BITS ? LB_HLB : C::HL;          // halo bytes per side
                constexpr int D0x = IN == LB_BITS ? 0 : C::D0;
                const int xb = IN == LB_BITS ? x0 >> 3 : x0, wb = IN == LB_BITS ? wout >> 3 : wout, Wb = IN == LB_BITS ? p.W >> 3 : p.W;
                const int lin = xb >= HLx ? HLx : 0;
                const int rin = min(HLx, Wb - (xb + wb));
                const unsigned mlen = lin + wb + rin;
                const unsigned rowbytes = 2 * HLx + wb;
                for (int c = 0; c < nchunks; c++, k++) {
                    const int slot = k % LB_STAGES;
                    mbar_wait_producer(&empty[slot], ((k / LB_STAGES) & 1) ^ 1);
                    uint8_t* sbase = ring + slot * (LB_CH * ROWB);
                    const int nrows = min(LB_CH, nsrc - c * LB_CH);
                    mbar_arrive_expect_tx(&full[slot], nrows * rowbytes);
                    for (int j = 0; j < nrows; j++) {
                        const uint8_t* g = p.src + life_map_row(p, y0 - G + c * LB_CH + j) * p.spitch;
                        uint8_t* srow = sbase + j * ROWB + D0x;
                        bulk_g2s(srow + HLx - lin, g + xb - lin, mlen, &full[slot]);
                        if (!lin) bulk_g2s(srow, g + Wb - HLx, HLx, &full[slot]);                           // wrapped left halo
                        if (rin < HLx) bulk_g2s(srow + HLx + wb + rin, g, HLx - rin, &full[slot]);          // wrapped right halo
                    }
                }
            } else {
                k += nchunks;
            }
            continue;
        }
        // ---------------- consumers ----------------
        if (warp * C::WO >= wout) {   // no final cell in this warp: keep the ring protocol only
            for (int c = 0; c < nchunks; c++, k++) {
                const int slot = k % LB_STAGES;
                mbar_wait(&full[slot], (k / LB_STAGES) & 1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
            }
            continue;
        }
        const int cell0 = warp * C::WO + (lane - C::HLN) * 32;   // first final cell of this lane inside the strip
        const bool active = lane >= C::HLN && lane <= 31 - C::HLN && cell0 < wout;
        const unsigned act_mask = __ballot_sync(0xffffffffu, active);
        const bool st_ok[2] = {((act_mask >> (lane >> 1)) & 1u) != 0, ((act_mask >> (16 + (lane >> 1))) & 1u) != 0};   // my two store slots
        // destination of the cells of lane 0 (a halo lane: never stored) in output row y0; packed dest: of this lane's own word
        uint8_t* __restrict__ wp = OUT_BITS ? p.dst + (long long)(y0 + p.doff1) * p.dpitch + ((x0 + warp * C::WO) >> 3) + (lane - C::HLN) * 4
                                            : p.dst + (long long)(y0 + p.doff1) * p.dpitch + x0 + warp * C::WO - C::HLN * 32;
        long long dpitch = p.dpitch;
        asm volatile("" : "+l"(dpitch));   // keep the pitch in registers: ptxas otherwise reloads it from the constant bank every row (LDCU + long-scoreboard stall)
        const int soff = IN == LB_BITS ? LB_HLB + ((warp * C::WO) >> 3) + (lane - C::HLN) * 4 : C::D0 + C::HL + cell0;
        BRow lv[G][3];
#pragma unroll
        for (int a = 0; a < G; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) lv[a][b] = BRow{0, 0, 0};
        for (int c = 0; c < nchunks; c++, k++) {
            const int slot = k % LB_STAGES;
            mbar_wait(&full[slot], (k / LB_STAGES) & 1);
            const uint8_t* sbase = ring + slot * (LB_CH * ROWB);
#pragma unroll
            for (int J = 0; J < LB_CH; J++) {      // stream index i = c * LB_CH + J, i % 3 == J
                const int i = c * LB_CH + J;
                {   // level 0: pack the source row (or load it packed)
                    const uint8_t* t = sbase + J * ROWB + soff;
                    unsigned w;
                    if constexpr (IN == LB_BITS) {
                        w = *reinterpret_cast<const unsigned*>(t);
                    } else {
                        const uint4 lo = *reinterpret_cast<const uint4*>(t), hi = *reinterpret_cast<const uint4*>(t + 16);
                        w = pack32<CELLS01>(lo, hi, mul);
                    }
                    // edge bits from the adjacent lanes' packed words. With one halo lane per side the outer edge bit of an end lane
                    // may be anything (the shuffle hands lane 0 / 31 its own word, as at every later level): a wrong bit moves one
                    // cell per generation and stays inside the end lane, which owns no final cell. The G - 1 halo-lane layout of
                    // round 1 reads the real halo byte.
                    unsigned prev = __shfl_up_sync(0xffffffffu, w, 1), next = __shfl_down_sync(0xffffffffu, w, 1);
#if !SB200_LB_ONE_HALO_LANE
                    if (lane == 0) prev = t[-1] != 0 ? 0x80000000u : 0u;
                    if (lane == 31) next = t[32] != 0;
#endif
                    lv[0][J] = brow(w, prev, next, mul);
                }
#pragma unroll
                for (int g = 1; g < G; g++) {  // generation g of row i - g from generation g-1 of rows i-g-1, i-g, i-g+1
                    const unsigned x = conway_bits(lv[g - 1][(J - g - 1 + 9) % 3], lv[g - 1][(J - g + 9) % 3], lv[g - 1][(J - g + 1 + 9) % 3]);
                    const unsigned prev = __shfl_up_sync(0xffffffffu, x, 1), next = __shfl_down_sync(0xffffffffu, x, 1);
                    lv[g][(J - g + 9) % 3] = brow(x, prev, next, mul);
                }
                if constexpr (OUT_BITS && SB200_LB_PRED_EMIT) {
                    // generation G of row i - G = the output row y0 + i - 2G, as ONE predicated store: no branch in the row body, so
                    // ptxas keeps the three rows of an iteration in one basic block (the branch around a store block costs a divergence
                    // check — BRA.DIV — in front of the next row's shuffles)
                    const bool emit = i >= 2 * G && i < nsrc;
                    const unsigned y = conway_bits(lv[G - 1][(J - G - 1 + 9) % 3], lv[G - 1][(J - G + 9) % 3], lv[G - 1][(J - G + 1 + 9) % 3]);
                    if (emit && active) *reinterpret_cast<unsigned*>(wp) = y;   // 30 consecutive words per warp row
                    wp += emit ? dpitch : 0;
                } else if (i >= 2 * G && i < nsrc) {   // generation G of row i - G: the output row y0 + i - 2G
                    const unsigned y = conway_bits(lv[G - 1][(J - G - 1 + 9) % 3], lv[G - 1][(J - G + 9) % 3], lv[G - 1][(J - G + 1 + 9) % 3]);
                    if constexpr (OUT_BITS) {
                        if (active) *reinterpret_cast<unsigned*>(wp) = y;   // 30 consecutive words per warp row
                    } else {
                    // two fully coalesced 512-byte stores per warp row: lane l writes 16 bytes at 16 l of the first / second
                    // half of the warp's 1 KiB, i.e. half (l & 1) of the cells of lane l/2 resp. 16 + l/2
#pragma unroll
                    for (int hb = 0; hb < 2; hb++) {
                        const int sl = hb * 16 + (lane >> 1);
                        const unsigned h16 = __shfl_sync(0xffffffffu, y, sl) >> (16 * (lane & 1));
                        const uint4 cells = make_uint4(unpack4(h16, 0, mul), unpack4(h16, 1, mul), unpack4(h16, 2, mul), unpack4(h16, 3, mul));
                        if (st_ok[hb]) *reinterpret_cast<uint4*>(wp + hb * 512 + lane * 16) = cells;   // a predicated store, no branch
                    }
                    }
                    wp += dpitch;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
    }
}

template <int G, int IN, bool OUT_BITS> static int launch_bit(const LifeParams& p, cudaStream_t st) {
    using C = LbCfg<G>;
    static thread_local int cfg_dev = -1, ctas_per_sm = 0;
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (dev != cfg_dev) {
        SB_CUDA(cudaFuncSetAttribute(life_bit_kernel<G, IN, OUT_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, lb_smem<IN>()));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, life_bit_kernel<G, IN, OUT_BITS>, (LB_WARPS + 1) * 32, lb_smem<IN>()) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        ctas_per_sm = per_sm;
        cfg_dev = dev;
    }
    LifeTmaParams q;
    q.lp = p;
    q.nstrips = (p.W + C::CAP - 1) / C::CAP;
    q.outb = std::min(C::CAP, ((p.W + q.nstrips - 1) / q.nstrips + 127) / 128 * 128);  // equal strips
    q.nstrips = (p.W + q.outb - 1) / q.outb;
    const long long ctas = (long long)ctas_per_sm * num_sms();
    // Runs per strip: strips x runs fills whole waves of the resident CTAs (t tasks per CTA); every run re-reads 2 G source
    // rows, so pick the t that minimises waves x (rows per run + 2 G). Measured r02a (16384^2, three strips, 444 CTAs): one task
    // per CTA 9843 Gcell-updates/s, a clipped second wave (768 tasks) 8283. SB200_LB_TASKS overrides t for A/B runs.
    static const int tpc_env = getenv("SB200_LB_TASKS") ? atoi(getenv("SB200_LB_TASKS")) : 0;
    // small grids: a run may be as short as G rows (it re-reads 2 G more) when that is what fills the SMs — the cost below decides;
    // SB200_LB_MIN_ROWS_X (default 1) x G rows is the floor (4 = the round-2 behaviour)
    static const int min_rows_x = getenv("SB200_LB_MIN_ROWS_X") ? std::max(1, atoi(getenv("SB200_LB_MIN_ROWS_X"))) : 1;
    const int lb_min_rows = min_rows_x * G;
    long long nruns = 1;
    double best_cost = 1e300;
    for (int t = (tpc_env > 0 ? tpc_env : 1); t <= (tpc_env > 0 ? tpc_env : 4); t++) {
        long long r = std::max<long long>(1, t * ctas / q.nstrips);
        r = std::min<long long>(r, std::max(1, p.rows / lb_min_rows));   // at least lb_min_rows rows per run
        const long long waves = (q.nstrips * r + ctas - 1) / ctas;
        const double cost = (double)waves * ((double)p.rows / (double)r + 2.0 * G);
        if (cost < best_cost * 0.999) { best_cost = cost; nruns = r; }
    }
    q.nruns = (int)nruns;
    const long long grid = std::min<long long>(ctas, (long long)q.nstrips * q.nruns);
    life_bit_kernel<G, IN, OUT_BITS><<<(unsigned)grid, (LB_WARPS + 1) * 32, lb_smem<IN>(), st>>>(q);
    return SB200_OK;
}

// One switch per translation unit: `gens` generations with the formats <IN, OUT_BITS>.
template <int IN, bool OUT_BITS> static int launch_bit_gens(int gens, const LifeParams& p, cudaStream_t st) {
    if constexpr (IN == LB_BITS) {   // a packed source also runs single generations (the remainder of a slab plan's cycle)
        if (gens == 1) return launch_bit<1, IN, OUT_BITS>(p, st);
    }
    switch (gens) {
        case 2: return launch_bit<2, IN, OUT_BITS>(p, st);
        case 3: return launch_bit<3, IN, OUT_BITS>(p, st);
        case 4: return launch_bit<4, IN, OUT_BITS>(p, st);
#if SB200_LB_ONE_HALO_LANE
        case 5: return launch_bit<5, IN, OUT_BITS>(p, st);
        case 6: return launch_bit<6, IN, OUT_BITS>(p, st);
        case 7: return launch_bit<7, IN, OUT_BITS>(p, st);
        case 8: return launch_bit<8, IN, OUT_BITS>(p, st);
#endif
        default: break;
    }
    set_error("life_bit_kernel: %d generations per launch are not built", gens);
    return SB200_EUNSUPPORTED;
}

}  // namespace sb
