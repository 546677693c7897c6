// common.cuh — shared host/device definitions of libstencils_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/stencils_b200.h"

namespace sb {

// ---------------------------------------------------------------- errors / diagnostics (thread-local)
void set_error(const char* fmt, ...);
void set_kernel_name(const char* name);
void count_launch(int n = 1);

#define SB_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            sb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SB200_ECUDA;                                                               \
        }                                                                                     \
    } while (0)

#define SB_LAUNCH_CHECK()                                                                     \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            sb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SB200_ECUDA;                                                               \
        }                                                                                     \
        sb::count_launch();                                                                   \
    } while (0)

// ---------------------------------------------------------------- element types
template <int E> struct ElType;
template <> struct ElType<SB200_BOOL> { using type = uint8_t; };
template <> struct ElType<SB200_U8> { using type = uint8_t; };
template <> struct ElType<SB200_I32> { using type = int32_t; };
template <> struct ElType<SB200_I64> { using type = int64_t; };
template <> struct ElType<SB200_F32> { using type = float; };
template <> struct ElType<SB200_F64> { using type = double; };

inline size_t elsize(int e) {
    switch (e) {
    case SB200_BOOL: case SB200_U8: return 1;
    case SB200_I32: case SB200_F32: return 4;
    case SB200_I64: case SB200_F64: return 8;
    default: return 0;
    }
}

// Device-side view of a sweep: strides in elements, everything 0-based.
struct DevDesc {
    int ndim, L, R, reducer;
    long long size[3];   // logical size
    long long sstr[3];   // source parent strides
    long long dstr[3];   // dest parent strides
    long long sext[3];   // source parent extents
    int soff[3], doff[3], bc[3];
    long long lo[3], n[3];  // output region origin and extents
    unsigned long long padbits;
    unsigned born, survive;
    double alpha;
    const int* offs;     // device [L][3]
    const void* weights; // device [L] of eltype
    int scatter_op, scatter_rule, flags;
};

// A cached, validated sweep: device copies of the tables plus the dispatch decision.
struct Plan {
    sb200_desc d;           // host copy (pointers replaced by owned tables)
    DevDesc dd;
    int* offs_dev = nullptr;
    void* weights_dev = nullptr;
    int* scatter_order_dev = nullptr;  // [2R+1][L] static fold order (scatter only)
    int shape_tag = -1;     // recognised named shape (sb200_shape) or -1
    int shape_ndim = 0;     // dimensionality of the recognised shape
    bool uniform_bc = true;
    unsigned long long stamp = 0;   // last lookup (plan cache eviction)
    std::string key;
};

// ---------------------------------------------------------------- device arithmetic with Julia semantics
#ifdef __CUDACC__
// Additions / multiplications that ptxas may never contract into an FMA (Julia does not contract).
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
// Integers wrap (Julia native integer arithmetic).
__device__ __forceinline__ uint8_t add_rn(uint8_t a, uint8_t b) { return (uint8_t)(a + b); }
__device__ __forceinline__ int32_t add_rn(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
__device__ __forceinline__ int64_t add_rn(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
__device__ __forceinline__ uint8_t mul_rn(uint8_t a, uint8_t b) { return (uint8_t)(a * b); }
__device__ __forceinline__ int32_t mul_rn(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
__device__ __forceinline__ int64_t mul_rn(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }

// Julia max/min (Base): NaN-propagating, -0.0 < +0.0. max.NaN.f32 is one FMNMX on sm_100a and
// orders the zeros the IEEE-754-2019 way (+0 > -0).
__device__ __forceinline__ float jl_max(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float jl_min(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
// Three-input forms (sm_100 FMNMX3, same issue rate as FMNMX: two comparisons per slot). Maximum / minimum with these
// semantics (NaN wins, +0 > -0) do not depend on the order of the operands.
__device__ __forceinline__ float jl_max3(float a, float b, float c) {
    float r;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float jl_min3(float a, float b, float c) {
    float r;
    asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ double jl_max(double a, double b) {
    if (a != a || b != b) return __longlong_as_double(0x7ff8000000000000LL);
    if (a > b) return a;
    if (a < b) return b;
    return (__double_as_longlong(a) < 0) ? b : a;  // equal: prefer +0.0
}
__device__ __forceinline__ double jl_min(double a, double b) {
    if (a != a || b != b) return __longlong_as_double(0x7ff8000000000000LL);
    if (a < b) return a;
    if (a > b) return b;
    return (__double_as_longlong(a) < 0) ? a : b;  // equal: prefer -0.0
}
__device__ __forceinline__ double jl_max3(double a, double b, double c) { return jl_max(jl_max(a, b), c); }
__device__ __forceinline__ double jl_min3(double a, double b, double c) { return jl_min(jl_min(a, b), c); }
__device__ __forceinline__ uint8_t jl_max(uint8_t a, uint8_t b) { return a > b ? a : b; }
__device__ __forceinline__ uint8_t jl_min(uint8_t a, uint8_t b) { return a < b ? a : b; }
__device__ __forceinline__ int32_t jl_max(int32_t a, int32_t b) { return a > b ? a : b; }
__device__ __forceinline__ int32_t jl_min(int32_t a, int32_t b) { return a < b ? a : b; }
__device__ __forceinline__ int64_t jl_max(int64_t a, int64_t b) { return a > b ? a : b; }
__device__ __forceinline__ int64_t jl_min(int64_t a, int64_t b) { return a < b ? a : b; }

// bounded_index (src/array.jl:146-179), 0-based; -1 = out of bounds under Remove.
__device__ __forceinline__ long long bounded(long long j, long long s, int bc) {
    if (j >= 0 && j < s) return j;
    if (bc == SB200_WRAP) return j < 0 ? j + s : j - s;
    if (bc == SB200_REFLECT) return j < 0 ? -j : 2 * (s - 1) - j;
    return -1;
}
#endif  // __CUDACC__

// Fused ghost-plane push requested by the current sweep (sb200_desc.mirror_*): set by do_gather around the dispatch.
// A kernel that stores the planes itself sets `honoured`; otherwise do_gather copies them after the sweep.
struct MirrorReq {
    void* ptr = nullptr;       // plane `lo` lands here; plane pitch = dest parent's
    long long lo = 0, hi = 0;  // logical planes [lo, hi) of the last axis
    bool honoured = false;
};
extern thread_local MirrorReq g_mirror;

// ---------------------------------------------------------------- kernel families (one .cu each)
int launch_generic_gather(const Plan& pl, const void* src, void* dst, cudaStream_t st);
int launch_update_halo(const Plan& pl, void* parent, cudaStream_t st);
int launch_generic_scatter(const Plan& pl, const void* src, void* dst, cudaStream_t st);
int launch_generic_scatter_rect(const Plan& pl, const void* src, void* dst, cudaStream_t st, long long lo0, long long hi0,
                                long long lo1, long long hi1);
// Specialised kernels return SB200_OK when they handled the sweep, -1 when the plan is not theirs.
int try_life_swar(const Plan& pl, const void* src, void* dst, cudaStream_t st);
bool life2_accepts(const sb200_desc& d, const Plan& pl);  // SB200_FLAG_DOUBLE_STEP
bool life_multi_accepts(const sb200_desc& d, const Plan& pl, int gens);  // gens = 2 (DOUBLE_STEP) or 4 (QUAD_STEP)
int try_tile2d(const Plan& pl, const void* src, void* dst, cudaStream_t st);
int try_small2d(const Plan& pl, const void* src, void* dst, cudaStream_t st);   // small2d.cu: radius-1 shapes, L2-resident grids
int try_diffusion3d(const Plan& pl, const void* src, void* dst, cudaStream_t st);
int try_diffusion3d_double(const Plan& pl, const void* src, void* dst, cudaStream_t st);  // SB200_FLAG_DOUBLE_STEP (stream3d2.cu)
bool diffusion2_accepts(const sb200_desc& d, const Plan& pl);
int try_gather_stream(const Plan& pl, const void* src, void* dst, cudaStream_t st);
int try_gather_stream3d(const Plan& pl, const void* src, void* dst, cudaStream_t st);
int try_box3d(const Plan& pl, const void* src, void* dst, cudaStream_t st);   // Window(1,3) / Moore(1,3) at compile time (box3d.cu)
int try_scatter_fast(const Plan& pl, const void* src, void* dst, cudaStream_t st);
int try_multi_tile2d(const Plan* const* plans, const sb200_term* terms, int nterms, void* dst, cudaStream_t st, bool dry);   // multi_tile.cu
int try_scatter_stream(const Plan& pl, const void* src, void* dst, cudaStream_t st, int x_lo, int x_hi, int y_lo, int y_hi);

int num_sms();

// api.cu: the dispatch behind sb200_gather, and the question sb200_iterate / the slab plans ask before scheduling
// several generations per launch.
int do_gather(const sb200_desc* d, const void* src, void* dst, cudaStream_t st);
bool multistep_accepts(const sb200_desc* d);
// Launch time of the Life kernels by generations per launch, relative to one generation (api.cu; measured, tools/life_gens_probe.py),
// and the size with the least time per generation (what long runs and the slab plans' cycles are made of).
constexpr int kMaxGens = 8;
extern const double kLifeLaunchCost[kMaxGens + 1];
// the same for packed -> packed launches (life_bit_kernel<G, bits, bits>), and the size packed runs are made of
extern const double kLifePackedLaunchCost[kMaxGens + 1];
constexpr int kLifePackedBulkGens = 6;
inline int life_bulk_gens() {
    int best = 1;
    for (int g = 2; g <= kMaxGens; g++)
        if (kLifeLaunchCost[g] / g < kLifeLaunchCost[best] / best) best = g;
    return best;
}

}  // namespace sb
