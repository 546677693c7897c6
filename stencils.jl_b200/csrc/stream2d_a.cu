// stream2d_a.cu — instantiations of the streaming 2-D gather (stream2d.cuh): Window R=1..3.
#include "stream2d.cuh"

namespace sb {

template <typename T> static int group_t(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    S2Params<T> p;
    if (!s2_accepts<T>(pl, src, dst, p)) return -1;
    if (pl.d.reducer == SB200_KERNELDOT && (pl.d.flags & SB200_FLAG_ALLOW_FMA) && pl.shape_tag == SB200_WINDOW) {
        // the caller allows contraction: acc = fma(v_k, w_k, acc), one rounding per tap instead of two (49 FFMA instead of 98
        // FMUL + FADD for the 7 x 7 kernel, which is bound by FP32 issue, not HBM)
        if (pl.d.radius == 1) return s2_launch<T, SB200_WINDOW, 1, S2_KDOT_FMA>(p, st);
        if (pl.d.radius == 2) return s2_launch<T, SB200_WINDOW, 2, S2_KDOT_FMA>(p, st);
        if (pl.d.radius == 3) return s2_launch<T, SB200_WINDOW, 3, S2_KDOT_FMA>(p, st);
    }
    if (pl.shape_tag == SB200_WINDOW && pl.d.radius == 1) return s2_dispatch_reducer<T, SB200_WINDOW, 1>(p, pl.d.reducer, st);
    if (pl.shape_tag == SB200_WINDOW && pl.d.radius == 2) return s2_dispatch_reducer<T, SB200_WINDOW, 2>(p, pl.d.reducer, st);
    if (pl.shape_tag == SB200_WINDOW && pl.d.radius == 3) return s2_dispatch_reducer<T, SB200_WINDOW, 3>(p, pl.d.reducer, st);
    return -1;
}

int s2_group_a(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    if (pl.d.eltype == SB200_F32) return group_t<float>(pl, src, dst, st);
    if (pl.d.eltype == SB200_F64) return group_t<double>(pl, src, dst, st);
    return -1;
}

}  // namespace sb
