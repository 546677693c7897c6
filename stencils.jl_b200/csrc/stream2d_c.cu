// stream2d_c.cu — instantiations of the streaming 2-D gather (stream2d.cuh): Circle R=2..4.
#include "stream2d.cuh"

namespace sb {

template <typename T> static int group_t(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    S2Params<T> p;
    if (!s2_accepts<T>(pl, src, dst, p)) return -1;
    if (pl.shape_tag == SB200_CIRCLE && pl.d.radius == 2) return s2_dispatch_reducer<T, SB200_CIRCLE, 2>(p, pl.d.reducer, st);
    if (pl.shape_tag == SB200_CIRCLE && pl.d.radius == 3) return s2_dispatch_reducer<T, SB200_CIRCLE, 3>(p, pl.d.reducer, st);
    if (pl.shape_tag == SB200_CIRCLE && pl.d.radius == 4) return s2_dispatch_reducer<T, SB200_CIRCLE, 4>(p, pl.d.reducer, st);
    return -1;
}

int s2_group_c(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    if (pl.d.eltype == SB200_F32) return group_t<float>(pl, src, dst, st);
    if (pl.d.eltype == SB200_F64) return group_t<double>(pl, src, dst, st);
    return -1;
}

}  // namespace sb
