// stream2d.cu — dispatch of 2-D float gathers to the TMA-fed streaming kernels (stream2d.cuh).
#include "stream2d.cuh"

namespace sb {

int try_tile2d(const Plan& pl, const void* src, void* dst, cudaStream_t st) {
    const sb200_desc& d = pl.d;
    if (d.flags & SB200_FLAG_NO_TMA) return -1;
    if (d.ndim != 2 || (d.eltype != SB200_F32 && d.eltype != SB200_F64)) return -1;
    if (d.reducer == SB200_LIFE || pl.shape_tag < 0 || pl.shape_ndim != 2) return -1;
    if (pl.dd.n[1] == 0) return SB200_OK;
    int rc = -1;
    switch (pl.shape_tag) {
    case SB200_WINDOW: rc = s2_group_a(pl, src, dst, st); break;
    case SB200_MOORE: case SB200_VONNEUMANN: rc = s2_group_b(pl, src, dst, st); break;
    case SB200_CIRCLE: rc = s2_group_c(pl, src, dst, st); break;
    case SB200_CROSS: case SB200_DIAMOND: rc = s2_group_d(pl, src, dst, st); break;
    default: break;
    }
    const bool fma = d.reducer == SB200_KERNELDOT && (d.flags & SB200_FLAG_ALLOW_FMA) && pl.shape_tag == SB200_WINDOW && d.radius <= 3;
    if (rc == SB200_OK) set_kernel_name(fma ? "stream2d_kernel<fma>" : "stream2d_kernel");
    return rc;
}

}  // namespace sb
