// life_bit_pk_a.cu — life_bit_kernel<G, byte source, packed dest>: the first launch of a run that keeps its state packed.
#include "life_bit.cuh"

namespace sb {

int launch_life_bit_to_bits(int gens, bool cells01, const LifeParams& p, cudaStream_t st) {
#if SB200_LB_ONE_HALO_LANE
    return cells01 ? launch_bit_gens<LB_U8_01, true>(gens, p, st) : launch_bit_gens<LB_U8, true>(gens, p, st);
#else
    set_error("packed Life state needs the one-halo-lane build");
    return SB200_EUNSUPPORTED;
#endif
}

}  // namespace sb
