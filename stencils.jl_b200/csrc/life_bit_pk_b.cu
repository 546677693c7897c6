// life_bit_pk_b.cu — life_bit_kernel<G, packed source, packed or byte dest>: the launches inside and at the end of a packed run.
#include "life_bit.cuh"

namespace sb {

int launch_life_bit_from_bits(int gens, bool out_bits, const LifeParams& p, cudaStream_t st) {
#if SB200_LB_ONE_HALO_LANE
    return out_bits ? launch_bit_gens<LB_BITS, true>(gens, p, st) : launch_bit_gens<LB_BITS, false>(gens, p, st);
#else
    set_error("packed Life state needs the one-halo-lane build");
    return SB200_EUNSUPPORTED;
#endif
}

}  // namespace sb
