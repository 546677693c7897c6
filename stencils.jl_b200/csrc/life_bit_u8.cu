// life_bit_u8.cu — life_bit_kernel<G, byte source, byte dest>: the launches a single sb200_gather with SB200_FLAG_GENS sees.
#include "life_bit.cuh"

namespace sb {

int launch_life_bit_u8(int gens, bool cells01, const LifeParams& p, cudaStream_t st) {
    return cells01 ? launch_bit_gens<LB_U8_01, false>(gens, p, st) : launch_bit_gens<LB_U8, false>(gens, p, st);
}

}  // namespace sb
