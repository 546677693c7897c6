#!/bin/bash
# r02av: stream3d2_kernel producers / ring depth re-tuned after the instruction cuts (A/B: 4 producer warps, 8 stages, both)
O=gpurun_out/r02av
mkdir -p $O
S=$O/status.txt
date > $S
for v in d2p4 d2st8 d2p4st8; do
  SB200_LIB=$PWD/stencils.jl_b200/lib/libstencils_b200_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "two_steps or diffusion" > $O/pytest_$v.log 2>&1; echo "$v pytest rc=$?" >> $S
done
for rep in 1 2; do
  for v in default d2p4 d2st8 d2p4st8; do
    if [ $v = default ]; then unset SB200_LIB; else export SB200_LIB=$PWD/stencils.jl_b200/lib/libstencils_b200_$v.so; fi
    timeout 200 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras > $O/diffusion_${v}_$rep.json 2> $O/diffusion_${v}_$rep.err; echo "$v diffusion $rep rc=$?" >> $S
  done
done
date >> $S
