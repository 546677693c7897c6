#!/bin/bash
# r02ay: stream3d2_kernel with the split-phase level barrier (mbarrier arrive after level 1 of plane i, wait before level 2 of plane i-1), A/B
O=gpurun_out/r02ay
mkdir -p $O
S=$O/status.txt
date > $S
export SB200_LIB=$PWD/stencils.jl_b200/lib/libstencils_b200_d2split.so
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_plan.py -m gpu -q -x -k "two_steps or diffusion" > $O/pytest_split.log 2>&1; echo "split pytest rc=$?" >> $S
unset SB200_LIB
for rep in 1 2; do
  for v in default d2split; do
    if [ $v = default ]; then unset SB200_LIB; else export SB200_LIB=$PWD/stencils.jl_b200/lib/libstencils_b200_$v.so; fi
    timeout 100 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras > $O/diffusion_${v}_$rep.json 2> $O/diffusion_${v}_$rep.err; echo "$v diffusion $rep rc=$?" >> $S
  done
done
date >> $S
