#!/bin/bash
# r02at: stream3d2_kernel with scalar FADDs for the x-neighbour adds (no pair assembly), A/B
O=gpurun_out/r02at
mkdir -p $O
S=$O/status.txt
date > $S
for v in default d2xs; do
  if [ $v = default ]; then unset SB200_LIB; else export SB200_LIB=$PWD/stencils.jl_b200/lib/libstencils_b200_$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "two_steps or diffusion" > $O/pytest_$v.log 2>&1; echo "$v pytest rc=$?" >> $S
  for rep in 1 2; do
    timeout 200 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras > $O/bench_${v}_$rep.json 2> $O/bench_${v}_$rep.err; echo "$v bench $rep rc=$?" >> $S
  done
done
date >> $S
