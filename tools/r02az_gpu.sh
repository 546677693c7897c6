#!/bin/bash
# r02az: box3d_kernel Float32 sums with the mixed fold (centre taps packed, x-neighbour taps scalar on the halves of the pair), A/B
O=gpurun_out/r02az
mkdir -p $O
S=$O/status.txt
date > $S
export SB200_LIB=$PWD/stencils.jl_b200/lib/libstencils_b200_b3mix.so
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "box3d or window3d or box or 3d" > $O/pytest_mix.log 2>&1; echo "mix pytest rc=$?" >> $S
unset SB200_LIB
for rep in 1 2; do
  for v in default b3mix; do
    if [ $v = default ]; then unset SB200_LIB; else export SB200_LIB=$PWD/stencils.jl_b200/lib/libstencils_b200_$v.so; fi
    timeout 100 python bench.py --workload window3d --steps 20 --warmup 4 --no-extras > $O/window3d_${v}_$rep.json 2> $O/window3d_${v}_$rep.err; echo "$v window3d $rep rc=$?" >> $S
  done
done
date >> $S
