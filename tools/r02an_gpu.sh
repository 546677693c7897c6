#!/bin/bash
# r02an: box3d_kernel with the plane loop unrolled by two, A/B on Window(1,3) mean 768^3
O=gpurun_out/r02an
mkdir -p $O
S=$O/status.txt
date > $S
for v in default b3u2; do
  if [ $v = default ]; then unset SB200_LIB; else export SB200_LIB=$PWD/stencils.jl_b200/lib/libstencils_b200_$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "box3d or window3d or 3d" > $O/pytest_$v.log 2>&1; echo "$v pytest rc=$?" >> $S
  for rep in 1 2; do
    timeout 200 python bench.py --workload window3d --steps 10 --warmup 3 --no-extras > $O/bench_${v}_$rep.json 2> $O/bench_${v}_$rep.err; echo "$v bench $rep rc=$?" >> $S
  done
done
date >> $S
