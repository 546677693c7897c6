#!/bin/bash
# r02o: nested extrema folds — width-ordered (default build) and without the pair fold (s2pf0) — on Circle(4) max
O=gpurun_out/r02o
mkdir -p $O
S=$O/status.txt
date > $S
LIBDIR=$PWD/stencils.jl_b200/lib
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_independent.py -q -x -k "stream2d or reducers_2d or circle or Circle or hand_built" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $S
SB200_LIB=$LIBDIR/libstencils_b200_s2pf0.so timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_independent.py -q -x -k "stream2d or reducers_2d or circle or Circle or hand_built" > $O/pytest_s2pf0.log 2>&1; echo "pytest s2pf0 rc=$?" >> $S
for i in 1 2; do
  timeout 200 python bench.py --workload circle --no-extras > $O/bench_circle_$i.json 2> $O/bench_circle_$i.err; echo "bench circle rc=$?" >> $S
  SB200_LIB=$LIBDIR/libstencils_b200_s2pf0.so timeout 200 python bench.py --workload circle --no-extras > $O/bench_circle_s2pf0_$i.json 2> $O/bench_circle_s2pf0_$i.err; echo "bench circle s2pf0 rc=$?" >> $S
done
date >> $S
