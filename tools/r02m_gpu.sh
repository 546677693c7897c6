#!/bin/bash
# r02m: the independent-path parity tests; runs per strip for the launch-bound 1000 x 1000 grid
O=gpurun_out/r02m
mkdir -p $O
S=$O/status.txt
date > $S
timeout 600 python -m pytest tests/test_gpu_independent.py -q > $O/pytest_independent.log 2>&1; echo "pytest independent rc=$?" >> $S
for n in 0 10 20 37 74 148 296; do
  SB200_S2_NRUNS=$n timeout 100 python bench.py --workload mean1000 --no-extras > $O/bench_mean1000_n$n.json 2> $O/bench_mean1000_n$n.err; echo "mean1000 nruns=$n rc=$?" >> $S
done
date >> $S
