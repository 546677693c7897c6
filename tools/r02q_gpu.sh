#!/bin/bash
# r02q: shift-free rolled folds in stream2d (R >= 2): parity + the issue-bound benches
O=gpurun_out/r02q
mkdir -p $O
S=$O/status.txt
date > $S
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_independent.py tests/test_gpu_api.py -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $S
for wl in kernel kernel_fma circle mean mean_halo; do
  timeout 200 python bench.py --workload $wl --no-extras > $O/bench_${wl}.json 2> $O/bench_${wl}.err; echo "bench $wl rc=$?" >> $S
done
date >> $S
