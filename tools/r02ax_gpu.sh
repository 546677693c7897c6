#!/bin/bash
# r02ax: the clean rebuild of the final sources on one GPU: sanitizers, whole suite, the bench with its own defaults (other_configs), the reference arm
O=gpurun_out/r02ax
mkdir -p $O
S=$O/status.txt
date > $S
(
  timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_cases.py > $O/memcheck.log 2>&1
  echo "memcheck rc=$?" >> $S
) &
timeout 300 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
wait
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?" >> $S
timeout 200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?" >> $S
date >> $S
