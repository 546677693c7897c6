#!/bin/bash
# r02n: whole suite + sanitizers with the slab-plan cases + driver-style bench after the last changes
O=gpurun_out/r02n
mkdir -p $O
S=$O/status.txt
date > $S
(
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_cases.py > $O/memcheck.log 2>&1
  echo "memcheck rc=$?" >> $S
  timeout 400 compute-sanitizer --tool synccheck --error-exitcode 9 python tests/sanitize_cases.py --quick > $O/synccheck.log 2>&1
  echo "synccheck rc=$?" >> $S
) &
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $S
wait
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver_k20.json 2> $O/bench_driver_k20.err; echo "bench k20 rc=$?" >> $S
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?" >> $S
date >> $S
