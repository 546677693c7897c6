#!/bin/bash
# r02x: packed Life state (SB200_FLAG_SRC_BITS / _DST_BITS): parity of the three launch forms, packed runs in sb200_iterate and in
# the slab plans, launch times, the driver's bench line
O=gpurun_out/r02x
mkdir -p $O
S=$O/status.txt
date > $S
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 200 python tools/life_gens_probe.py > $O/probe.log 2>&1; echo "probe rc=$?" >> $S
timeout 300 python tools/plan_probe.py life > $O/plan_probe_life.log 2>&1; echo "plan probe rc=$?" >> $S
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench k20 rc=$?" >> $S
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_cases.py --quick > $O/memcheck_quick.log 2>&1; echo "memcheck rc=$?" >> $S
date >> $S
