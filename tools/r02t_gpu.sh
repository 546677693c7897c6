#!/bin/bash
# r02t: the launch-size split of sb200_iterate (SB200_FLAG_GENS, 1 .. 8 generations per launch): whole GPU suite, smoke, per-size launch
# times, the driver's bench line with the new split and with the round-2 power-of-two split
O=gpurun_out/r02t
mkdir -p $O
S=$O/status.txt
date > $S
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 300 python tools/life_gens_probe.py > $O/life_gens_probe.log 2>&1; echo "probe rc=$?" >> $S
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver_k20.json 2> $O/bench_driver_k20.err; echo "bench k20 rc=$?" >> $S
SB200_POW2_STEPS=1 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > $O/bench_driver_k20_pow2.json 2> $O/bench_driver_k20_pow2.err; echo "bench k20 pow2 rc=$?" >> $S
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_cases.py --quick > $O/memcheck_quick.log 2>&1; echo "memcheck quick rc=$?" >> $S
date >> $S
