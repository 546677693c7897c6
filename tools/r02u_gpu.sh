#!/bin/bash
# r02u: launches of seven Life generations (sb200_iterate bulk size, slab-plan cycles of 126 = 18 x 7), bench N = 1 on whole plan cycles
O=gpurun_out/r02u
mkdir -p $O
S=$O/status.txt
date > $S
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver_k20.json 2> $O/bench_driver_k20.err; echo "bench k20 rc=$?" >> $S
timeout 300 python tools/plan_probe.py life > $O/plan_probe_life.log 2>&1; echo "plan probe life rc=$?" >> $S
SB200_POW2_STEPS=1 timeout 300 python tools/plan_probe.py life > $O/plan_probe_life_pow2.log 2>&1; echo "plan probe life pow2 rc=$?" >> $S
timeout 300 python bench.py --gpus 1 --workload diffusion --steps 20 --warmup 5 --no-extras > $O/bench_diffusion_k20.json 2> $O/bench_diffusion_k20.err; echo "bench diffusion k20 rc=$?" >> $S
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_cases.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $S
date >> $S
