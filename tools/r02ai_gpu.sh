#!/bin/bash
# r02ai: final evidence run (packed Life runs) on one GPU — whole suite, smoke, sanitizers, ncu launch list + one --set full capture per dominant kernel,
# the bench exactly as the driver runs it (--steps 20 --warmup 5) and with its own defaults
O=gpurun_out/r02ai
mkdir -p $O
S=$O/status.txt
date > $S
(
  timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_cases.py > $O/memcheck.log 2>&1
  echo "memcheck rc=$?" >> $S
  timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python tests/sanitize_cases.py --quick > $O/synccheck.log 2>&1
  echo "synccheck rc=$?" >> $S
) &
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $S
wait
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver_k20.json 2> $O/bench_driver_k20.err; echo "bench k20 rc=$?" >> $S
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?" >> $S
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?" >> $S
bash profiles/collect.sh r02 > $O/collect.log 2>&1; echo "collect rc=$?" >> $S
date >> $S
timeout 150 ncu --set full --import-source on --clock-control none -k regex:stream3d2 -s 2 -c 1 -f -o gpurun_out/r02_diffusion2 \
    python bench.py --workload diffusion --steps 8 --warmup 4 --no-extras > $O/ncu_diffusion2.log 2>&1; echo "ncu diffusion2 rc=$?" >> $S
ncu -i gpurun_out/r02_diffusion2.ncu-rep --page raw --csv > gpurun_out/r02_diffusion2_raw.csv 2>/dev/null; rm -f gpurun_out/r02_diffusion2.ncu-rep
date >> $S
