#!/bin/bash
# Final 1-GPU evidence run of the round: whole GPU suite, sanitizers on the final kernels, bench lines, ncu captures.
O=gpurun_out/r01l
mkdir -p $O
S=$O/status.txt
date > $S
(
  timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_cases.py > $O/memcheck.log 2>&1
  echo "memcheck rc=$?" >> $S
  timeout 200 compute-sanitizer --tool synccheck --error-exitcode 9 python tests/sanitize_cases.py --quick > $O/synccheck.log 2>&1
  echo "synccheck rc=$?" >> $S
) &
timeout 420 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $S
wait
for v in 1 0; do
  SB200_DIFFUSION_DOUBLE_STEP=$v timeout 200 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras \
      > $O/bench_diffusion_ds$v.json 2> $O/bench_diffusion_ds$v.err; echo "bench diffusion ds$v rc=$?" >> $S
done
timeout 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?" >> $S
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stream3d|halo|push|wait|signal" -c 40 --csv --log-file $O/launches_diffusion.csv \
    python bench.py --workload diffusion --steps 24 --warmup 4 --no-extras > $O/launches_diffusion.log 2>&1
SB200_DIFFUSION_DOUBLE_STEP=0 timeout 150 ncu --set full --import-source on --clock-control none -k regex:stream3d -s 3 -c 1 -f -o gpurun_out/r01l_diffusion \
    python bench.py --workload diffusion --steps 4 --warmup 4 --no-extras > $O/ncu_diffusion.log 2>&1; echo "ncu diffusion rc=$?" >> $S
timeout 150 ncu --set full --import-source on --clock-control none -k regex:stream3d2 -s 1 -c 1 -f -o gpurun_out/r01l_diffusion2 \
    python bench.py --workload diffusion --steps 4 --warmup 4 --no-extras > $O/ncu_diffusion2.log 2>&1; echo "ncu diffusion2 rc=$?" >> $S
date >> $S
