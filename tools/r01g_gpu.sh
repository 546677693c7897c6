#!/bin/bash
# One GPU call, most important first; every leg has its own timeout and writes under gpurun_out/r01g/ as it goes.
# gpurun --timeout 660 -- 'bash tools/r01g_gpu.sh'
O=gpurun_out/r01g
mkdir -p $O
S=$O/status.txt
date > $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv >> $S 2>&1

# 1. torch-free parity of every kernel family (incl. stream3d2_kernel) against the oracle: seconds
timeout 120 python tests/sanitize_cases.py > $O/cases.log 2>&1; echo "cases rc=$?" >> $S

# 2. A/B: iterated diffusion 1024^3, one vs two steps per launch (clean GPU)
for v in 0 1; do
  SB200_DIFFUSION_DOUBLE_STEP=$v timeout 200 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras \
      > $O/bench_diffusion_ds$v.json 2> $O/bench_diffusion_ds$v.err; echo "bench ds$v rc=$?" >> $S
done

# 3. the new kernel's parity tests (small + full size)
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -k "two_steps" -x -q > $O/pytest_two_steps.log 2>&1
echo "pytest two_steps rc=$?" >> $S

# 4. sanitizers (background) while the whole GPU suite runs with the two-step schedule switched on
(
  timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_cases.py > $O/memcheck.log 2>&1
  echo "memcheck rc=$?" >> $S
  timeout 200 compute-sanitizer --tool synccheck --error-exitcode 9 python tests/sanitize_cases.py --quick > $O/synccheck.log 2>&1
  echo "synccheck rc=$?" >> $S
  timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/sanitize_cases.py --quick > $O/racecheck.log 2>&1
  echo "racecheck rc=$?" >> $S
) &
SB200_DIFFUSION_DOUBLE_STEP=1 timeout 420 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_ds1.log 2>&1
echo "pytest gpu (ds=1) rc=$?" >> $S
wait

# 5. ncu of the two-step kernel (one launch, full set)
SB200_DIFFUSION_DOUBLE_STEP=1 timeout 150 ncu --set full --clock-control none -k regex:stream3d2 -c 1 -f -o $O/diffusion2 \
    python bench.py --workload diffusion --steps 4 --warmup 4 --no-extras > $O/ncu_diffusion2.log 2>&1
echo "ncu rc=$?" >> $S
ncu -i $O/diffusion2.ncu-rep --page raw --csv > $O/diffusion2_raw.csv 2>/dev/null

# 6. the default bench line
timeout 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?" >> $S
date >> $S
