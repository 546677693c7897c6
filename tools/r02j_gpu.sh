#!/bin/bash
# r02j: evidence run on one GPU — whole suite, smoke, sanitizers, ncu launch list + one --set full capture per dominant kernel,
# the bench exactly as the driver runs it (--steps 20 --warmup 5) and with its own defaults
O=gpurun_out/r02j
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
LIBDIR=$PWD/stencils.jl_b200/lib
for wl in circle kernel mean; do
  SB200_LIB=$LIBDIR/libstencils_b200_s2ew.so timeout 200 python bench.py --workload $wl --no-extras > $O/bench_${wl}_s2ew.json 2> $O/bench_${wl}_s2ew.err; echo "bench $wl s2ew rc=$?" >> $S
  timeout 200 python bench.py --workload $wl --no-extras > $O/bench_${wl}.json 2> $O/bench_${wl}.err; echo "bench $wl rc=$?" >> $S
done
SB200_LIB=$LIBDIR/libstencils_b200_s2ew.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "stream2d or reducers_2d" > $O/pytest_s2ew.log 2>&1; echo "pytest s2ew rc=$?" >> $S
du -sh gpurun_out >> $S
date >> $S
