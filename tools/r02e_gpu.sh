#!/bin/bash
# r02e: box3d (compile-time Window(1,3) / Moore(1,3)) parity + rows-per-thread variants, the new sb200_iterate schedule, full bench line
O=gpurun_out/r02e
mkdir -p $O
S=$O/status.txt
date > $S
LIBDIR=$PWD/stencils.jl_b200/lib
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 300 python tests/sanitize_cases.py > $O/cases.log 2>&1; echo "cases rc=$?" >> $S
timeout 200 python bench.py --workload window3d --no-extras > $O/bench_window3d_rt4.json 2> $O/bench_window3d_rt4.err; echo "bench window3d rt4 rc=$?" >> $S
for v in b3rt2 b3rt3; do
  SB200_LIB=$LIBDIR/libstencils_b200_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "box3d or stream3d" > $O/pytest_$v.log 2>&1; echo "pytest $v rc=$?" >> $S
  SB200_LIB=$LIBDIR/libstencils_b200_$v.so timeout 200 python bench.py --workload window3d --no-extras > $O/bench_window3d_$v.json 2> $O/bench_window3d_$v.err; echo "bench $v rc=$?" >> $S
done
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_default_k20.json 2> $O/bench_default_k20.err; echo "bench k20 rc=$?" >> $S
timeout 150 ncu --set full --import-source on --clock-control none -k regex:box3d -s 2 -c 1 -f -o $O/window3d \
    python bench.py --workload window3d --no-extras > $O/ncu_window3d.log 2>&1; echo "ncu window3d rc=$?" >> $S
ncu -i $O/window3d.ncu-rep --page raw --csv > $O/window3d_raw.csv 2>/dev/null
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_cases.py --quick > $O/memcheck_quick.log 2>&1; echo "memcheck quick rc=$?" >> $S
date >> $S
