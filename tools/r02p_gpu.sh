#!/bin/bash
# r02p: producer back-off (try_wait suspend hint) on the issue-bound stream2d folds
O=gpurun_out/r02p
mkdir -p $O
S=$O/status.txt
date > $S
LIBDIR=$PWD/stencils.jl_b200/lib
for wl in kernel kernel_fma circle mean; do
  timeout 200 python bench.py --workload $wl --no-extras > $O/bench_${wl}.json 2> $O/bench_${wl}.err; echo "bench $wl rc=$?" >> $S
  SB200_LIB=$LIBDIR/libstencils_b200_s2bo1k.so timeout 200 python bench.py --workload $wl --no-extras > $O/bench_${wl}_s2bo1k.json 2> $O/bench_${wl}_s2bo1k.err; echo "bench $wl s2bo1k rc=$?" >> $S
done
timeout 150 ncu --set full --import-source on --clock-control none -k regex:stream2d -s 3 -c 1 -f -o $O/kernel python bench.py --workload kernel --no-extras > $O/ncu_kernel.log 2>&1
ncu -i $O/kernel.ncu-rep --page source --csv --print-source sass > $O/kernel_sass.csv 2>/dev/null; rm -f $O/kernel.ncu-rep
date >> $S
