#!/bin/bash
# r02h: life_bit_kernel with the 9-op Conway table + funnel shifts, packed box3d vs scalar, whole suite
O=gpurun_out/r02h
mkdir -p $O
S=$O/status.txt
date > $S
LIBDIR=$PWD/stencils.jl_b200/lib
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 200 python bench.py --steps 1000 --no-extras > $O/bench_life.json 2> $O/bench_life.err; echo "bench life rc=$?" >> $S
timeout 200 python bench.py --steps 20 --warmup 5 --no-extras > $O/bench_life_k20.json 2> $O/bench_life_k20.err; echo "bench life k20 rc=$?" >> $S
timeout 200 python bench.py --workload window3d --no-extras > $O/bench_window3d.json 2> $O/bench_window3d.err; echo "bench window3d rc=$?" >> $S
SB200_LIB=$LIBDIR/libstencils_b200_b3np.so timeout 200 python bench.py --workload window3d --no-extras > $O/bench_window3d_b3np.json 2> $O/bench_window3d_b3np.err; echo "bench b3np rc=$?" >> $S
timeout 150 ncu --set full --import-source on --clock-control none -k regex:life_bit -s 6 -c 1 -f -o $O/life \
    python bench.py --steps 200 --warmup 16 --no-extras > $O/ncu_life.log 2>&1; echo "ncu life rc=$?" >> $S
ncu -i $O/life.ncu-rep --page raw --csv > $O/life_raw.csv 2>/dev/null
ncu -i $O/life.ncu-rep --page source --csv --print-source sass > $O/life_sass.csv 2>/dev/null
rm -f $O/*.ncu-rep
date >> $S
