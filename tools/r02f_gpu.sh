#!/bin/bash
# r02f: box3d v2 (producer back-off, predicated end-lane loads, PAD instantiation), bulk-copy producer for Halo-padded sources,
# producer back-off in every kernel as an A/B variant (bo200), new tests (Layered, update_boundary, FMA, long runs)
O=gpurun_out/r02f
mkdir -p $O
S=$O/status.txt
date > $S
LIBDIR=$PWD/stencils.jl_b200/lib
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 600 python bench.py --steps 1000 > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?" >> $S
SB200_LIB=$LIBDIR/libstencils_b200_bo200.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -x > $O/pytest_bo200.log 2>&1; echo "pytest bo200 rc=$?" >> $S
SB200_LIB=$LIBDIR/libstencils_b200_bo200.so timeout 600 python bench.py --steps 1000 > $O/bench_bo200.json 2> $O/bench_bo200.err; echo "bench bo200 rc=$?" >> $S
timeout 150 ncu --set full --import-source on --clock-control none -k regex:box3d -s 2 -c 1 -f -o $O/window3d \
    python bench.py --workload window3d --no-extras > $O/ncu_window3d.log 2>&1; echo "ncu window3d rc=$?" >> $S
ncu -i $O/window3d.ncu-rep --page raw --csv > $O/window3d_raw.csv 2>/dev/null
ncu -i $O/window3d.ncu-rep --page source --csv --print-source sass > $O/window3d_sass.csv 2>/dev/null
timeout 150 ncu --set full --import-source on --clock-control none -k regex:stream2d -s 2 -c 1 -f -o $O/mean_halo \
    python bench.py --workload mean_halo --no-extras > $O/ncu_mean_halo.log 2>&1; echo "ncu mean_halo rc=$?" >> $S
ncu -i $O/mean_halo.ncu-rep --page raw --csv > $O/mean_halo_raw.csv 2>/dev/null
rm -f $O/*.ncu-rep
date >> $S
