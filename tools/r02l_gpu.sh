#!/bin/bash
# r02l: vector halo loads A/B (stream2d), one-slab plan vs sb200_iterate on one box, launch list + stream3d2 capture
O=gpurun_out/r02l
mkdir -p $O
S=$O/status.txt
date > $S
LIBDIR=$PWD/stencils.jl_b200/lib
for wl in circle kernel kernel_fma; do
  timeout 200 python bench.py --workload $wl --no-extras > $O/bench_${wl}.json 2> $O/bench_${wl}.err; echo "bench $wl rc=$?" >> $S
  SB200_LIB=$LIBDIR/libstencils_b200_s2vh.so timeout 200 python bench.py --workload $wl --no-extras > $O/bench_${wl}_s2vh.json 2> $O/bench_${wl}_s2vh.err; echo "bench $wl s2vh rc=$?" >> $S
done
SB200_LIB=$LIBDIR/libstencils_b200_s2vh.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -k "stream2d or reducers_2d or circle or kernel" > $O/pytest_s2vh.log 2>&1; echo "pytest s2vh rc=$?" >> $S
timeout 300 python tools/plan_probe.py diffusion > $O/plan_probe_diffusion.log 2>&1; echo "probe diffusion rc=$?" >> $S
timeout 300 python tools/plan_probe.py life > $O/plan_probe_life.log 2>&1; echo "probe life rc=$?" >> $S
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"life_|stream|gather_|scatter_|box3d|halo_kernel|plan_|combine" -c 400 --csv --log-file gpurun_out/r02_launches_life.csv \
    python bench.py --steps 200 --warmup 16 --no-extras > $O/launches_life.log 2>&1; echo "launch list rc=$?" >> $S
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stream2d" -c 40 --csv --log-file $O/launches_mean1000.csv \
    python bench.py --workload mean1000 --no-extras > $O/launches_mean1000.log 2>&1; echo "launch list mean1000 rc=$?" >> $S
timeout 150 ncu --set full --import-source on --clock-control none -k regex:stream3d2 -s 2 -c 1 -f -o gpurun_out/r02_diffusion2 \
    python bench.py --workload diffusion --steps 8 --warmup 4 --no-extras > $O/ncu_diffusion2.log 2>&1; echo "ncu diffusion2 rc=$?" >> $S
ncu -i gpurun_out/r02_diffusion2.ncu-rep --page raw --csv > gpurun_out/r02_diffusion2_raw.csv 2>/dev/null; rm -f gpurun_out/r02_diffusion2.ncu-rep
date >> $S
