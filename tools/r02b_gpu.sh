#!/bin/bash
# r02b: one-halo-lane Life + eight generations per launch + Remove axes in stream3d2 as the DEFAULT build; packed folds re-tested
O=gpurun_out/r02b
mkdir -p $O
S=$O/status.txt
date > $S
LIBDIR=$PWD/stencils.jl_b200/lib
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 300 python bench.py --no-extras --steps 1000 > $O/bench_life_default.json 2> $O/bench_life_default.err; echo "bench life default rc=$?" >> $S
for t in 1 2 3; do
  SB200_LB_TASKS=$t timeout 300 python bench.py --no-extras --steps 1000 > $O/bench_life_t$t.json 2> $O/bench_life_t$t.err; echo "bench tasks=$t rc=$?" >> $S
done
SB200_OCT_STEP=0 timeout 300 python bench.py --no-extras --steps 1000 > $O/bench_life_quad.json 2> $O/bench_life_quad.err; echo "bench quad rc=$?" >> $S
timeout 300 python bench.py --no-extras --steps 20 --warmup 5 > $O/bench_life_20.json 2> $O/bench_life_20.err; echo "bench 20 steps rc=$?" >> $S
L=$LIBDIR/libstencils_b200_pk.so
SB200_LIB=$L timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -k "two_steps" -x -q > $O/pytest_pk.log 2>&1; echo "pytest pk rc=$?" >> $S
SB200_LIB=$L timeout 200 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras > $O/bench_diffusion_pk.json 2> $O/bench_diffusion_pk.err; echo "bench diffusion pk rc=$?" >> $S
timeout 200 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras > $O/bench_diffusion_default.json 2> $O/bench_diffusion_default.err; echo "bench diffusion rc=$?" >> $S
timeout 150 ncu --set full --import-source on --clock-control none -k regex:life_bit -s 4 -c 1 -f -o $O/life_oct \
    python bench.py --steps 200 --warmup 4 --no-extras > $O/ncu_life.log 2>&1; echo "ncu life rc=$?" >> $S
ncu -i $O/life_oct.ncu-rep --page raw --csv > $O/life_oct_raw.csv 2>/dev/null
date >> $S
