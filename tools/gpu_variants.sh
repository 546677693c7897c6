#!/bin/bash
# A/B of library variants built with tools/build_variant.sh:  gpu_variants.sh <tag> <variant> [<variant> ...]
O=gpurun_out/$1; shift
mkdir -p $O
S=$O/status.txt
date > $S
timeout 120 python tests/sanitize_cases.py > $O/cases.log 2>&1; echo "cases rc=$?" >> $S
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -k "two_steps" -x -q > $O/pytest_two_steps.log 2>&1
echo "pytest two_steps rc=$?" >> $S
SB200_DIFFUSION_DOUBLE_STEP=0 timeout 200 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras > $O/bench_ds0.json 2> $O/bench_ds0.err
SB200_DIFFUSION_DOUBLE_STEP=1 timeout 200 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras > $O/bench_default.json 2> $O/bench_default.err
for spec in "$@"; do        # <variant>[:<SB200_DIFFUSION_DOUBLE_STEP value>]
  v=${spec%%:*}; ds=1; [[ "$spec" == *:* ]] && ds=${spec##*:}
  L=$PWD/stencils.jl_b200/lib/libstencils_b200_$v.so
  SB200_LIB=$L timeout 120 python tests/sanitize_cases.py > $O/cases_$v.log 2>&1; echo "cases $v rc=$?" >> $S
  SB200_LIB=$L timeout 200 python -m pytest tests/test_gpu_parity.py -k "stream3d or two_steps or 1d_3d or iterate" -x -q > $O/pytest_$v.log 2>&1; echo "pytest $v rc=$?" >> $S
  SB200_LIB=$L SB200_DIFFUSION_DOUBLE_STEP=$ds timeout 200 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras > $O/bench_$v.json 2> $O/bench_$v.err
  echo "bench $v (ds=$ds) rc=$?" >> $S
done
SB200_DIFFUSION_DOUBLE_STEP=1 timeout 150 ncu --set full --import-source on --clock-control none -k regex:stream3d2 -c 1 -f -o $O/diffusion2 \
    python bench.py --workload diffusion --steps 4 --warmup 4 --no-extras > $O/ncu_diffusion2.log 2>&1
echo "ncu rc=$?" >> $S
ncu -i $O/diffusion2.ncu-rep --page raw --csv > $O/diffusion2_raw.csv 2>/dev/null
ncu -i $O/diffusion2.ncu-rep --page source --csv > $O/diffusion2_source.csv 2>/dev/null
rm -f $O/diffusion2.ncu-rep
date >> $S
