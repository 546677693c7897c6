#!/bin/bash
# Quick A/B of the two-step diffusion kernel: parity first, then timing, then one ncu capture.  usage: gpu_quick.sh <tag>
O=gpurun_out/$1
mkdir -p $O
S=$O/status.txt
date > $S
timeout 120 python tests/sanitize_cases.py > $O/cases.log 2>&1; echo "cases rc=$?" >> $S
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -k "two_steps" -x -q > $O/pytest_two_steps.log 2>&1
echo "pytest two_steps rc=$?" >> $S
for v in 0 1; do
  SB200_DIFFUSION_DOUBLE_STEP=$v timeout 200 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras \
      > $O/bench_diffusion_ds$v.json 2> $O/bench_diffusion_ds$v.err; echo "bench ds$v rc=$?" >> $S
done
SB200_DIFFUSION_DOUBLE_STEP=1 timeout 150 ncu --set full --import-source on --clock-control none -k regex:stream3d2 -c 1 -f -o $O/diffusion2 \
    python bench.py --workload diffusion --steps 4 --warmup 4 --no-extras > $O/ncu_diffusion2.log 2>&1
echo "ncu rc=$?" >> $S
ncu -i $O/diffusion2.ncu-rep --page raw --csv > $O/diffusion2_raw.csv 2>/dev/null
ncu -i $O/diffusion2.ncu-rep --page source --csv > $O/diffusion2_source.csv 2>/dev/null
rm -f $O/diffusion2.ncu-rep
date >> $S
