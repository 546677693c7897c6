#!/bin/bash
# r02au: final evidence of the build with the scalar x-adds in stream3d2_kernel: whole GPU suite, smoke, the driver's bench command,
# the C5 diffusion line (default and an UNROLL=4 variant A/B), the launch list of the bench command, one --set full capture of stream3d2_kernel
O=gpurun_out/r02au
mkdir -p $O
S=$O/status.txt
date > $S
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver_k20.json 2> $O/bench_driver_k20.err; echo "bench k20 rc=$?" >> $S
for rep in 1 2; do
  for v in default d2u4; do
    if [ $v = default ]; then unset SB200_LIB; else export SB200_LIB=$PWD/stencils.jl_b200/lib/libstencils_b200_$v.so; fi
    timeout 200 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras > $O/diffusion_${v}_$rep.json 2> $O/diffusion_${v}_$rep.err; echo "$v diffusion $rep rc=$?" >> $S
  done
done
unset SB200_LIB
date >> $S
R=r02au
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"life_|stream|gather_|scatter_|box3d|halo_kernel|plan_|combine" -c 400 --csv --log-file gpurun_out/${R}_launches_life.csv \
    python bench.py --steps 200 --warmup 16 --no-extras > $O/launches_life.log 2>&1; echo "launch list rc=$?" >> $S
timeout 150 ncu --set full --import-source on --clock-control none -k regex:stream3d2 -s 2 -c 1 -f -o gpurun_out/${R}_diffusion \
    python bench.py --workload diffusion --steps 8 --warmup 4 --no-extras > $O/ncu_diffusion.log 2>&1; echo "ncu diffusion rc=$?" >> $S
ncu -i gpurun_out/${R}_diffusion.ncu-rep --page raw --csv > gpurun_out/${R}_diffusion_raw.csv 2>/dev/null; rm -f gpurun_out/${R}_diffusion.ncu-rep
date >> $S
