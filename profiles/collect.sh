#!/bin/bash
# Run on the GPU box (under gpurun): ncu launch list of the default bench command + one `--set full` capture
# of each dominant kernel. Outputs go to gpurun_out/; summarise here with `python profiles/summarize.py`.
set -x
mkdir -p gpurun_out
R=${1:-r02}
# launch list of the default bench command: only the library's kernels (the synthetic field generator launches hundreds of torch kernels first)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"life_|stream|gather_|scatter_|box3d|halo_kernel|plan_|combine" -c 400 --csv --log-file gpurun_out/${R}_launches_life.csv \
    python bench.py --steps 200 --warmup 16 --no-extras > gpurun_out/${R}_launches_life.log 2>&1
# (the .ncu-rep captures are 5-8 MB each and gpurun brings back at most 64 MiB: export the raw page as CSV on the box, drop the capture)
export_rep() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
ncu --set full --clock-control none --import-source on -k regex:life_bit -s 6 -c 1 -f -o gpurun_out/${R}_life \
    python bench.py --steps 200 --warmup 16 --no-extras > /dev/null 2>&1
export_rep ${R}_life
[ "$ONLY_LIFE" = 1 ] && exit 0
for wl in mean mean_halo kernel kernel_fma circle positional scatter window3d diffusion; do
  ncu --set full --clock-control none --import-source on -k regex:"stream2d|stream3d|scatter_fast|scatter_stream|gather_stream|box3d" -s 3 -c 1 -f -o gpurun_out/${R}_${wl} \
      python bench.py --workload ${wl} --steps 4 --warmup 3 --no-extras > /dev/null 2>&1
  export_rep ${R}_${wl}
done
ls -la gpurun_out
