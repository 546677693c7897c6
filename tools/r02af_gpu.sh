#!/bin/bash
# r02af: ncu --set full of multi_tile2d_kernel (tools/multi_probe.py)
O=gpurun_out/r02af
mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:multi_tile -s 2 -c 1 -f -o $O/multi python tools/multi_probe.py > $O/ncu.log 2>&1; echo "ncu rc=$?" > $O/status.txt
ncu -i $O/multi.ncu-rep --page raw --csv > $O/multi_raw.csv 2>/dev/null
ncu -i $O/multi.ncu-rep --page source --csv > $O/multi_source.csv 2>/dev/null
ncu -i $O/multi.ncu-rep --page details > $O/multi_details.txt 2>/dev/null
rm -f $O/multi.ncu-rep
