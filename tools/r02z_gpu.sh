#!/bin/bash
# r02z: ncu --set full of one packed -> packed launch (8 and 6 generations): pipe utilisation and stall reasons
O=gpurun_out/r02z
mkdir -p $O
S=$O/status.txt
date > $S
for g in 8 6; do
  SB200_LIFE_PACKED_GENS=$g timeout 200 ncu --set full --clock-control none --import-source on -k regex:life_bit -s 6 -c 1 -f -o $O/pk$g \
      python bench.py --steps 200 --warmup 16 --no-extras > $O/ncu_pk$g.log 2>&1; echo "ncu pk$g rc=$?" >> $S
  ncu -i $O/pk$g.ncu-rep --page raw --csv > $O/pk${g}_raw.csv 2>/dev/null
  ncu -i $O/pk$g.ncu-rep --page source --csv > $O/pk${g}_source.csv 2>/dev/null
  rm -f $O/pk$g.ncu-rep
done
date >> $S
