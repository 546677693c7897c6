#!/bin/bash
# r02y: packed -> packed launches: resident CTAs per SM (3 everywhere / 4 up to G = 6 / up to G = 7 / up to G = 8) and the 8-op rule
O=gpurun_out/r02y
mkdir -p $O
S=$O/status.txt
date > $S
for v in default pk3 pk7 pk8; do
  if [ $v = default ]; then unset SB200_LIB; else export SB200_LIB=$PWD/stencils.jl_b200/lib/libstencils_b200_$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "packed or any_generations or eight_generations" > $O/pytest_$v.log 2>&1; echo "$v pytest rc=$?" >> $S
  timeout 300 python tools/life_gens_probe.py > $O/probe_$v.log 2>&1; echo "$v probe rc=$?" >> $S
done
unset SB200_LIB
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
date >> $S
