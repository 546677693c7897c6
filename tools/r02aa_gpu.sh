#!/bin/bash
# r02aa: packed -> packed launches with the one-bit shifts on the FMA pipe (IMAD / IMAD.HI) against the funnel shifts
O=gpurun_out/r02aa
mkdir -p $O
S=$O/status.txt
date > $S
for v in default ims; do
  if [ $v = default ]; then unset SB200_LIB; else export SB200_LIB=$PWD/stencils.jl_b200/lib/libstencils_b200_$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "packed" > $O/pytest_$v.log 2>&1; echo "$v pytest rc=$?" >> $S
  timeout 300 python tools/life_gens_probe.py > $O/probe_$v.log 2>&1; echo "$v probe rc=$?" >> $S
done
date >> $S
