#!/bin/bash
# r02v: A/B of the Life kernel's pipe balance (SHF / LOP3 work moved to the FMA pipe as IMAD / IMAD.HI): parity first, then launch times
O=gpurun_out/r02v
mkdir -p $O
S=$O/status.txt
date > $S
for v in default imad imad2 imad3; do
  if [ $v = default ]; then unset SB200_LIB; else export SB200_LIB=$PWD/stencils.jl_b200/lib/libstencils_b200_$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "life or Life or iterate" > $O/pytest_$v.log 2>&1; echo "$v pytest rc=$?" >> $S
  timeout 200 python tools/life_gens_probe.py > $O/probe_$v.log 2>&1; echo "$v probe rc=$?" >> $S
done
date >> $S
