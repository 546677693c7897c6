#!/bin/bash
# r02ar: packed Life kernels with the output row as a predicated store (no branch around the store block), A/B
O=gpurun_out/r02ar
mkdir -p $O
S=$O/status.txt
date > $S
for v in default pred; do
  if [ $v = default ]; then unset SB200_LIB; else export SB200_LIB=$PWD/stencils.jl_b200/lib/libstencils_b200_$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plan.py -m gpu -q -k "packed or life or Life or plan" > $O/pytest_$v.log 2>&1; echo "$v pytest rc=$?" >> $S
  timeout 300 python tools/life_gens_probe.py > $O/probe_$v.log 2>&1; echo "$v probe rc=$?" >> $S
done
date >> $S
