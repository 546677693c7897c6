#!/bin/bash
# r02c: the C-ABI slab plans on one GPU (several slabs per device) + the whole suite with the new defaults
O=gpurun_out/r02c
mkdir -p $O
S=$O/status.txt
date > $S
timeout 600 python -m pytest tests/test_gpu_plan.py -x -q > $O/pytest_plan.log 2>&1; echo "pytest plan rc=$?" >> $S
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 200 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras > $O/bench_diffusion.json 2> $O/bench_diffusion.err; echo "bench diffusion rc=$?" >> $S
date >> $S
