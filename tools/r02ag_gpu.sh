#!/bin/bash
# r02ag: small2d_kernel (radius-1 shapes on L2-resident grids): parity matrix, whole suite, the 1000 x 1000 probe, sanitizer
O=gpurun_out/r02ag
mkdir -p $O
S=$O/status.txt
date > $S
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "small_grid" > $O/pytest_small.log 2>&1; echo "pytest small rc=$?" >> $S
timeout 300 python tools/mean1000_probe.py > $O/mean1000_probe.log 2>&1; echo "probe rc=$?" >> $S
SB200_SMALL2D_MAX_CELLS=0 timeout 300 python tools/mean1000_probe.py > $O/mean1000_probe_stream2d.log 2>&1; echo "probe stream2d rc=$?" >> $S
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "small_grid and float32" > $O/memcheck_small.log 2>&1; echo "memcheck rc=$?" >> $S
date >> $S
