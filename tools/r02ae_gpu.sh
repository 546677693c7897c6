#!/bin/bash
# r02ae: single-pass multi-array gathers (multi_tile2d_kernel): parity, A/B against the sweep-per-argument path, sanitizer
O=gpurun_out/r02ae
mkdir -p $O
S=$O/status.txt
date > $S
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -q > $O/pytest_api.log 2>&1; echo "pytest api rc=$?" >> $S
timeout 300 python tools/multi_probe.py > $O/multi_probe.log 2>&1; echo "probe rc=$?" >> $S
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_api.py -m gpu -q -k "multi or layered" > $O/memcheck_multi.log 2>&1; echo "memcheck rc=$?" >> $S
date >> $S
