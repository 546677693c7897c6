#!/bin/bash
# r02w: life_bit_kernel with the funnel-shift pack, no edge-byte reads, predicated stores: whole GPU suite + launch times + sanitizer
O=gpurun_out/r02w
mkdir -p $O
S=$O/status.txt
date > $S
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 200 python tools/life_gens_probe.py > $O/probe.log 2>&1; echo "probe rc=$?" >> $S
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_cases.py --quick > $O/memcheck_quick.log 2>&1; echo "memcheck rc=$?" >> $S
date >> $S
