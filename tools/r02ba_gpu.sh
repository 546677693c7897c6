#!/bin/bash
# r02ba: the library as committed at the end of round 2 (SASS identical to the build of r02au / r02ax): whole GPU suite, smoke, the driver's bench command
O=gpurun_out/r02ba
mkdir -p $O
S=$O/status.txt
date > $S
timeout 300 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver_k20.json 2> $O/bench_driver_k20.err; echo "bench k20 rc=$?" >> $S
date >> $S
