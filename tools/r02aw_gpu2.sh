#!/bin/bash
# r02aw (2 GPUs): the final build (scalar x-adds in stream3d2_kernel, also in its MIRROR instantiation) — whole GPU suite incl. the
# cross-device plan tests, multigpu_check, and the driver's bench command at N = 2 (Life headline + c5_diffusion beside it)
O=gpurun_out/r02aw
mkdir -p $O
S=$O/status.txt
date > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu (2 devices) rc=$?" >> $S
timeout 300 $TR --master-port 29639 tests/multigpu_check.py --plan-only > $O/multigpu_check.log 2>&1; echo "multigpu_check rc=$?" >> $S
timeout 300 $TR --master-port 29641 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" >> $S
date >> $S
