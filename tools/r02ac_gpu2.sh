#!/bin/bash
# r02ac (2 GPUs): packed plans with the conversions fused into the first / last sweep of a call
O=gpurun_out/r02ac
mkdir -p $O
S=$O/status.txt
date > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=29640
run() { name=$1; shift; P=$((P+1)); env "$@" timeout 600 $TR --master-port $P bench.py --gpus 2 ${ARGS} > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $S; }
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu (2 devices) rc=$?" >> $S
timeout 600 $TR --master-port 29639 tests/multigpu_check.py --plan-only > $O/multigpu_check.log 2>&1; echo "multigpu_check rc=$?" >> $S
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > $O/life_n1.json 2> $O/life_n1.err; echo "life n1 rc=$?" >> $S
ARGS="--steps 20 --warmup 5 --no-extras"
run life_default A=1
run life_g256 SB200_PLAN_GHOST=256
timeout 300 python tools/plan_probe.py life > $O/plan_probe_life.log 2>&1; echo "plan probe rc=$?" >> $S
date >> $S
