#!/bin/bash
# r02ab (2 GPUs): packed Life runs in the slab plans across devices: plan tests, multigpu_check, ghost thickness, the N = 2 bench line
O=gpurun_out/r02ab
mkdir -p $O
S=$O/status.txt
date > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=29620
run() { name=$1; shift; P=$((P+1)); env "$@" timeout 600 $TR --master-port $P bench.py --gpus 2 ${ARGS} > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $S; }
timeout 600 python -m pytest tests/test_gpu_plan.py -x -q > $O/pytest_plan.log 2>&1; echo "pytest plan rc=$?" >> $S
timeout 600 $TR --master-port 29619 tests/multigpu_check.py --plan-only > $O/multigpu_check.log 2>&1; echo "multigpu_check rc=$?" >> $S
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > $O/life_n1.json 2> $O/life_n1.err; echo "life n1 rc=$?" >> $S
ARGS="--steps 20 --warmup 5 --no-extras"
run life_default A=1
run life_g64 SB200_PLAN_GHOST=64
run life_g256 SB200_PLAN_GHOST=256
run life_unpacked SB200_LIFE_PACKED=0
ARGS="--steps 20 --warmup 5"
run bench_n2 A=1
date >> $S
