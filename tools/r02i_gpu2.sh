#!/bin/bash
# r02i (2 GPUs): fused push / pull kernels of the flag form, ghost thickness and overlap for Life, diffusion ghost 4 / 8 and the
# fused mirror store of stream3d2, the N = 2 bench line
O=gpurun_out/r02i
mkdir -p $O
S=$O/status.txt
date > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=29520
run() { name=$1; shift; P=$((P+1)); env "$@" timeout 600 $TR --master-port $P bench.py --gpus 2 ${ARGS} > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $S; }
timeout 600 python -m pytest tests/test_gpu_plan.py -x -q > $O/pytest_plan.log 2>&1; echo "pytest plan rc=$?" >> $S
timeout 600 $TR --master-port 29519 tests/multigpu_check.py --plan-only > $O/multigpu_check.log 2>&1; echo "multigpu_check rc=$?" >> $S
ARGS="--steps 128 --no-extras"
run life_g32 SB200_PLAN_GHOST=32
run life_g32_unfused SB200_PLAN_GHOST=32 SB200_PLAN_FUSED_XFER=0
run life_g64 SB200_PLAN_GHOST=64
run life_g128 SB200_PLAN_GHOST=128
run life_g256 SB200_PLAN_GHOST=256
run life_g128_ov SB200_PLAN_GHOST=128 SB200_OVERLAP=1
ARGS="--workload diffusion --steps 100 --no-extras"
run diff_g4 SB200_PLAN_GHOST=4
run diff_g4_mirror SB200_PLAN_GHOST=4 SB200_D2_MIRROR=1
run diff_g8 SB200_PLAN_GHOST=8
run diff_g8_mirror SB200_PLAN_GHOST=8 SB200_D2_MIRROR=1
run diff_g4_noov SB200_PLAN_GHOST=4 SB200_OVERLAP=0
timeout 300 python bench.py --workload diffusion --steps 100 --no-extras > $O/diff_n1.json 2> $O/diff_n1.err; echo "diff n1 rc=$?" >> $S
timeout 300 python bench.py --steps 1000 --no-extras > $O/life_n1.json 2> $O/life_n1.err; echo "life n1 rc=$?" >> $S
date >> $S
