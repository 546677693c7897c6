#!/bin/bash
# 2-GPU run: N-GPU == 1-GPU bitwise for the slab iterator (incl. two diffusion steps per launch) and the one-shot slab sweeps.
O=gpurun_out/r01m
mkdir -p $O
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tests/multigpu_check.py --quick > $O/multigpu_check.log 2>&1
echo "multigpu_check rc=$?" > $O/status.txt
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload diffusion --steps 100 --warmup 8 --no-extras > $O/bench_diffusion_2gpu.json 2> $O/bench_diffusion_2gpu.err
echo "bench diffusion 2 gpus rc=$?" >> $O/status.txt
