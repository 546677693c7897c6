#!/bin/bash
# r02am (final build, second run: packed Life plans): the whole weak-scaling series N = 1, 2, 4, 8 on ONE 8-GPU box (same silicon, same thermal state history as the driver's SCALE run)
O=gpurun_out/r02am
mkdir -p $O
S=$O/status.txt
date > $S
nvidia-smi --query-gpu=index,name,power.limit,clocks.max.sm --format=csv > $O/gpus.txt 2>&1
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > $O/life_n1.json 2> $O/life_n1.err; echo "life n1 rc=$?" >> $S
timeout 300 python bench.py --gpus 1 --workload diffusion --steps 100 --no-extras > $O/diffusion_n1.json 2> $O/diffusion_n1.err; echo "diffusion n1 rc=$?" >> $S
P=29550
for N in 2 4 8; do
  P=$((P+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench n$N rc=$?" >> $S
done
timeout 300 python bench.py --gpus 1 --workload diffusion --steps 100 --no-extras > $O/diffusion_n1_after.json 2> $O/diffusion_n1_after.err; echo "diffusion n1 (after) rc=$?" >> $S
P=$((P+1))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --workload diffusion --steps 100 --no-extras --strong > $O/diffusion_strong_n8.json 2> $O/diffusion_strong_n8.err; echo "diffusion strong n8 rc=$?" >> $S
date >> $S
