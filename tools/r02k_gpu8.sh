#!/bin/bash
# r02k (N GPUs of one box, N = $1): the bench exactly as the driver launches it (weak scaling, Life headline + C5 diffusion beside it)
N=${1:-8}
O=gpurun_out/r02k_n$N
mkdir -p $O
S=$O/status.txt
date > $S
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $O/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench n$N rc=$?" >> $S
if [ "$2" = "more" ]; then
  timeout 600 $TR --master-port 29542 bench.py --gpus $N --workload diffusion --steps 100 --no-extras --strong > $O/bench_diffusion_strong.json 2> $O/bench_diffusion_strong.err; echo "diffusion strong rc=$?" >> $S
  SB200_EXCHANGE=nccl timeout 600 $TR --master-port 29543 bench.py --gpus $N --workload diffusion --steps 100 --no-extras > $O/bench_diffusion_nccl.json 2> $O/bench_diffusion_nccl.err; echo "diffusion nccl fallback rc=$?" >> $S
  timeout 600 $TR --master-port 29544 tests/multigpu_check.py --plan-only > $O/multigpu_check.log 2>&1; echo "multigpu_check rc=$?" >> $S
fi
date >> $S
