#!/bin/bash
# Run on an 8-GPU box (gpurun --gpus 8): weak scaling of the two iterated configs at N = 1, 2, 4, 8, the way the
# driver launches bench.py. One JSON line per (workload, N) in gpurun_out/scale_<round>_<workload>_<N>.json.
R=${1:-r01}
mkdir -p gpurun_out
port=29600
for wl in ${WORKLOADS:-life diffusion}; do
  steps=1000; [ $wl = diffusion ] && steps=100
  for n in ${NGPUS:-1 2 4 8}; do
    port=$((port+1))
    if [ $n = 1 ]; then
      python bench.py --gpus 1 --workload $wl --steps $steps --warmup 16 --no-extras 2>/dev/null | tail -1 > gpurun_out/scale_${R}_${wl}_$n.json
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n --workload $wl --steps $steps --warmup 16 2>/dev/null | tail -1 > gpurun_out/scale_${R}_${wl}_$n.json
    fi
    python -c "import json;d=json.load(open('gpurun_out/scale_${R}_${wl}_$n.json'));print('$wl',d['n_gpus'],round(d['value'],1),round(d['ms_per_step'],4),d['config'].get('exchange','')[:30])"
  done
done
