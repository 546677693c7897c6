#!/bin/bash
# r02d (2 GPUs): the slab plans across devices (single-process form in pytest, rank form under torchrun), the scaling bench
# with whole exchange cycles in the timed region, fused ghost store in stream3d2 on / off
O=gpurun_out/r02d
mkdir -p $O
S=$O/status.txt
date > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_plan.py -x -q > $O/pytest_plan.log 2>&1; echo "pytest plan rc=$?" >> $S
timeout 600 $TR --master-port 29511 tests/multigpu_check.py --quick > $O/multigpu_check.log 2>&1; echo "multigpu_check rc=$?" >> $S
SB200_D2_MIRROR=1 timeout 600 $TR --master-port 29512 tests/multigpu_check.py --plan-only > $O/multigpu_check_mirror.log 2>&1; echo "multigpu_check mirror rc=$?" >> $S
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?" >> $S
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" >> $S
SB200_D2_MIRROR=1 timeout 600 $TR --master-port 29514 bench.py --gpus 2 --workload diffusion --steps 100 --no-extras > $O/bench_n2_diffusion_mirror.json 2> $O/bench_n2_diffusion_mirror.err; echo "bench n2 diffusion mirror rc=$?" >> $S
timeout 600 $TR --master-port 29515 bench.py --gpus 2 --workload diffusion --steps 100 --no-extras > $O/bench_n2_diffusion.json 2> $O/bench_n2_diffusion.err; echo "bench n2 diffusion rc=$?" >> $S
SB200_OVERLAP=0 timeout 600 $TR --master-port 29516 bench.py --gpus 2 --workload diffusion --steps 100 --no-extras > $O/bench_n2_diffusion_noov.json 2> $O/bench_n2_diffusion_noov.err; echo "bench n2 diffusion no overlap rc=$?" >> $S
timeout 600 $TR --master-port 29517 bench.py --gpus 2 --workload diffusion --steps 100 --no-extras --strong > $O/bench_n2_diffusion_strong.json 2> $O/bench_n2_diffusion_strong.err; echo "bench n2 diffusion strong rc=$?" >> $S
timeout 300 $TR --master-port 29518 bench.py --gpus 2 --impl reference --steps 20 --warmup 5 > $O/bench_n2_ref.json 2> $O/bench_n2_ref.err; echo "bench n2 ref rc=$?" >> $S
date >> $S
