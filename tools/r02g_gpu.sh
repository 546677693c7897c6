#!/bin/bash
# r02g: producer back-off through try_wait's suspend-time hint (box3d default 1000 ns; life / stream3d2 as A/B variants), the
# compile-time shifted stream2d instantiation (kernel / circle must be back at their r02e rates)
O=gpurun_out/r02g
mkdir -p $O
S=$O/status.txt
date > $S
LIBDIR=$PWD/stencils.jl_b200/lib
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $S
for wl in window3d kernel circle mean mean_halo; do
  timeout 200 python bench.py --workload $wl --no-extras > $O/bench_$wl.json 2> $O/bench_$wl.err; echo "bench $wl rc=$?" >> $S
done
for v in b3bo4k b3bo0; do
  SB200_LIB=$LIBDIR/libstencils_b200_$v.so timeout 200 python bench.py --workload window3d --no-extras > $O/bench_window3d_$v.json 2> $O/bench_window3d_$v.err; echo "bench $v rc=$?" >> $S
done
timeout 200 python bench.py --steps 1000 --no-extras > $O/bench_life.json 2> $O/bench_life.err; echo "bench life rc=$?" >> $S
SB200_LIB=$LIBDIR/libstencils_b200_lbbo1k.so timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k "life or Life" > $O/pytest_lbbo1k.log 2>&1; echo "pytest lbbo1k rc=$?" >> $S
SB200_LIB=$LIBDIR/libstencils_b200_lbbo1k.so timeout 200 python bench.py --steps 1000 --no-extras > $O/bench_life_lbbo1k.json 2> $O/bench_life_lbbo1k.err; echo "bench lbbo1k rc=$?" >> $S
timeout 200 python bench.py --workload diffusion --steps 100 --no-extras > $O/bench_diffusion.json 2> $O/bench_diffusion.err; echo "bench diffusion rc=$?" >> $S
SB200_LIB=$LIBDIR/libstencils_b200_d2bo1k.so timeout 200 python bench.py --workload diffusion --steps 100 --no-extras > $O/bench_diffusion_d2bo1k.json 2> $O/bench_diffusion_d2bo1k.err; echo "bench d2bo1k rc=$?" >> $S
date >> $S
