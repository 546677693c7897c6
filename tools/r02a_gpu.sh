#!/bin/bash
# r02a: first GPU call of round 2 — the experiments prepared at the end of round 1, A/B against the default build.
#   variants built here first with tools/build_variant.sh: hl1 (life one halo lane), s2p2 / s2p4 (cp.async producers),
#   g3p2 (gather_stream3d producers), pk (packed f32x2 folds in stream3d2)
O=gpurun_out/r02a
mkdir -p $O
S=$O/status.txt
date > $S
LIBDIR=$PWD/stencils.jl_b200/lib
# ---- default build: whole suite (sanity of the restored tree) ----
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $S
timeout 300 python bench.py --no-extras --steps 1000 > $O/bench_life_default.json 2> $O/bench_life_default.err; echo "bench life default rc=$?" >> $S
# ---- Life: one halo lane, then eight generations per launch ----
L=$LIBDIR/libstencils_b200_hl1.so
SB200_LIB=$L timeout 300 python -m pytest tests -m gpu -x -q -k "life or Life" > $O/pytest_hl1.log 2>&1; echo "pytest hl1 rc=$?" >> $S
SB200_LIB=$L SB200_EXPERIMENTS=1 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "eight_generations" > $O/pytest_hl1_oct.log 2>&1; echo "pytest hl1 oct rc=$?" >> $S
SB200_LIB=$L timeout 300 python bench.py --no-extras --steps 1000 > $O/bench_life_hl1.json 2> $O/bench_life_hl1.err; echo "bench hl1 rc=$?" >> $S
for t in 1 3 4; do
  SB200_LIB=$L SB200_LB_TASKS=$t timeout 300 python bench.py --no-extras --steps 1000 > $O/bench_life_hl1_t$t.json 2> $O/bench_life_hl1_t$t.err; echo "bench hl1 tasks=$t rc=$?" >> $S
done
SB200_LIB=$L SB200_OCT_STEP=1 timeout 300 python bench.py --no-extras --steps 1000 > $O/bench_life_hl1_oct.json 2> $O/bench_life_hl1_oct.err; echo "bench hl1 oct rc=$?" >> $S
# ---- diffusion: Remove axes experiment, packed folds ----
SB200_EXPERIMENTS=1 SB200_D2_REMOVE=1 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "remove_axes" > $O/pytest_d2_remove.log 2>&1; echo "pytest d2 remove rc=$?" >> $S
timeout 200 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras > $O/bench_diffusion_default.json 2> $O/bench_diffusion_default.err; echo "bench diffusion rc=$?" >> $S
L=$LIBDIR/libstencils_b200_pk.so
SB200_LIB=$L timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -k "two_steps" -x -q > $O/pytest_pk.log 2>&1; echo "pytest pk rc=$?" >> $S
SB200_LIB=$L timeout 200 python bench.py --workload diffusion --steps 100 --warmup 4 --no-extras > $O/bench_diffusion_pk.json 2> $O/bench_diffusion_pk.err; echo "bench diffusion pk rc=$?" >> $S
# ---- Halo-padded mean: cp.async producers ----
timeout 200 python bench.py --workload mean_halo --no-extras > $O/bench_mean_halo_default.json 2> $O/bench_mean_halo_default.err
for v in s2p2 s2p4; do
  L=$LIBDIR/libstencils_b200_$v.so
  SB200_LIB=$L timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "halo or Halo or reducers_2d" > $O/pytest_$v.log 2>&1; echo "pytest $v rc=$?" >> $S
  SB200_LIB=$L timeout 200 python bench.py --workload mean_halo --no-extras > $O/bench_mean_halo_$v.json 2> $O/bench_mean_halo_$v.err; echo "bench $v rc=$?" >> $S
done
# ---- Window(1,3): more producers ----
timeout 200 python bench.py --workload window3d --no-extras > $O/bench_window3d_default.json 2> $O/bench_window3d_default.err
L=$LIBDIR/libstencils_b200_g3p2.so
SB200_LIB=$L timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "stream3d" > $O/pytest_g3p2.log 2>&1; echo "pytest g3p2 rc=$?" >> $S
SB200_LIB=$L timeout 200 python bench.py --workload window3d --no-extras > $O/bench_window3d_g3p2.json 2> $O/bench_window3d_g3p2.err; echo "bench g3p2 rc=$?" >> $S
date >> $S
