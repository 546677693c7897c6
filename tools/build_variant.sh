#!/bin/bash
# Build a variant of libstencils_b200.so with extra -D flags for some source files (A/B runs on the GPU box):
#   tools/build_variant.sh p3 stream3d2.cu -DSB200_D2_PRODUCERS=3           ->  stencils.jl_b200/lib/libstencils_b200_p3.so
#   tools/build_variant.sh s2p2 stream2d_a.cu,stream2d_b.cu,stream2d_c.cu,stream2d_d.cu -DSB200_S2_PRODUCERS=2
# Run with SB200_LIB=$PWD/stencils.jl_b200/lib/libstencils_b200_<name>.so. The default library must be built first
# (the other objects are taken from csrc/build/).
set -e
name=$1; srcs=${2//,/ }; shift 2
cd "$(dirname "$0")/../stencils.jl_b200/csrc"
NV=/usr/local/cuda/bin/nvcc
objs=$(ls build/*.o)
new=""
for src in $srcs; do
  o=/tmp/variant_${name}_${src%.cu}.o
  $NV -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -ccbin /usr/bin/g++ "$@" -c $src -o $o
  objs=$(echo "$objs" | grep -v "build/${src%.cu}.o")
  new="$new $o"
done
$NV -shared -o ../lib/libstencils_b200_$name.so $objs $new -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fPIC -ldl
echo ../lib/libstencils_b200_$name.so
