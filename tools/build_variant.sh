#!/bin/bash
# Build a variant of libstencils_b200.so with extra -D flags for ONE source file (A/B runs on the GPU box):
#   tools/build_variant.sh p3 stream3d2.cu -DSB200_D2_PRODUCERS=3   ->  stencils.jl_b200/lib/libstencils_b200_p3.so
# Run with SB200_LIB=stencils.jl_b200/lib/libstencils_b200_p3.so. The default library must be built first.
set -e
name=$1; src=$2; shift 2
cd "$(dirname "$0")/../stencils.jl_b200/csrc"
NV=/usr/local/cuda/bin/nvcc
$NV -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -ccbin /usr/bin/g++ "$@" -c $src -o /tmp/variant_$name.o
objs=$(ls build/*.o | grep -v "build/${src%.cu}.o")
$NV -shared -o ../lib/libstencils_b200_$name.so $objs /tmp/variant_$name.o -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fPIC -ldl
echo ../lib/libstencils_b200_$name.so
