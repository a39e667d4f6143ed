#!/bin/bash
# Builds a variant of libshipenv.so with extra nvcc flags into build_variants/<tag>/libshipenv.so
# (compare with AST_SAC_B200_LIB=build_variants/<tag>/libshipenv.so python bench.py ...).
# usage: tools/build_variant.sh <tag> [-DSENV_... flags]
set -e
tag=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/build_variants/$tag
mkdir -p $out
cd $root/ast_sac_b200/csrc
COMMON="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../../include $*"
nvcc $COMMON -fmad=false -c -o $out/kernels_strict.o kernels_strict.cu &
nvcc $COMMON -fmad=true -c -o $out/kernels_fast.o kernels_fast.cu &
nvcc $COMMON -c -o $out/shipenv.o shipenv.cu &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libshipenv.so $out/kernels_strict.o $out/kernels_fast.o $out/shipenv.o
rm -f $out/*.o
echo built $out/libshipenv.so
