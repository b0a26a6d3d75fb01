#!/bin/bash
# Developer A/B builds: recompiles ONE translation unit with extra -D flags and links it with the other objects of the
# regular build into variants/<name>.so (select with OGS_LIB_PATH).  usage: build_variants.sh <name> <unit.cu> <flags...>
set -e
name=$1; unit=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
B=$ROOT/opengaussian_b200/csrc/build
extra=""
case $unit in preprocess.cu|binning.cu) extra="--fmad=false";; esac
mkdir -p $ROOT/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
  -Xptxas -v $extra "$@" -c $ROOT/opengaussian_b200/csrc/$unit -o $ROOT/variants/$name.o 2> $ROOT/variants/$name.log
objs=$(ls $B/*.o | grep -v "/${unit%.cu}.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/variants/$name.so $ROOT/variants/$name.o $objs -lcudart
grep -E "spill|Used" $ROOT/variants/$name.log | paste - - | sed 's/ptxas info    ://g' | cut -c1-160
