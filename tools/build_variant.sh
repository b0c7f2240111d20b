#!/usr/bin/env bash
# tuning aid: build libgsb with different -D settings into build_variants/<name>.so (select at run time with GSB_LIB=...)
# usage: tools/build_variant.sh name "-DGSB_SBW=8 -DGSB_SBH=4"
set -e
name=$1; defs=$2
cd "$(dirname "$0")/.."
out=build_variants/$name; mkdir -p $out
csrc=gaussiansplattingmlx_b200/csrc
common="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr $defs"
for f in project adam densify; do nvcc $common --fmad=false -c $csrc/$f.cu -o $out/$f.o & done
for f in binning tilelists raster loss api project_bwd; do nvcc $common -c $csrc/$f.cu -o $out/$f.o & done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o build_variants/$name.so $out/*.o
rm -rf $out
echo built build_variants/$name.so
