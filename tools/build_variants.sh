#!/bin/bash
# Build variant libraries of the CUDA path for an A/B on the GPU box (run HERE, before gpurun: build/ travels with
# the snapshot).  Each argument is "name:flags", e.g.
#   tools/build_variants.sh base: xwants:-DWB_CL_XWANTS=1 "xr:-DWB_CL_XWANTS=1 -DWB_CL_REFILTER=1"
# -> build/variants/lib_<name>.so; select one on the box with WB_LIB=$PWD/build/variants/lib_<name>.so.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for spec in "$@"; do
  name=${spec%%:*}
  flags=${spec#*:}
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -shared \
      $flags -Xptxas -v -o build/variants/lib_$name.so wolkenbase_b200/csrc/wolken_b200.cu 2>&1 \
      | grep -A2 "wb_classify_kernelILi1" | grep "spill\|registers" | tr '\n' ' '; echo " <- $name ($flags)" ) &
done
wait
ls -la build/variants
