#!/bin/bash
# build_variant.sh NAME [-DMACRO=..]...  ->  build_variants/NAME.so  (A/B experiments; select with RTGPU_LIB_PATH)
# SRC=<dir> builds from another source tree (e.g. a `git worktree` of an older commit)
name=$1; shift
mkdir -p build_variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true \
  -Xcompiler -fPIC -shared -Xptxas -v "$@" -o build_variants/$name.so ${SRC:-.}/raytracer-2022_b200/csrc/rt_kernels.cu 2>&1 | grep -A2 "k_extendE" | grep -E "Used|stack"
