#!/bin/bash
# build_many.sh "NAME:-DFLAG=1 -DOTHER=2" ...   -> build_variants/NAME.so, built in parallel (A/B experiments)
mkdir -p build_variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true \
      -Xcompiler -fPIC -shared -Xptxas -v $flags -o build_variants/$name.so raytracer-2022_b200/csrc/rt_kernels.cu > build_variants/$name.log 2>&1;
    echo "$name: $(grep -A3 'k_extendILi0' build_variants/$name.log | grep -E 'Used|stack frame' | tr -s ' ' | tr '\n' ' ' | cut -c1-200)" ) &
done
wait
