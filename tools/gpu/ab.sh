#!/bin/bash
# ab.sh OUTDIR SPP VARIANT...  : kernel_times of final_scene for each build_variants/VARIANT.so
O=gpurun_out/$1; SPP=$2; shift; shift
mkdir -p $O
for v in "$@"; do
  echo "== $v" >> $O/ab.txt
  RTGPU_LIB_PATH=$PWD/build_variants/$v.so KT_SPP=$SPP timeout 300 python tools/kernel_times.py final_scene ${AB_SCENES} >> $O/ab.txt 2>&1
done
cat $O/ab.txt
