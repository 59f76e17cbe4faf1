#!/bin/bash
# RTGPU_OVERLAP on / off, same build (side streams: k_regen beside the sort, the class kernels beside each other)
O=gpurun_out/overlap; mkdir -p $O; rm -f $O/ab.txt
for e in "RTGPU_OVERLAP=0" "RTGPU_OVERLAP=1" "RTGPU_OVERLAP=0" "RTGPU_OVERLAP=1"; do
  echo "== $e" >> $O/ab.txt
  env $e KT_SPP=192 timeout 300 python tools/kernel_times.py final_scene >> $O/ab.txt 2>&1
  env $e KT_SPP=64 timeout 300 python tools/kernel_times.py cornell_box random_scene obj_cornell >> $O/ab.txt 2>&1
done
cat $O/ab.txt
