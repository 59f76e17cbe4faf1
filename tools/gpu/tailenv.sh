#!/bin/bash
O=gpurun_out/tailenv; mkdir -p $O
for e in "RTGPU_TAIL_MAX=65536" "RTGPU_TAIL_MAX=131072" "RTGPU_TAIL_MAX=262144" "RTGPU_TAIL_MAX=524288" "RTGPU_TAIL_MAX=1048576" "RTGPU_RUN_AHEAD=1" "RTGPU_RUN_AHEAD=3" "RTGPU_TAIL_MAX=65536"; do
  echo "== $e" >> $O/ab.txt
  env $e KT_SPP=192 timeout 300 python tools/kernel_times.py final_scene >> $O/ab.txt 2>&1
  env $e KT_SPP=100 timeout 300 python tools/kernel_times.py random_scene >> $O/ab.txt 2>&1
done
cat $O/ab.txt
