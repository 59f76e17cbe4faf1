#!/bin/bash
# sort on/off and threshold per scene (library env knobs), default build
O=gpurun_out/sortenv; mkdir -p $O
for e in "RTGPU_SORT_RAYS=1" "RTGPU_SORT_RAYS=0" "RTGPU_SORT_MIN=65536" "RTGPU_SORT_MIN=1048576" "RTGPU_SORT_MIN=4194304"; do
  echo "== $e" >> $O/ab.txt
  env $e KT_SPP=${1:-64} timeout 300 python tools/kernel_times.py final_scene cornell_box random_scene obj_cornell >> $O/ab.txt 2>&1
done
echo "== C1 full (100 spp)" >> $O/ab.txt
KT_SPP=100 timeout 300 python tools/kernel_times.py random_scene >> $O/ab.txt 2>&1
RTGPU_SORT_RAYS=0 KT_SPP=100 timeout 300 python tools/kernel_times.py random_scene >> $O/ab.txt 2>&1
cat $O/ab.txt
