#!/bin/bash
# round 2, GPU call 1: FP64 latency micro, parity of the f32-cull extend, A/B of occupancy / unroll variants against round 1
mkdir -p gpurun_out/c1
O=gpurun_out/c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
timeout 120 tools/micro/fp64_lat > $O/fp64_lat.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
for v in base cur mb5 mb6 u1 u3; do
  echo "== $v" >> $O/ab.txt
  RTGPU_LIB_PATH=$PWD/build_variants/$v.so KT_SPP=48 timeout 300 python tools/kernel_times.py final_scene >> $O/ab.txt 2>&1
done
for v in base cur; do
  echo "== $v" >> $O/ab_scenes.txt
  RTGPU_LIB_PATH=$PWD/build_variants/$v.so timeout 300 python tools/kernel_times.py cornell_box obj_cornell random_scene cornell_smoke wwscene >> $O/ab_scenes.txt 2>&1
  echo "== $v" >> $O/bounces.txt
  RTGPU_LIB_PATH=$PWD/build_variants/$v.so RTGPU_TRACE_BOUNCES=1 KT_SPP=48 timeout 300 python tools/kernel_times.py final_scene 2>&1 | tail -52 >> $O/bounces.txt
done
tail -3 $O/pytest_gpu.txt; cat $O/ab.txt; cat $O/fp64_lat.txt
