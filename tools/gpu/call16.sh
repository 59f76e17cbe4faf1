#!/bin/bash
for v in q2 q3 q4 q4_mb3 q4_mb5 q4_q25 q4_q50 q4_mq4 q4_mq12 q2 q4; do
  echo "== $v"
  RTGPU_LIB_PATH=$PWD/build_variants/$v.so KT_SPP=192 timeout 300 python tools/kernel_times.py final_scene 2>&1 | tail -1 | cut -d: -f2-
done
