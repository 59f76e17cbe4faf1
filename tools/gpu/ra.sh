#!/bin/bash
for ra in 1 2 3 4 8; do
  echo "== RTGPU_RUN_AHEAD=$ra"
  RTGPU_RUN_AHEAD=$ra KT_SPP=100 timeout 300 python tools/kernel_times.py random_scene 2>&1 | tail -1 | cut -d: -f2- | cut -c1-150
  RTGPU_RUN_AHEAD=$ra KT_SPP=192 timeout 300 python tools/kernel_times.py final_scene 2>&1 | tail -1 | cut -d: -f2- | cut -c1-150
done
