#!/bin/bash
for rows in 0 25 50 100 200; do
  echo "== band rows $rows"
  RTGPU_BAND_ROWS=$rows KT_SPP=192 timeout 300 python tools/kernel_times.py final_scene 2>&1 | tail -1 | cut -d: -f2-
done
echo "== default"; KT_SPP=192 timeout 300 python tools/kernel_times.py final_scene 2>&1 | tail -1 | cut -d: -f2-
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pool or sorted or deterministic or c5" 2>&1 | tail -3
