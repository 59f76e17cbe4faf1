#!/bin/bash
O=gpurun_out/c9; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "exit $?" >> $O/pytest_gpu.txt
tail -8 $O/pytest_gpu.txt
for pool in 4194304 8388608 12582912 16777216 33554432 125000000; do
  KT_POOL=$pool KT_SPP=192 timeout 300 python tools/kernel_times.py final_scene 2>&1 | tail -1
done
KT_SPP=100 timeout 300 python tools/kernel_times.py random_scene 2>&1 | tail -1
KT_POOL=1048576 KT_SPP=100 timeout 300 python tools/kernel_times.py random_scene 2>&1 | tail -1
KT_SPP=64 timeout 300 python tools/kernel_times.py cornell_box obj_cornell 2>&1 | tail -2
