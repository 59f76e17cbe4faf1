#!/bin/bash
O=gpurun_out/c14; mkdir -p $O
for v in p3pf_u3 q1 q2; do
  echo "== $v"
  RTGPU_LIB_PATH=$PWD/build_variants/$v.so KT_SPP=192 timeout 300 python tools/kernel_times.py final_scene 2>&1 | tail -1
done
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "exit $?" >> $O/pytest_gpu.txt
tail -4 $O/pytest_gpu.txt
