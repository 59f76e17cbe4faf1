#!/bin/bash
O=gpurun_out/c10; mkdir -p $O
for v in p2 p3nopf p3pf p3pf_u3; do
  echo "== $v"
  RTGPU_LIB_PATH=$PWD/build_variants/$v.so KT_SPP=192 timeout 300 python tools/kernel_times.py final_scene 2>&1 | tail -1
done
RTGPU_LIB_PATH=$PWD/build_variants/p3pf.so KT_POOL=16252928 KT_SPP=192 timeout 300 python tools/kernel_times.py final_scene 2>&1 | tail -1
RTGPU_LIB_PATH=$PWD/build_variants/p3pf.so KT_SPP=100 timeout 300 python tools/kernel_times.py random_scene 2>&1 | tail -1
RTGPU_LIB_PATH=$PWD/build_variants/p3pf.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sorted or pool or deterministic or radiance" 2>&1 | tail -3
