#!/bin/bash
O=gpurun_out/tail; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tail or pool or sorted or deterministic" 2>&1 | tail -4
for tm in 0 16384 65536 262144 1048576; do
  echo "== RTGPU_TAIL_MAX=$tm"
  RTGPU_TAIL_MAX=$tm KT_SPP=100 timeout 300 python tools/kernel_times.py random_scene 2>&1 | tail -1 | cut -d: -f2- | cut -c1-150
  RTGPU_TAIL_MAX=$tm KT_SPP=192 timeout 300 python tools/kernel_times.py final_scene 2>&1 | tail -1 | cut -d: -f2- | cut -c1-150
done
RTGPU_TAIL_MAX=65536 KT_SPP=64 timeout 300 python tools/kernel_times.py cornell_box obj_cornell wwscene 2>&1 | tail -3 | cut -c1-200
echo "== compute-sanitizer memcheck on the smoke render"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer.txt 2>&1; echo "sanitizer exit $?"; tail -4 $O/sanitizer.txt
