#!/bin/bash
O=gpurun_out/c8; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "exit $?" >> $O/pytest_gpu.txt
tail -15 $O/pytest_gpu.txt
AB_SCENES="random_scene cornell_box obj_cornell" bash tools/gpu/ab.sh c8 48 p2 p2_u3
KT_SPP=192 timeout 300 python tools/kernel_times.py final_scene
