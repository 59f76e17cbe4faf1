#!/bin/bash
O=gpurun_out/c3; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "primitives or ties or medium or trace_batch" > $O/pytest_quick.txt 2>&1; echo "exit $?" >> $O/pytest_quick.txt
tail -5 $O/pytest_quick.txt
bash tools/gpu/ab.sh c3 48 base svc svc_u1 svc_u3 svc_nopair svc_r8
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "exit $?" >> $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
