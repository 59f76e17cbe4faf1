#!/bin/bash
O=gpurun_out/c7; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "primitives or ties or medium or trace_batch" > $O/pytest_quick.txt 2>&1; echo "exit $?" >> $O/pytest_quick.txt
tail -3 $O/pytest_quick.txt
bash tools/gpu/ab.sh c7 48 base w32 w64 w64p w32_u1 w32_u3 w32_u4
bash tools/gpu/ncu_ab.sh c7 w32 > $O/ncu_summary.txt 2>&1
