#!/bin/bash
O=gpurun_out/c5; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "primitives or ties or medium or trace_batch" > $O/pytest_quick.txt 2>&1; echo "exit $?" >> $O/pytest_quick.txt
tail -5 $O/pytest_quick.txt
bash tools/gpu/ab.sh c5 48 base pool pool_r512 pool_t640 pool_t768 pool_nopair pool_stay25
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "exit $?" >> $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
