#!/bin/bash
O=gpurun_out/c12; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_multi.py -m gpu -x -q --durations=12 > $O/pytest_full.txt 2>&1; echo "exit $?" >> $O/pytest_full.txt
tail -25 $O/pytest_full.txt
timeout 900 python tools/parity_report.py $O/parity_r02.json > $O/parity_report.txt 2>&1; tail -8 $O/parity_report.txt
