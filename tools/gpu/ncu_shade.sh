#!/bin/bash
# full ncu capture of the six class kernels (and k_classify) of a steady-state iteration of the timed bench step
O=${1:-gpurun_out/ncu_shade}; mkdir -p $O
PROF="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-workloads"
timeout 300 $PROF > $O/plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_classify" -s 665 -c 7 -o $O/prof_shade -f $PROF > $O/ncu.log 2>&1
tail -n 2 $O/ncu.log; tail -1 $O/plain.log | cut -c1-300
