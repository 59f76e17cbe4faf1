#!/bin/bash
PROF="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --spp 8"
timeout 600 $PROF > gpurun_out/plain3.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 51 -c 2 -o gpurun_out/prof_shade -f $PROF > gpurun_out/ncu_shade.log 2>&1
tail -2 gpurun_out/ncu_shade.log
