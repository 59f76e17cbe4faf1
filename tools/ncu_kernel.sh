#!/bin/bash
# ncu_kernel.sh <kernel regex> <skip> <count> <outname> : one --set full capture inside a short bench run
PROF="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --spp 8"
timeout 600 $PROF > gpurun_out/plain_$4.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -o gpurun_out/$4 -f $PROF > gpurun_out/ncu_$4.log 2>&1
tail -2 gpurun_out/ncu_$4.log
