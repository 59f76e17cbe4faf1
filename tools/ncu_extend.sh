#!/bin/bash
# one ncu --set full capture of k_extend (bounces 1..3 of the timed step) + the launch list
PROF="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --spp 8"
timeout 600 $PROF > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1
timeout 600 $PROF > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 51 -c 2 -o gpurun_out/prof_extend -f $PROF > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
