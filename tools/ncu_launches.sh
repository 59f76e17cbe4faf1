#!/bin/bash
PROF="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --spp ${1:-24}"
timeout 600 $PROF > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_full.csv $PROF > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-300
