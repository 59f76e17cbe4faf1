#!/bin/bash
# ncu evidence for profiles/ (round 2): launch list of one bench step + full captures of k_extend (an early and a steady-state
# iteration of the timed step) and of k_shade<1>.  Every ncu pass runs only after the same command has exited 0 without ncu.
# usage: tools/ncu_round.sh [outdir]   (default gpurun_out/ncu_r02)
O=${1:-gpurun_out/ncu_r02}; mkdir -p $O
PROF="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-workloads"
timeout 300 $PROF > $O/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -c 6000 --csv --log-file $O/launches_full.csv $PROF > $O/ncu_launches.log 2>&1
# the timed step is the second render: its k_extend launches are numbers 80.. (80 iterations per 192-spp step at the default pool of 15 Mi paths)
timeout 300 $PROF > $O/plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 81 -c 2 -o $O/prof_extend -f $PROF > $O/ncu_full.log 2>&1
timeout 300 $PROF > $O/plain3.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 95 -c 1 -o $O/prof_extend_steady -f $PROF > $O/ncu_steady.log 2>&1
tail -n 2 $O/ncu_full.log; tail -1 $O/plain.log | cut -c1-300
