#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full captures of k_extend (bounces 1, 2 and 20 of the timed step)
# and k_shade<1>.  Every ncu pass runs only after the same command has exited 0 without ncu (its JSON line is kept:
# rays per bounce of the captured launches come from there).
PROF="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
mkdir -p gpurun_out
timeout 300 $PROF > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -c 1700 --csv --log-file gpurun_out/launches_full.csv $PROF > gpurun_out/ncu_launches.log 2>&1
timeout 300 $PROF > gpurun_out/plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 51 -c 2 -o gpurun_out/prof_extend -f $PROF > gpurun_out/ncu_full.log 2>&1
timeout 300 $PROF > gpurun_out/plain3.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 70 -c 1 -o gpurun_out/prof_extend_deep -f $PROF > gpurun_out/ncu_deep.log 2>&1
timeout 300 $PROF > gpurun_out/plain4.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:k_shade<1>" -s 51 -c 1 -o gpurun_out/prof_shade1 -f $PROF > gpurun_out/ncu_shade1.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
