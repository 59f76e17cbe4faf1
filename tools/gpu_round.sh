#!/bin/bash
# One GPU-box session: tests, bench (both arms), ncu launch list + one full capture of k_extend.
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_ref.json
PROF="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --spp 4"
timeout 600 $PROF > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1
timeout 600 $PROF > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 2 -c 3 -o gpurun_out/prof_extend -f $PROF > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
