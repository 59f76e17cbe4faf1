#!/bin/bash
# full-set capture of k_extend (bounce 1 of the second render) with source, for SASS-level accounting
O=gpurun_out/c13; mkdir -p $O
KT_SPP=48 KT_POOL=40000000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 53 -c 1 -o $O/prof_extend -f python tools/kernel_times.py final_scene > $O/ncu.log 2>&1
tail -3 $O/ncu.log
ls -la $O
