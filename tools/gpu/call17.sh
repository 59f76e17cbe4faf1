#!/bin/bash
O=gpurun_out/c17; mkdir -p $O
KT_SPP=192 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 637 -c 1 -o $O/prof_shade1 -f python tools/kernel_times.py final_scene > $O/ncu1.log 2>&1
KT_SPP=192 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 641 -c 1 -o $O/prof_shade5 -f python tools/kernel_times.py final_scene > $O/ncu5.log 2>&1
grep -i "profiling\|k_shade" $O/ncu1.log $O/ncu5.log | head
