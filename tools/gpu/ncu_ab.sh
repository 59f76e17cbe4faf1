#!/bin/bash
# ncu_ab.sh OUTDIR VARIANT...  : selected counters of k_extend (bounces 0,1,2 and 20 of the second 48-spp render of final_scene)
O=gpurun_out/$1; shift
mkdir -p $O
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,\
l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,\
l1tex__data_pipe_lsu_wavefronts.sum,sass__inst_executed_local_loads,sass__inst_executed_local_stores,sass__inst_executed_shared_loads,sass__inst_executed_shared_stores,\
sass__inst_executed_global_loads,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,\
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,\
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,\
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,\
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,\
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,\
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,\
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed_op_shared_atom.sum,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
for v in "$@"; do
  for sk in "50 3" "70 1"; do
    set -- $sk
    RTGPU_LIB_PATH=$PWD/build_variants/$v.so KT_SPP=48 timeout 600 ncu --metrics $M --clock-control none -k regex:k_extend -s $1 -c $2 --csv --log-file $O/${v}_s$1.csv \
      python tools/kernel_times.py final_scene > $O/${v}_s$1.log 2>&1
  done
done
python - "$O" <<'PY'
import csv, sys, glob, os
from collections import defaultdict
O = sys.argv[1]
for f in sorted(glob.glob(O + "/*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    if not rows: continue
    h = rows[0]; mi, vi, ii = h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    d = defaultdict(dict)
    for r in rows[1:]:
        d[r[ii]][r[mi]] = r[vi]
    print("==", os.path.basename(f))
    ids = sorted(d, key=int)
    for m in sorted(d[ids[0]]):
        print(f"{m:90s}", "  ".join(f"{d[i].get(m, ''):>16s}" for i in ids))
PY
