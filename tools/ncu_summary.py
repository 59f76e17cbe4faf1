"""Summarise gpurun_out/{launches_full.csv, prof_extend.ncu-rep} into profiles/ (tracked)."""
import csv, json, os, re, subprocess, sys
from collections import defaultdict

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = "profiles"
os.makedirs(out_dir, exist_ok=True)
G = "gpurun_out"

# 1. launch list: per-kernel device time, warp instructions and active lanes per instruction
rows = [r for r in csv.reader(open(f"{G}/launches_full.csv")) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
d = defaultdict(dict)
for r in rows[1:]:
    try:
        d[int(r[ii])][r[mi]] = float(r[vi].replace(",", ""))
    except ValueError:
        pass
    d[int(r[ii])]["name"] = r[ki]
ids = sorted(d)
half = len(ids) // 2  # first half = warm-up step, second half = the timed step
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
seq = []
for i in ids[half:]:
    x = d[i]
    m = re.search(r"k_shade<(\d)>", x["name"])
    key = f"k_shade<{m.group(1)}>" if m else x["name"].split("(")[0].replace("<unnamed>::", "")[:60]
    a = agg[key]
    a[0] += 1
    a[1] += x.get("gpu__time_duration.sum", 0)
    a[2] += x.get("smsp__inst_executed.sum", 0)
    a[3] += x.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0) * x.get("smsp__inst_executed.sum", 0)
    seq.append((key, x.get("gpu__time_duration.sum", 0)))
tot = sum(a[1] for a in agg.values())
with open(f"{out_dir}/{tag}_launches.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none\n")
    f.write("# command: python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --spp 24   (C3 final_scene 800x800, 24 spp = one wave)\n")
    f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.  Timed step only.\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:44s} launches {a[0]:4d}  time {a[1] / 1e6:8.2f} ms  share {a[1] / tot:.3f}  warp-instr {a[2] / 1e6:9.1f} M  "
                f"active lanes/instr {a[3] / max(a[2], 1):5.1f}\n")
    f.write("\n# first 24 launches of the timed step (kernel, us)\n")
    for n, v in seq[:24]:
        f.write(f"{n:40s} {v / 1e3:10.1f}\n")
print(open(f"{out_dir}/{tag}_launches.txt").read())
ext_share = agg["k_extend"][1] / tot

# 2. full-set metrics of k_extend
raw = subprocess.run(["ncu", "-i", f"{G}/prof_extend.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
with open(f"{out_dir}/{tag}_k_extend_ncu.txt", "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:k_extend -s 51 -c 2\n")
    f.write("# = bounces 1 and 2 of the timed step (54 M and 34 M incoherent rays), same bench command as the launch list\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            f.write(f"{w:85s} {units[i]:16s} {[r[i] for r in rows[2:]]}\n")
print(open(f"{out_dir}/{tag}_k_extend_ncu.txt").read())
i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tr = [float(r[i_r]) * scale[units[i_r]] + float(r[i_w]) * scale[units[i_w]] for r in rows[2:]]
json.dump({"k_extend_dram_bytes_per_launch": tr[0], "captured_launch": "bounce 1 of a 24-spp wave (54.1 M rays)",
           "all_captured_launches": tr, "k_extend_share_of_step_under_ncu": ext_share,
           "source": f"profiles/{tag}_k_extend_ncu.txt, profiles/{tag}_launches.txt"}, open(f"{out_dir}/roofline_ncu.json", "w"), indent=1)
