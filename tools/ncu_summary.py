"""Summarise gpurun_out/{launches.csv, prof_extend.ncu-rep} into profiles/ (tracked)."""
import csv, json, os, subprocess, sys
from collections import defaultdict

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = "profiles"
os.makedirs(out_dir, exist_ok=True)
G = "gpurun_out"

# 1. launch list
rows = [r for r in csv.reader(open(f"{G}/launches.csv")) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt, seq = defaultdict(float), defaultdict(int), []
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0].replace("<unnamed>::", "")
    tot[name] += v
    cnt[name] += 1
    seq.append((name, v))
s = sum(tot.values())
with open(f"{out_dir}/{tag}_launches.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)\n")
    f.write("# command: python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --spp 8   (C3 final_scene 800x800)\n")
    for k in sorted(tot, key=lambda k: -tot[k]):
        f.write(f"{k:70s} launches {cnt[k]:4d}  total {tot[k] / 1e6:9.3f} ms  share {tot[k] / s:.3f}\n")
    f.write("\n# first 40 launches (name, us)\n")
    for n, v in seq[:40]:
        f.write(f"{n:40s} {v / 1e3:10.1f}\n")
print(open(f"{out_dir}/{tag}_launches.txt").read()[:1500])

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
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
with open(f"{out_dir}/{tag}_k_extend_ncu.txt", "w") as f:
    f.write("# ncu --set full --clock-control none -k regex:k_extend (consecutive launches of one wave; bounce 1, 2 of the timed step)\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            f.write(f"{w:85s} {units[i]:16s} {[r[i] for r in rows[2:]]}\n")
print(open(f"{out_dir}/{tag}_k_extend_ncu.txt").read())
try:
    i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tr = [float(r[i_r]) * scale[units[i_r]] + float(r[i_w]) * scale[units[i_w]] for r in rows[2:]]
    json.dump({"k_extend_dram_bytes_per_launch": tr[0], "all_captured_launches": tr, "source": f"profiles/{tag}_k_extend_ncu.txt"},
              open(f"{out_dir}/roofline_ncu.json", "w"))
except Exception as e:
    print("no traffic", e)
