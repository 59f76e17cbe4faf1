"""Summarise <dir>/{launches_full.csv, prof_extend*.ncu-rep} (tools/ncu_round.sh) into profiles/ (tracked).

    python tools/ncu_summary.py [TAG] [DIR]     default: r02 gpurun_out/ncu_r02
"""
import csv, json, os, re, subprocess, sys
from collections import defaultdict

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/ncu_r02"
out_dir = "profiles"
os.makedirs(out_dir, exist_ok=True)


def bench_line(path):
    for ln in open(path):
        if ln.startswith("{"):
            return json.loads(ln)
    return {}


plain = bench_line(f"{G}/plain.log")
pool = int(plain.get("run", {}).get("paths_in_flight_per_gpu", 0))

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
# a render ends with k_finish: the timed step is what lies between the last two k_finish launches
fin = [k for k, i in enumerate(ids) if "k_finish" in d[i]["name"]]
lo, hi = (fin[-2] + 1, fin[-1] + 1) if len(fin) >= 2 else (0, len(ids))
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
seq = []
for i in ids[lo:hi]:
    x = d[i]
    m = re.search(r"k_shade<(\d)>", x["name"])
    key = f"k_shade<{m.group(1)}>" if m else x["name"].split("(")[0].replace("<unnamed>::", "")[:60]
    key = re.sub(r"^(?:void )?(k_extend|k_tail|k_trace_extend)<[^>]*>$", r"\1", key)  # NEST / MESH instantiations
    a = agg[key]
    a[0] += 1
    a[1] += x.get("gpu__time_duration.sum", 0)
    a[2] += x.get("smsp__inst_executed.sum", 0)
    a[3] += x.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0) * x.get("smsp__inst_executed.sum", 0)
    seq.append((key, x.get("gpu__time_duration.sum", 0)))
tot = sum(a[1] for a in agg.values())
with open(f"{out_dir}/{tag}_launches.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none\n")
    f.write("# command: python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-workloads   (C3 final_scene 800x800, 192 spp per step,\n")
    f.write(f"#          pool of {pool} paths: see the bench line next to this file)\n")
    f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.  Timed step only (between two k_finish).\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:44s} launches {a[0]:4d}  time {a[1] / 1e6:8.2f} ms  share {a[1] / tot:.3f}  warp-instr {a[2] / 1e6:9.1f} M  "
                f"active lanes/instr {a[3] / max(a[2], 1):5.1f}\n")
    lib = sum(a[1] for k, a in agg.items() if "cub" in k or "Radix" in k)
    f.write(f"\n# library (cub radix sort) share of the step: {lib / tot:.3f}\n")
    f.write("\n# launches of the second iteration of the timed step (kernel, us)\n")
    starts = [k for k, (n, _) in enumerate(seq) if n == "k_iter_begin"]
    if len(starts) >= 3:
        for n, v in seq[starts[1]:starts[2]]:
            f.write(f"{n:60s} {v / 1e3:10.1f}\n")
print(open(f"{out_dir}/{tag}_launches.txt").read())
ext_share = agg["k_extend"][1] / tot


# 2. full-set metrics of k_extend
def raw_rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    return rows[0], rows[1], rows[2:]


want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__icc_request_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
hdr, units, body = raw_rows(f"{G}/prof_extend.ncu-rep")
cols = [(f"iteration {b} of the step ({pool / 1e6:.1f} M rays: a full pool)", r) for b, r in zip((2, 3), body)]
if os.path.exists(f"{G}/prof_extend_steady.ncu-rep"):
    h2, u2, b2 = raw_rows(f"{G}/prof_extend_steady.ncu-rep")
    if b2 and h2 == hdr:
        row = list(b2[0])
        sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
        for i, (ua, ub) in enumerate(zip(units, u2)):  # ncu picks the unit per report: bring this one to the first report's units
            if ua != ub and ua in sc and ub in sc:
                try:
                    row[i] = f"{float(row[i]) * sc[ub] / sc[ua]:.6f}"
                except ValueError:
                    pass
        cols.append(("iteration 16 (steady state of the regenerating pool: 38 % new camera rays, the rest from bounces 1..15)", row))
with open(f"{out_dir}/{tag}_k_extend_ncu.txt", "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:k_extend -s 81 -c 2   (and -s 95 -c 1)\n")
    f.write("# launches of the timed step of: python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-workloads\n")
    f.write("# columns: " + " | ".join(n for n, _ in cols) + "\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            f.write(f"{w:85s} {units[i]:16s} {[c[1][i] for c in cols]}\n")
    i_n = hdr.index("smsp__inst_executed.sum")
    f.write(f"# warp instructions per ray (first column): {float(body[0][i_n]) / max(pool, 1):.1f}\n")
print(open(f"{out_dir}/{tag}_k_extend_ncu.txt").read())
i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tr = [float(r[i_r]) * scale[units[i_r]] + float(r[i_w]) * scale[units[i_w]] for r in body]
json.dump({"k_extend_dram_bytes_per_launch": tr[0], "captured_launch": "second iteration of the timed step (a full pool of rays)",
           "captured_launch_rays": pool,
           "all_captured_launches": tr, "k_extend_share_of_step_under_ncu": ext_share,
           "source": f"profiles/{tag}_k_extend_ncu.txt, profiles/{tag}_launches.txt"}, open(f"{out_dir}/roofline_ncu.json", "w"), indent=1)
