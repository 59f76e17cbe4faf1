"""Attribute the SASS-level counters of an ncu report to CUDA source lines (the ncu CLI prints them per SASS instruction only).

    python tools/ncu_lines.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [LAUNCH_INDEX] [TOP]

LIB.so must be the very build the report was captured from (compiled with -lineinfo): its cubin is disassembled with
`nvdisasm --print-line-info-inline` and joined with the report's SASS page by instruction offset.
Prints, per source line (innermost inlined frame; NCU_LINES_CALLER=1: the line that called it), warp instructions, stall samples
and active lanes.
"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40

# 1. SASS page of the chosen launch
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
kname = rows[0][1] if rows and len(rows[0]) > 1 else "?"
hdr = rows[1]
c_addr, c_src = hdr.index("Address"), hdr.index("Source")
c_inst, c_thr, c_smp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
sass = []
for r in rows[2:]:
    if len(r) <= c_inst or not r[c_addr].startswith("0x"):
        continue
    try:
        sass.append((int(r[c_addr], 16), r[c_src], float(r[c_inst]), float(r[c_thr]), float(r[c_smp])))
    except ValueError:
        pass
base = sass[0][0]
# a report page may hold the same kernel twice (ncu concatenates): keep the first run of increasing addresses
for k in range(1, len(sass)):
    if sass[k][0] <= sass[k - 1][0]:
        sass = sass[:k]
        break

# 2. line table of the same function from the library's cubin
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
func_want = re.sub(r"[^A-Za-z0-9_]", "", kern)
line_of = {}
for cb in cubins:
    dis = subprocess.run(["nvdisasm", "--print-line-info-inline", cb], capture_output=True, text=True).stdout
    in_func, cur, stack, fresh = False, None, [], True
    for ln in dis.splitlines():
        if ln.startswith("\t.section\t.text."):
            name = ln.split(".text.", 1)[1].split(",")[0]
            in_func = func_want in name and "$" not in name and not line_of
            cur, stack, fresh = None, [], True
            continue
        if not in_func:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:  # a stack of frames, innermost first, precedes the instructions it covers
            if fresh:
                stack, fresh = [], False
            stack.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
        if m:
            if not fresh:  # innermost frame that is this repo's code (not a CUDA header)
                own = [f for f in stack if "/cuda/" not in f[0]] or stack
                cur = (os.path.basename(own[0][0]), own[0][1], None)
                if os.environ.get("NCU_LINES_CALLER") and len(own) > 1:  # attribute to the line that called the innermost frame
                    cur = (os.path.basename(own[1][0]), own[1][1], None)
                fresh = True
            if cur:
                line_of[int(m.group(1), 16)] = cur
if not line_of:
    sys.exit(f"no function matching {kern!r} with line info in {lib}")

src_cache = {}


def text(fn, n):
    if fn not in src_cache:
        for root in ("raytracer-2022_b200/csrc", "include", "."):
            p = os.path.join(root, fn)
            if os.path.exists(p):
                src_cache[fn] = open(p, errors="replace").read().splitlines()
                break
        else:
            src_cache[fn] = []
    L = src_cache[fn]
    return L[n - 1].strip()[:100] if 0 < n <= len(L) else ""


by_line = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
tot = [0.0, 0.0, 0.0]
miss = 0
for addr, src, inst, thr, smp in sass:
    key = line_of.get(addr - base)
    if key is None:
        miss += 1
        key = ("?", 0, None)
    a = by_line[(key[0], key[1])]
    a[0] += inst; a[1] += thr; a[2] += smp
    tot[0] += inst; tot[1] += thr; tot[2] += smp
print(f"# {kname[:110]}")
print(f"# launch {launch}: {len(sass)} SASS instructions ({miss} without a line), {tot[0] / 1e6:.1f} M warp instructions, "
      f"{tot[1] / max(tot[0], 1):.1f} active lanes / instruction, {int(tot[2])} stall samples")
print(f"# {'instr %':>8s} {'samples %':>9s} {'lanes':>5s}  line")
for (fn, n), a in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {a[0] / tot[0] * 100:8.2f} {a[2] / max(tot[2], 1) * 100:9.2f} {a[1] / max(a[0], 1):5.1f}  {fn}:{n}  {text(fn, n)}")
