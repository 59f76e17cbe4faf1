"""Share of a kernel's executed instructions spent in the out-of-line subroutines ptxas appends to it (the full path of an f64
division, the medium path of sqrt, reciprocal slow paths, noinline device functions): these carry no line info, so tools/ncu_lines.py
cannot name them.

    python tools/ncu_subfuncs.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [LAUNCH_INDEX]

Then lists the call sites of each subroutine (source line of the CALL, times executed, active lanes).
"""
import csv, os, re, subprocess, sys, tempfile

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
c_addr, c_inst, c_thr, c_smp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
sass = []
for r in rows[2:]:
    if len(r) > c_inst and r[c_addr].startswith("0x"):
        try:
            sass.append((int(r[c_addr], 16), float(r[c_inst]), float(r[c_thr]), float(r[c_smp])))
        except ValueError:
            pass
for k in range(1, len(sass)):
    if sass[k][0] <= sass[k - 1][0]:
        sass = sass[:k]
        break
base = sass[0][0]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
want = re.sub(r"[^A-Za-z0-9_]", "", kern)
starts = []  # (offset, name)
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    dis = subprocess.run(["nvdisasm", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    inside, pending = False, None
    for ln in dis.splitlines():
        if ln.startswith("\t.section\t.text."):
            inside = want in ln and not starts
            pending = "main"
            continue
        if not inside:
            continue
        m = re.match(r"\s*\.type\s+(\S+),@function", ln)
        if m:
            pending = m.group(1)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and pending:
            name = pending.split("$")[-1] if "$" in pending else "main body"
            starts.append((int(m.group(1), 16), name))
            pending = None
starts.sort()
tot = [sum(x[1] for x in sass), sum(x[3] for x in sass)]
print(f"# {want}: {len(sass)} SASS instructions, {tot[0] / 1e6:.1f} M warp instructions")
for k, (off, name) in enumerate(starts):
    end = starts[k + 1][0] if k + 1 < len(starts) else 1 << 60
    seg = [x for x in sass if off <= x[0] - base < end]
    n, t, smp = sum(x[1] for x in seg), sum(x[2] for x in seg), sum(x[3] for x in seg)
    print(f"  {n / tot[0] * 100:6.2f} % instr  {smp / max(tot[1], 1) * 100:6.2f} % samples  {t / max(n, 1):5.1f} lanes  {len(seg):5d} SASS  {name}")

# call sites: CALL instructions of the main body, with the source line that precedes them in the line-info listing
calls = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    dis = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    inside, stack, fresh = False, [], True
    for ln in dis.splitlines():
        if ln.startswith("\t.section\t.text."):
            inside = want in ln and not calls
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            if fresh:
                stack, fresh = [], False
            stack.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
        if m:
            fresh = True
            if "CALL" in m.group(2) and "`(" in m.group(2):
                tgt = m.group(2).split("`(")[1].split(")")[0].split("$")[-1]
                own = [x for x in stack if x[0].startswith("rt_")] or stack
                calls[int(m.group(1), 16)] = (tgt, own[:2])
by_off = {x[0] - base: x for x in sass}
agg = {}
for off, (tgt, own) in calls.items():
    x = by_off.get(off)
    if not x or x[1] == 0:
        continue
    key = (tgt, tuple(own))
    a = agg.setdefault(key, [0.0, 0.0])
    a[0] += x[1]; a[1] += x[2]
print("# call sites (executed CALLs, lanes): callee <- line [<- caller line]")
for (tgt, own), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"  {a[0] / 1e6:8.2f} M  {a[1] / max(a[0], 1):5.1f}  {tgt[:40]:40s} <- " + " <- ".join(f"{f}:{n}" for f, n in own))
