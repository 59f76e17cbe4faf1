"""Share of a kernel's executed instructions spent in the out-of-line subroutines ptxas appends to it (the full path of an f64
division, the medium path of sqrt, reciprocal slow paths, noinline device functions): these carry no line info, so tools/ncu_lines.py
cannot name them.

    python tools/ncu_subfuncs.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [LAUNCH_INDEX]
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
