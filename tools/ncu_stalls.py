#!/usr/bin/env python
"""Stall-sample breakdown of a kernel from an ncu report: total by reason, and by section / source line of odg_core.cuh.

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [kernel-substring] [--lines=30]
"""
import csv, os, re, subprocess, sys, tempfile, bisect, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
top = next((int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("--lines=")), 30)
rep = argv[0]; kern = argv[1] if len(argv) > 1 else "k_stepILi2EE"
so = os.path.join(ROOT, "opendog_b200", "libodgsim.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = next(f for f in os.listdir(tmp) if f.endswith(".cubin") and "odg_sim." in f)
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
src = open(os.path.join(ROOT, "opendog_b200", "csrc", "odg_core.cuh")).read().splitlines()
SUB_LO = next(i + 1 for i, l in enumerate(src) if l.startswith("ODG_DEV void substep("))
SUB_HI = next(i + 1 for i, l in enumerate(src) if i + 1 > SUB_LO and l.startswith("}"))
in_k, lines, chain, fresh, cur, ops = False, [], [], True, None, []
for l in dis:
    if l.startswith("//---") and ".text." in l:
        in_k = kern in l; continue
    if not in_k: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh: chain, fresh = [], False
        chain.append((m.group(1), int(m.group(2)))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", l)
    if m:
        if not fresh:
            core = [ln for f, ln in chain if f.endswith("odg_core.cuh")]
            inside = [ln for ln in core if SUB_LO <= ln <= SUB_HI]
            cur = inside[-1] if inside else (core[-1] if core else -1)   # outermost frame inside substep(), else outermost
            fresh = True
        lines.append(cur); ops.append(m.group(1))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [dict(zip(hdr, r)) for r in rows[hi + 1:] if len(r) == len(hdr)]
n = min(len(data), len(lines))
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); totS = 0; totI = 0
per = {}
for i in range(n):
    d = data[i]; s = num(d["# Samples"]); totS += s; totI += num(d["Instructions Executed"])
    a = per.setdefault(lines[i], [0.0, 0.0, collections.Counter(), 0.0])
    a[0] += s; a[1] += num(d["Instructions Executed"]); a[3] += num(d["Thread Instructions Executed"])
    for r in reasons:
        v = num(d[r]); tot[r] += v; a[2][r] += v
print(f"kernel {kern}: sass {n}, samples {totS:.0f}, warp-instructions {totI:.4g}")
print("stall reasons (share of samples):", ", ".join(f"{r[6:]} {v / totS * 100:.1f}%" for r, v in tot.most_common(10)))
marks = [(i + 1, l.strip()) for i, l in enumerate(src) if l.strip().startswith("// ----")]
ml = [m[0] for m in marks]
sec = {}
for ln, a in per.items():
    k = bisect.bisect_right(ml, ln) - 1 if ln and ln > 0 else -1
    key = marks[k] if k >= 0 else (0, "other")
    b = sec.setdefault(key, [0.0, 0.0, collections.Counter(), 0.0])
    b[0] += a[0]; b[1] += a[1]; b[2].update(a[2]); b[3] += a[3]
print("\n  line  %samples  %inst  cyc/inst  lanes  top stalls   section")
for (ln, name), a in sorted(sec.items()):
    ts = ", ".join(f"{r[6:]} {v / max(a[0], 1) * 100:.0f}%" for r, v in a[2].most_common(3))
    print(f"{ln:6d}  {a[0] / totS * 100:6.2f}%  {a[1] / totI * 100:5.2f}%  {a[0] / totS / max(a[1] / totI, 1e-9):5.2f}  {a[3] / max(a[1], 1):5.1f}  {ts:42s} {name[:70]}")
print(f"\ntop {top} source lines by samples")
for ln, a in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    ts = ", ".join(f"{r[6:]} {v / max(a[0], 1) * 100:.0f}%" for r, v in a[2].most_common(3))
    text = src[ln - 1].strip() if ln and 0 < ln <= len(src) else "?"
    print(f"{ln or 0:6d}  {a[0] / totS * 100:6.2f}%  {a[1] / totI * 100:5.2f}%  lanes {a[3] / max(a[1], 1):4.1f}  {ts:40s} {text[:90]}")
