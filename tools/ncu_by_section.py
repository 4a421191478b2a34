#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counters to sections of odg_core.cuh.

    python tools/ncu_by_section.py gpurun_out/prof.ncu-rep [kernel-substring] [--lines=40]

Joins `ncu --page source --csv` (SASS rows, in program order) with `nvdisasm -g` line info of the
same cubin (extracted from opendog_b200/libodgsim.so), then buckets by the `// ----` section markers.
"""
import bisect
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--lines")]
    top_lines = next((int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("--lines=")), 0)
    rep = argv[0]
    kern = argv[1] if len(argv) > 1 else "k_stepILi2EE"
    so = os.environ.get("ODG_LIB_PATH", os.path.join(ROOT, "opendog_b200", "libodgsim.so"))
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cubins = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
    cubin = next((f for f in cubins if "odg_sim." in f), cubins[0])
    dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    src_path = os.environ.get("ODG_CORE_SRC", os.path.join(ROOT, "opendog_b200", "csrc", "odg_core.cuh"))
    src = open(src_path).read().splitlines()
    sub_lo = next(i + 1 for i, l in enumerate(src) if l.startswith("ODG_DEV void substep("))
    sub_hi = next(i + 1 for i, l in enumerate(src) if i + 1 > sub_lo and l.startswith("}"))
    # instructions of the kernel, each attributed to the OUTERMOST odg_core.cuh frame inside substep()
    # (so inlined helpers like dot()/cross() count for the section that calls them), else the outermost frame
    in_k, lines_of_inst, cur, chain, fresh = False, [], None, [], True
    for l in dis:
        if l.startswith("//---") and ".text." in l:
            in_k = kern in l
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((m.group(1), int(m.group(2))))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
            if not fresh:
                core = [ln for f, ln in chain if f.endswith("odg_core.cuh")]
                inside = [ln for ln in core if sub_lo <= ln <= sub_hi]
                cur = inside[-1] if inside else (core[-1] if core else -1)
                fresh = True
            lines_of_inst.append(cur)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
    rows = list(csv.reader(out))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    data = [dict(zip(hdr, r)) for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    n = min(len(data), len(lines_of_inst))
    print(f"sass rows {len(data)}, disasm instructions {len(lines_of_inst)}")
    marks = [(i + 1, l.strip()) for i, l in enumerate(src) if l.strip().startswith("// ----") or l.startswith("template <int NJL>") or l.startswith("ODG_DEV")]
    mlines = [m[0] for m in marks]
    agg = {}
    tot = sum(num(d["Instructions Executed"]) for d in data)
    tots = sum(num(d["# Samples"]) for d in data)
    for i in range(n):
        ln = lines_of_inst[i]
        if ln is None or ln < 0:
            key = (0, "odg_sim.cu / other")
        else:
            k = bisect.bisect_right(mlines, ln) - 1
            key = marks[k] if k >= 0 else (0, "pre")
        a = agg.setdefault(key, [0.0, 0.0, 0.0, 0])
        d = data[i]
        a[0] += num(d["Instructions Executed"]); a[1] += num(d["Thread Instructions Executed"])
        a[2] += num(d["# Samples"]); a[3] += 1
    print(f"total warp-instructions {tot:.4g}, samples {tots:.0f}")
    print("  line   %inst  thr/inst  %samples  static  section")
    for (ln, name), a in sorted(agg.items()):
        print(f"{ln:6d} {a[0] / tot * 100:6.2f}%  {a[1] / max(a[0], 1):6.1f}  {a[2] / max(tots, 1) * 100:7.2f}%  {a[3]:6d}  {name[:80]}")
    if top_lines:
        per = {}
        for i in range(n):
            a = per.setdefault(lines_of_inst[i], [0.0, 0.0, 0.0, 0])
            d = data[i]
            a[0] += num(d["Instructions Executed"]); a[1] += num(d["Thread Instructions Executed"])
            a[2] += num(d["# Samples"]); a[3] += 1
        print(f"\ntop {top_lines} source lines of odg_core.cuh by executed warp-instructions")
        print("  line   %inst  thr/inst  %samples  static  source")
        for ln, a in sorted(per.items(), key=lambda kv: -kv[1][0])[:top_lines]:
            text = src[ln - 1].strip() if ln and 0 < ln <= len(src) else "?"
            print(f"{ln or 0:6d} {a[0] / tot * 100:6.2f}%  {a[1] / max(a[0], 1):6.1f}  {a[2] / max(tots, 1) * 100:7.2f}%  {a[3]:6d}  {text[:110]}")


if __name__ == "__main__":
    main()
