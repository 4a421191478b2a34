#!/usr/bin/env python
"""Static control-flow map of k_step: BSSY / BRA / BSYNC instructions per source line of odg_core.cuh (innermost
inlined frame), from `nvdisasm -gi` of the built library. Branches end the basic blocks ptxas can schedule over and
cost branch_resolving stalls; in this latency-bound kernel removing them paid far more than their instruction share.

    python tools/sass_branches.py [lo_line hi_line]
"""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("ODG_LIB_PATH", os.path.join(ROOT, "opendog_b200", "libodgsim.so"))
kern = os.environ.get("ODG_KERNEL", "k_stepILi2ELb1")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = next(f for f in os.listdir(tmp) if f.endswith(".cubin") and "odg_sim." in f)
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
src = open(os.path.join(ROOT, "opendog_b200/csrc/odg_core.cuh")).read().splitlines()
lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1, len(src))
in_k, chain, fresh, cur = False, [], True, -1
per, total = {}, 0
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
    m2 = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m2:
        if not fresh:
            core = [ln for f, ln in chain if f.endswith("odg_core.cuh")]
            cur = core[0] if core else -1
            fresh = True
        total += 1
        op = m2.group(1)
        if op in ("BSSY", "BRA", "BSYNC", "BRX", "CALL", "RET", "WARPSYNC") and lo <= cur <= hi:
            per.setdefault(cur, {}).setdefault(op, 0)
            per[cur][op] += 1
print(f"{total} instructions in {kern}; control-flow instructions by source line {lo}-{hi}:")
for ln in sorted(per):
    print(f"{ln:5d} {' '.join(f'{k}:{v}' for k, v in sorted(per[ln].items())):28s} {src[ln - 1].strip()[:100] if ln > 0 else '?'}")
