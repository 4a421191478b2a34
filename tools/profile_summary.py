#!/usr/bin/env python
"""Summarise an `ncu --set full` report (and optionally a launch list) into profiles/<name>.md.

    python tools/profile_summary.py gpurun_out/prof_65536_r01.ncu-rep profiles/r01_k_step_65536 \
        [--launches gpurun_out/launches_4096_r01.csv] [--note "..."]

Writes <out>.md (human summary: duration, DRAM traffic per launch, occupancy, pipe utilisation, stall
reasons, per-section instruction attribution) and <out>.json (the raw numbers bench.py's roofline.traffic
reads). The .ncu-rep itself stays in gpurun_out/ (scratch).
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.sum.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict((h, (u, v)) for h, u, v in zip(hdr, units, r)) for r in rows[2:]]


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    rep, outp = sys.argv[1], sys.argv[2]
    launches = sys.argv[sys.argv.index("--launches") + 1] if "--launches" in sys.argv else None
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    kernels = raw(rep)
    md = [f"# ncu summary: {os.path.basename(rep)}", "", note, ""]
    js = {"report": os.path.basename(rep), "kernels": []}
    for k in kernels:
        name = k.get("Kernel Name", ("", ""))[1]
        d = {"kernel": name}
        md += [f"## {name}", "", "| metric | value | unit |", "|---|---|---|"]
        for key in KEYS:
            if key in k:
                u, v = k[key]
                d[key] = fnum(v)
                md.append(f"| {key} | {v} | {u} |")
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            unit = k["dram__bytes_read.sum"][0]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
            unit_w = k["dram__bytes_write.sum"][0]
            scale_w = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit_w]
            d["dram_traffic_bytes"] = rd * scale + wr * scale_w
            md.append(f"| **DRAM traffic per launch (read+write)** | {d['dram_traffic_bytes']:.4g} | byte |")
        # counted FP32 work (SURVEY section 8d ii): thread-level FFMA (x2), FADD, FMUL per launch = per-cycle rate x elapsed cycles
        fp = {op: fnum(k.get(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed", ("", ""))[1]) for op in ("ffma", "fadd", "fmul")}
        # (smsp__ per_cycle_elapsed rates are per SM-clock cycle; sm__cycles_elapsed.max is the same clock)
        cyc = fnum(k.get("sm__cycles_elapsed.max", k.get("smsp__cycles_elapsed.max", ("", "")))[1])
        if all(v is not None for v in fp.values()) and cyc:
            per_cycle = 2 * fp["ffma"] + fp["fadd"] + fp["fmul"]
            d["fp32_flops_per_cycle"] = per_cycle
            d["fp32_flops_per_launch"] = per_cycle * cyc
            d["fp32_frac_of_peak_under_ncu"] = per_cycle / (148 * 128 * 2)
            md.append(f"| **counted FP32 flops per launch (2 FFMA + FADD + FMUL, thread level)** | {d['fp32_flops_per_launch']:.4g} | flop |")
            md.append(f"| FP32 flops per SM-cycle, chip-wide (peak 148 x 128 x 2 = 37888) | {per_cycle:.1f} = {100 * per_cycle / 37888:.2f} % | flop/cycle |")
        stalls = sorted(((fnum(v[1]) or 0.0, h) for h, v in k.items()
                         if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")),
                        reverse=True)[:8]
        md += ["", "Warp stall reasons (warps stalled per issue-active cycle):", ""]
        for v, h in stalls:
            md.append(f"- {h.split('stalled_')[1].split('_per_issue')[0]}: {v:.3f}")
            d.setdefault("stalls", {})[h.split('stalled_')[1].split('_per_issue')[0]] = v
        md.append("")
        js["kernels"].append(d)
    if "k_step" in (kernels[0].get("Kernel Name", ("", ""))[1] if kernels else ""):
        env = dict(os.environ)
        sec = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_section.py"), rep], capture_output=True,
                             text=True, env=env).stdout
        md += ["## Instruction attribution by section of odg_core.cuh (tools/ncu_by_section.py)", "", "```", sec.rstrip(), "```", ""]
    if launches and os.path.exists(launches):
        rows = [r for r in csv.reader(open(launches)) if len(r) > 14 and r[0].isdigit()]
        agg = {}
        for r in rows:
            nm = r[4].split("(")[0]
            a = agg.setdefault(nm, [0, 0.0])
            a[0] += 1; a[1] += float(r[14])
        tot = sum(a[1] for a in agg.values())
        md += [f"## Launch list ({os.path.basename(launches)}): every launch of the bench command, serialised, cold cache", "",
               "| kernel | launches | total ns | share |", "|---|---|---|---|"]
        for nm, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            md.append(f"| {nm} | {a[0]} | {a[1]:.0f} | {a[1] / tot * 100:.1f}% |")
        js["launches"] = {nm: {"count": a[0], "ns": a[1]} for nm, a in agg.items()}
    open(outp + ".md", "w").write("\n".join(md) + "\n")
    json.dump(js, open(outp + ".json", "w"), indent=1)
    print("wrote", outp + ".md")


if __name__ == "__main__":
    main()
