"""Turn the ncu reports brought back in gpurun_out/ into the tracked summaries under profiles/.
Usage (CPU box, after a gpurun profiling call):  python profiles/summarize.py r01"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def raw_rows(rep):
    res = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(res.stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def unit_scale(u):
    return {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3}.get(u, None)


lines = [f"# ncu summaries, round {tag[1:]} (raw reports stay in gpurun_out/, not tracked)\n"]
traffic = {}
for rep_name in sorted(os.listdir(OUT)):
    if not (rep_name.startswith(tag + "_") and rep_name.endswith(".ncu-rep")):
        continue
    hdr, units, rows = raw_rows(os.path.join(OUT, rep_name))
    ki = hdr.index("Kernel Name")
    lines.append(f"\n## {rep_name}  (`ncu --set full --clock-control none`)\n")
    seen = collections.OrderedDict()
    for r in rows:
        seen.setdefault(r[ki], r)  # first captured launch of each kernel
    for kname, r in seen.items():
        short = kname.split("(")[0].replace("void ", "")
        lines.append(f"\n### `{short}`\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"| {k} | {r[i]} | {units[i]} |")
        try:
            rd = float(r[hdr.index('dram__bytes_read.sum')]) * unit_scale(units[hdr.index('dram__bytes_read.sum')])
            wr = float(r[hdr.index('dram__bytes_write.sum')]) * unit_scale(units[hdr.index('dram__bytes_write.sum')])
            traffic[short] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
                              "duration_us": float(r[hdr.index('gpu__time_duration.sum')]) * unit_scale(units[hdr.index('gpu__time_duration.sum')]) * 1e6}
        except Exception:
            pass

# launch list of the bench command
lpath = os.path.join(OUT, f"{tag}_launches.csv")
if os.path.exists(lpath):
    rows = [r for r in csv.reader(open(lpath)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki].split("(")[0].replace("void ", "")[:90], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    lines.append(f"\n## {tag}_launches.csv — every launch of `python bench.py --steps 3 --warmup 3 --no-extras` "
                 "(`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)\n")
    lines.append("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, v in agg.items():
        lines.append(f"| `{k}` | {len(v)} | {sum(v)/1e3:.1f} | {sum(v)/len(v)/1e3:.1f} | {sum(v)/tot:.3f} |")
    with open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as fh:
        fh.write(open(lpath).read())
# launch list of the TD3 gradient step (profiles/run_td3.py --batch 4096 --once, steady-state launches)
tpath = os.path.join(OUT, f"{tag}_td3_launches.csv")
if os.path.exists(tpath):
    rows = [r for r in csv.reader(open(tpath)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki].split("(")[0].replace("void ", "")[:90], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    lines.append(f"\n## {tag}_td3_launches.csv — launches 60..180 of `python profiles/run_td3.py --batch 4096 --once` (TD3 gradient steps, fp32 GEMM; "
                 "cold-cache, serialised: compare shares)\n")
    lines.append("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, v in agg.items():
        lines.append(f"| `{k}` | {len(v)} | {sum(v)/1e3:.1f} | {sum(v)/len(v)/1e3:.1f} | {sum(v)/tot:.3f} |")
    with open(os.path.join(PROF, f"{tag}_td3_launches.csv"), "w") as fh:
        fh.write(open(tpath).read())
open(os.path.join(PROF, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(os.path.join(PROF, f"{tag}_traffic.json"), "w"), indent=1)
print("\n".join(lines)[:6000])
