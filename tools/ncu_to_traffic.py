#!/usr/bin/env python
"""Regenerates profiles/ncu_traffic.json (read by bench.py's `roofline.traffic`) and a text summary from ONE
`ncu --set full --clock-control none` pass over every own kernel of one cfg2 view.  Runs in the CPU container.
usage: python tools/ncu_to_traffic.py gpurun_out/<all>.ncu-rep profiles/r02/ncu_all_kernels_v1.txt [source-tag]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, txt = sys.argv[1], sys.argv[2]
tag = sys.argv[3] if len(sys.argv) > 3 else os.path.relpath(txt, ROOT)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_seconds(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1, "second": 1}.get(unit, 1e-9)


kernels = {}
lines = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "?")
    short = re.sub(r"\(.*", "", name).split("::")[-1].strip()
    short = re.sub(r"^void ", "", short)
    lines.append("===== " + name[:110])
    for w in WANT:
        if w in d:
            lines.append(f"  {w:78s} {d[w]:>16s} {units[hdr.index(w)]}")
    u = lambda k: units[hdr.index(k)]
    entry = {
        "dram_bytes": to_bytes(d["dram__bytes_read.sum"], u("dram__bytes_read.sum")) + to_bytes(d["dram__bytes_write.sum"], u("dram__bytes_write.sum")),
        "duration_s": to_seconds(d["gpu__time_duration.sum"], u("gpu__time_duration.sum")),
        "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
        "warp_instructions": float(d["smsp__inst_executed.sum"].replace(",", "")),
        "registers": int(d["launch__registers_per_thread"]),
    }
    key = short
    k = 2
    while key in kernels:  # the same kernel launched several times in a view (sort passes)
        key = f"{short}#{k}"
        k += 1
    kernels[key] = entry
os.makedirs(os.path.dirname(txt), exist_ok=True)
open(txt, "w").write("\n".join(lines) + "\n")
out = {"source": f"{tag} (ncu --set full --clock-control none --import-source on, one launch per own kernel of one cfg2 "
                 f"view through bench.py, B200; per-launch times are cold-cache and serialised)", "kernels": kernels}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
for k, v in kernels.items():
    print(f"{k:42s} {v['duration_s'] * 1e6:8.1f} us  dram {v['dram_bytes'] / 1e6:8.1f} MB  issue {v['issue_active_pct']:5.1f} %  regs {v['registers']}")
