#!/usr/bin/env python
"""Summarise an .ncu-rep (run in the CPU container): key raw metrics per kernel + hottest SASS lines.
usage: python tools/ncu_summary.py gpurun_out/prof_blend.ncu-rep [kernel-regex]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=====", d.get("Kernel Name", "?")[:80])
    for w in want:
        if w in d:
            print(f"  {w:72s} {d[w]:>16s} {units[hdr.index(w)]}")

def hot(kernel):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) < 3:
        return
    hdr = rows[1]
    ia, ie, it, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed"), hdr.index("# Samples")
    tot = sum(int(r[ie]) for r in rows[2:])
    stot = sum(int(r[isamp]) for r in rows[2:])
    print(f"===== hot SASS of {kernel}: total warp-instructions {tot/1e6:.1f}M, samples {stot}")
    prev = None
    for r in rows[2:]:
        n = int(r[ie])
        if n > tot * 0.004:
            key = round(n / 2e5)
            if key != prev:
                print(f"  ---- {n/1e6:.2f}M executions")
            prev = key
            print(f"     thr={r[it]:>4s} samp={100*int(r[isamp])/max(1,stot):5.2f}%  {r[ia].strip()}")

if rx:
    hot(rx)
