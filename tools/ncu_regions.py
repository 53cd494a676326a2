#!/usr/bin/env python
"""SASS regions of a kernel in an .ncu-rep (consecutive instructions with the same execution count): instructions executed
and stall samples per region.  usage: python tools/ncu_regions.py report.ncu-rep kernel-regex [min-share-percent]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ia, isrc, iss, ie = h.index("Address"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
seen, data = set(), []
for r in rows[2:]:
    if len(r) < len(h) or r[ia] in seen:
        continue
    try:
        data.append((r[isrc], int(r[iss]), int(r[ie])))
    except ValueError:
        continue
    seen.add(r[ia])
tot_i, tot_s = sum(d[2] for d in data), sum(d[1] for d in data)
print(f"== {rows[0][1][:70]}: {tot_i} warp-instructions executed, {tot_s} stall samples")
i = 0
while i < len(data):
    j = i
    while j + 1 < len(data) and data[j + 1][2] == data[i][2]:
        j += 1
    n, ex = j - i + 1, data[i][2]
    inst, samp = n * ex, sum(d[1] for d in data[i:j + 1])
    if ex and (100.0 * inst / tot_i >= min_share or 100.0 * samp / max(1, tot_s) >= 2.0):
        print(f"   SASS lines {i:4d}-{j:4d}: executed {ex:9d} x {n:3d} instr = {100.0 * inst / tot_i:5.1f} % of instructions, "
              f"{100.0 * samp / max(1, tot_s):5.1f} % of samples   first: {data[i][0][:48]}")
    i = j + 1
