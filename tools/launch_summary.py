"""Per-kernel mean of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv"""
import collections
import csv
import sys

for fn in sys.argv[1:]:
    rows = list(csv.reader(open(fn)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d["Metric Unit"], 1.0)
        agg.setdefault(d["Kernel Name"][:64], []).append(v)
    print(fn)
    for k, v in agg.items():
        print(f"  {k:64s} n={len(v):3d} mean={sum(v) / len(v):8.1f} us  min={min(v):8.1f}")
