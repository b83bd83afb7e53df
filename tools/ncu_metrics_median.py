"""Median per kernel of an `ncu --metrics ... --csv --log-file X` launch list:  python tools/ncu_metrics_median.py X.csv"""
import csv, statistics, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
acc = defaultdict(lambda: defaultdict(list))
unit = {}
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    acc[r[ix["Kernel Name"]]][r[ix["Metric Name"]]].append(v)
    unit[r[ix["Metric Name"]]] = r[ix["Metric Unit"]]
for k, m in acc.items():
    n = max(len(v) for v in m.values())
    print(f"{k[:62]:62s} x{n:3d}  " + "  ".join(f"{name} [{unit[name]}]={statistics.median(v):.4g}" for name, v in sorted(m.items())))
