"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv CMD`)
into per-kernel launch counts, total time and share:   python tools/ncu_launch_summary.py X.csv > profiles/....txt"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(int\)|\(bool\)", "", name)
    name = re.sub(r"\(CUtensorMap_st.*$|\(TapGemmParams.*$|\(.*$", "", name)
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
    agg[name][0] += 1
    agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
print(f"# total {tot:.2f} ms over {n} launches")
print("kernel,launches,ms,share")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{c},{ms:.3f},{ms / tot:.4f}")
