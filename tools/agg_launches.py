#!/usr/bin/env python
"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list by
kernel name: launches, total time, share, DRAM bytes.   python tools/agg_launches.py file.csv [top]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tunit = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("slfp::", "")[:100]
    v = float(row["Metric Value"].replace(",", ""))
    a = agg[name]
    if row["Metric Name"] == "gpu__time_duration.sum":
        a[0] += 1
        a[1] += v * tunit.get(row["Metric Unit"].strip(), 1.0)
    elif row["Metric Name"].startswith("dram__bytes"):
        a[2] += v * unit.get(row["Metric Unit"].strip(), 1)
tot = sum(a[1] for a in agg.values())
print(f"| kernel | launches | total us | share | DRAM MB | GB/s |\n|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f} % | {a[2] / 1e6:.0f} | {a[2] / a[1] / 1e3 if a[1] else 0:.0f} |")
print(f"| **total** | {sum(a[0] for a in agg.values())} | {tot:.1f} | | {sum(a[2] for a in agg.values()) / 1e6:.0f} | |")
