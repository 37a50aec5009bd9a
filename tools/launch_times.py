"""Per-kernel mean of gpu__time_duration.sum from an `ncu --csv` launch list:  python tools/launch_times.py file.csv"""
import collections
import csv
import sys

hdr, agg = None, collections.OrderedDict()
for r in csv.reader(open(sys.argv[1])):
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            agg.setdefault(d["Kernel Name"][:70], []).append(float(d["Metric Value"].replace(",", "")))
for k, v in agg.items():
    print(f"{k:72s} {len(v):4d} x {sum(v) / len(v) / 1e3:8.1f} us   (min {min(v) / 1e3:.1f})")
