"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_shares.py file.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, data = rows[0], rows[1:]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
agg = collections.OrderedDict()
for r in data:
    short = r[ik].split("(")[0].split("::")[-1]
    a = agg.setdefault(short, [0.0, 0])
    a[0] += float(r[iv].replace(",", "")) * scale[r[iu]]
    a[1] += 1
tot = sum(a[0] for a in agg.values())
for k, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{ms:10.3f} ms {n:4d} launches {100 * ms / tot:5.1f}%  {k[:60]}")
print(f"{len(data)} launches; sum {tot:.3f} ms")
