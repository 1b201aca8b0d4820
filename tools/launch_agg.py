"""Aggregate the second half (= the last pass) of an ncu launch-list CSV by kernel name."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
rows = rows[1:]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = rows[int(len(rows) * frac):]
agg, tot = collections.OrderedDict(), 0.0
for r in rows:
    t = float(r[vi].replace(",", "")) / 1000
    tot += t
    a = agg.setdefault(r[ki][:84], [0, 0.0])
    a[0] += 1
    a[1] += t
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:4d} {t:9.1f} us  {k}")
print(f"total {tot:.1f} us, {len(rows)} launches")
