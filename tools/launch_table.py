"""Print an ncu launch-list CSV (gpu__time_duration.sum) as a table: python tools/launch_table.py file.csv [substr-to-skip]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
skip = sys.argv[2:] or []
tot = 0.0
for r in rows[1:]:
    if any(s in r[ki] for s in skip):
        continue
    t = float(r[vi].replace(",", "")) / 1000
    tot += t
    print(f"{t:8.1f} us  {r[ki][:70]}")
print(f"total {tot:.1f} us")
