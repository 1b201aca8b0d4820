#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from an ncu report (needs -lineinfo + --import-source on).

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [min_pct]
"""
import csv, subprocess, sys, io
rep = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue   # keep source-level rows only
    d = dict(zip(hdr, r))
    try:
        lines.append((cur, int(r[0]), r[1], int(d["Instructions Executed"]), int(d["# Samples"]), d))
    except ValueError:
        pass
ti = sum(l[3] for l in lines); ts = sum(l[4] for l in lines)
print(f"total warp instr {ti}, samples {ts}")
for f, ln, src, ins, smp, d in lines:
    if ins / ti * 100 >= minpct or smp / ts * 100 >= minpct:
        tops = sorted(((k, int(v)) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()), key=lambda kv: -kv[1])[:2]
        print(f"{f[:16]:16} {ln:4d} ins {ins/ti*100:5.1f}% smp {smp/ts*100:5.1f}% {tops[0][0][6:]:>9}/{tops[1][0][6:]:<9} {src.strip()[:90]}")
