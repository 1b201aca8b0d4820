#!/usr/bin/env python
"""Static SASS opcode mix of one kernel in an object / .so (no GPU needed).
  python tools/sass_mix.py <file.o|.so> <kernel-substring> [top_n]"""
import collections, re, subprocess, sys
path, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
cur, mix, total = None, collections.Counter(), 0
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1).split(".")[0]
            mix[op] += 1; total += 1
print(f"{pat}: {total} static instructions")
for op, n in mix.most_common(top):
    print(f"  {op:12s} {n:6d} {100*n/total:5.1f}%")
