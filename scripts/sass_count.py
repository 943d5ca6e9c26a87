#!/usr/bin/env python
"""Static SASS instruction counts per kernel of libpbh_b200.so (cuobjdump -sass): total and by opcode class.  The per-item
routines are straight-line code inside a tile loop, so the static count of a kernel is a good first estimate of the
executed instructions per item (the ncu captures under profiles/ have the executed counts).
usage: python scripts/sass_count.py <regex> [top-n opcodes]"""
import collections, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "plonk-by-fingers_b200", "libpbh_b200.so")
pat = re.compile(sys.argv[1]); top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name, counts = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and name:
        counts[name][m.group(1).split(".")[0]] += 1
for k, c in counts.items():
    if pat.search(k):
        print(f"{sum(c.values()):6d}  {k[:110]}")
        print("        " + "  ".join(f"{o}:{n}" for o, n in c.most_common(top)))
