#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libbng_b200.so (cuobjdump -sass): what the hand-written kernels compile to.
usage: python profiles/sass_hist.py [libbng_b200.so] > profiles/rNN_sass_hist.txt
Lists, per kernel, the instruction count, registers are in the ptxas logs; the memory / sync / warp-collective opcodes
(LDG/STG widths, ATOMG/RED, UBLKCP = cp.async.bulk (TMA), SYNCS = mbarrier, MATCH, REDUX, VOTE, SHFL, BAR) and the
ten most frequent opcodes."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "bng_b200", "libbng_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
kern, hist = None, {}
arch = set(re.findall(r"arch = (sm_\w+)", out))
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print(f"# {os.path.basename(lib)}: {len(hist)} kernels, architectures {sorted(arch)}")
KEY = ("LDG", "STG", "LDS", "STS", "ATOMG", "ATOMS", "RED", "UBLKCP", "SYNCS", "MATCH", "REDUX", "VOTE", "SHFL", "BAR", "LDC", "LDL", "STL",
       "MEMBAR", "FENCE", "NANOSLEEP", "CCTL", "ERRBAR")
for k in sorted(hist, key=lambda k: -sum(hist[k].values())):
    h = hist[k]
    tot = sum(h.values())
    name = demangle(k)
    print(f"\n== {name[:150]}\n   {tot} instructions")
    fam = collections.Counter()
    for op, c in h.items():
        base = op.split(".")[0]
        if base in KEY:
            fam[op] += c
    for op, c in sorted(fam.items()):
        print(f"   {op:34s} {c}")
    print("   top:", ", ".join(f"{op} {c}" for op, c in h.most_common(10)))
