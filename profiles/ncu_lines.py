#!/usr/bin/env python
"""Top source lines of a kernel by warp-stall samples, from an .ncu-rep captured with --import-source on:
    python profiles/ncu_lines.py report.ncu-rep [top N] [kernel substring]
Aggregates the per-SASS-instruction samples of `ncu --page source --print-source cuda,sass` onto (file, line)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
filt = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = {}
fname = func = None
hdr = None
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        func = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        si = [i for i, c in enumerate(hdr) if c.startswith("Warp Stall Sampling (All")][0]
        ii = hdr.index("Instructions Executed")
        lsb = hdr.index("stall_long_sb")
        continue
    if hdr is None or len(r) <= si or (filt and filt not in (func or "")):
        continue
    if r[0]:
        cur = (fname, int(r[0]), r[1].strip()[:100])
        continue
    if cur is None:
        continue
    try:
        s, n, l = int(r[si] or 0), int(r[ii] or 0), int(r[lsb] or 0)
    except ValueError:
        continue
    a = agg.setdefault(cur, [0, 0, 0])
    a[0] += s
    a[1] += n
    a[2] += l
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot}, warp instructions {toti}")
print(f"{'samples':>8} {'%':>5} {'long_sb':>7} {'inst %':>6}  file:line  source")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0]:8d} {100 * a[0] / tot:5.1f} {a[2]:7d} {100 * a[1] / toti:6.1f}  {k[0]}:{k[1]}  {k[2]}")
