#!/usr/bin/env python
"""Many more seeds of the mutation fuzz than the test suite runs (tests/test_oracle_fuzz.py): every program, seeds
[lo, hi), pageable and pinned alternating, device vs the reference oracle.  Prints the combinations that differ.
    gpurun --timeout 900 -- 'python tools/fuzz_wide.py 100 140'"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import harness
import test_oracle_fuzz as F
from oracle import pyoracle

lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100, 120)
kind = "reference" if pyoracle.available("reference") else "port"
bad = ran = 0
t0 = time.time()
for prog in F.TARGETS:
    for seed in range(lo, hi):
        pinned = bool(seed & 1)
        want = harness.run_script(harness.OracleBackend(kind), F.fuzz_script(prog, seed))
        be = harness.GpuBackend(pinned=pinned)
        try:
            got = harness.run_script(be, F.fuzz_script(prog, seed))
        finally:
            be.close()
        d = harness.diff_keys(want, got)
        ran += 1
        if d:
            bad += 1
            print(f"DIFF {prog} seed {seed} pinned={pinned}: {[k for k, _ in d][:6]}", flush=True)
print(f"{ran} runs, {bad} with differences, {time.time() - t0:.0f} s ({kind} oracle)")
sys.exit(1 if bad else 0)
