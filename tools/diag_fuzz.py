"""Prints how the device and the oracle differ on one mutated corpus (tests/test_oracle_fuzz.py):
    gpurun --timeout 200 -- 'python tools/diag_fuzz.py [prog] [seed]'
Every differing result key, the surplus / missing nat_log_rb records decoded, and the three stats vectors."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
PROG = sys.argv[1] if len(sys.argv) > 1 else "pipeline_up"
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 11
import numpy as np, harness, collections
import test_oracle_fuzz as F
from bng_b200 import layouts as L
from oracle import pyoracle
kind = "reference" if pyoracle.available("reference") else "port"
sc = F.fuzz_script(PROG, SEED)
want = harness.run_script(harness.OracleBackend(kind), F.fuzz_script(PROG, SEED))
be = harness.GpuBackend()
got = harness.run_script(be, F.fuzz_script(PROG, SEED)); be.close()
for k in sorted(want):
    a, b = np.asarray(want[k]), np.asarray(got[k])
    if a.shape != b.shape or not np.array_equal(a, b):
        print("DIFF", k, a.shape, b.shape)
a = want["ev_nat_log_rb"]; b = got["ev_nat_log_rb"]
ca = collections.Counter(bytes(r[8:36]) for r in a); cb = collections.Counter(bytes(r[8:36]) for r in b)
extra = cb - ca; missing = ca - cb
def dec(r):
    e = np.frombuffer(r, np.uint8)
    typ = int.from_bytes(r[0:4],'little'); sub = int.from_bytes(r[4:8],'little')
    return dict(type=typ, sub=sub, priv='.'.join(map(str,r[8:12])), pub='.'.join(map(str,r[12:16])), pport=int.from_bytes(r[16:18],'big'), pubport=int.from_bytes(r[18:20],'big'), dst='.'.join(map(str,r[20:24])), dport=int.from_bytes(r[24:26],'big'), proto=r[26], flags=r[27])
print("extra on gpu:", len(extra)); [print("  +", dec(r), n) for r, n in list(extra.items())[:12]]
print("missing on gpu:", len(missing)); [print("  -", dec(r), n) for r, n in list(missing.items())[:12]]
for name in ("st_nat_stats_map","st_qos_stats_map","st_antispoof_stats"):
    print(name, want[name], got[name])
