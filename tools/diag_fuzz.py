"""Prints how the device and the oracle differ on one mutated corpus (tests/test_oracle_fuzz.py):
    gpurun --timeout 300 -- 'python tools/diag_fuzz.py pipeline_up 11 13 [--pinned]'
Every differing result key; per differing frame the verdicts and the input/output header bytes; table keys present on
one side only; the surplus / missing nat_log_rb records decoded; the stats vectors."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import numpy as np

import harness
import test_oracle_fuzz as F
from oracle import pyoracle

args = [a for a in sys.argv[1:] if not a.startswith("--")]
PINNED = "--pinned" in sys.argv
PROG = args[0] if args else "pipeline_up"
SEEDS = [int(x) for x in args[1:]] or [11]
kind = "reference" if pyoracle.available("reference") else "port"


def hx(b):
    return bytes(b).hex()


def dec(r):
    r = bytes(r)
    return dict(type=int.from_bytes(r[0:4], 'little'), sub=int.from_bytes(r[4:8], 'little'), priv='.'.join(map(str, r[8:12])),
                pub='.'.join(map(str, r[12:16])), pport=int.from_bytes(r[16:18], 'big'), pubport=int.from_bytes(r[18:20], 'big'),
                dst='.'.join(map(str, r[20:24])), dport=int.from_bytes(r[24:26], 'big'), proto=r[26], flags=r[27])


for SEED in SEEDS:
    print(f"===== {PROG} seed {SEED} pinned={PINNED} oracle={kind}")
    sc = F.fuzz_script(PROG, SEED)
    want = harness.run_script(harness.OracleBackend(kind), F.fuzz_script(PROG, SEED))
    be = harness.GpuBackend(pinned=PINNED)
    got = harness.run_script(be, F.fuzz_script(PROG, SEED))
    be.close()
    diffs = harness.diff_keys(want, got)
    for k, msg in diffs:
        print("DIFF", k, msg)
    runs = [(si, st) for si, st in enumerate(sc.steps) if st[0] == "run"]
    for si, st in runs:
        tag = f"s{si:03d}"
        _, prog, arena, lens, now, off16, stride, prio, _ = st
        vw, vg = want[tag + "_verdict"], got[tag + "_verdict"]
        fw = want[tag + "_frames"].reshape(-1, stride)
        fg = got[tag + "_frames"].reshape(-1, stride)
        fin = arena.reshape(-1, stride)
        bad = np.nonzero((vw != vg) | (fw != fg).any(axis=1))[0]
        print(f"-- run {tag}: {len(bad)} frames differ (verdict or bytes)")
        for i in bad[:40]:
            print(f"   frame {i} len {lens[i]} verdict want {vw[i]} got {vg[i]}")
            print(f"      in   {hx(fin[i][:64])}")
            print(f"      want {hx(fw[i][:64])}")
            print(f"      got  {hx(fg[i][:64])}")
    for t in ("nat_sessions", "nat_reverse", "eim_table", "subscriber_nat", "qos_ingress"):
        kw = {bytes(k): bytes(v) for k, v in zip(want["tk_" + t], want["tv_" + t])}
        kg = {bytes(k): bytes(v) for k, v in zip(got["tk_" + t], got["tv_" + t])}
        only_w = [k for k in kw if k not in kg]
        only_g = [k for k in kg if k not in kw]
        dv = [k for k in kw if k in kg and kw[k] != kg[k]]
        print(f"-- table {t}: want {len(kw)} got {len(kg)}; only oracle {len(only_w)}, only gpu {len(only_g)}, value differs {len(dv)}")
        for k in only_w[:10]:
            print("   only oracle", hx(k), hx(kw[k]))
        for k in only_g[:10]:
            print("   only gpu   ", hx(k), hx(kg[k]))
        for k in dv[:10]:
            print("   value", hx(k), "\n      want", hx(kw[k]), "\n      got ", hx(kg[k]))
    a = want["ev_nat_log_rb"]
    b = got["ev_nat_log_rb"]
    print("nat_log_rb records: want", a.shape, "got", b.shape)
    if a.shape[0] and b.shape[0] and a.shape[1] >= 36 and b.shape[1] >= 36:
        ca = collections.Counter(bytes(r[8:36]) for r in a)
        cb = collections.Counter(bytes(r[8:36]) for r in b)
        extra = cb - ca
        missing = ca - cb
        print("extra on gpu:", sum(extra.values()))
        for r, n in list(extra.items())[:20]:
            print("  +", dec(r), n)
        print("missing on gpu:", sum(missing.values()))
        for r, n in list(missing.items())[:20]:
            print("  -", dec(r), n)
    for name in ("st_nat_stats_map", "st_qos_stats_map", "st_antispoof_stats"):
        print(name, want[name].tolist(), got[name].tolist())
