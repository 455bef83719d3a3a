"""The oracles against the committed golden vectors.

tests/golden/*.npz were produced by the reference's own eBPF C sources run
natively (tests/golden/make_golden.py).  Every oracle library present — the
reference build and the plain-C port — must reproduce them bit for bit; this
is what pins the port before it is trusted as a checker."""
import os

import numpy as np
import pytest

import harness
import scenarios

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("script", sorted(scenarios.ALL_SCRIPTS))
def test_oracle_reproduces_golden(script, ora_kind):
    if ora_kind == "none":
        pytest.fail("no oracle library built: run `make -C oracle`")
    gold = harness.load_golden(os.path.join(GOLD, script + ".npz"))
    be = harness.OracleBackend(ora_kind)
    res = harness.run_script(be, scenarios.ALL_SCRIPTS[script]())
    harness.compare(gold, res, f"{script}: golden vs {ora_kind} oracle")


def test_goldens_exercise_the_quirks():
    """Guards against vacuous fixtures: the corpora must reach the paths SURVEY.md §8a calls out."""
    g = harness.load_golden(os.path.join(GOLD, "antispoof.npz"))
    st = g["st_antispoof_stats"]
    assert st[0] > 0 and st[1] > 0 and st[2] > 0 and st[3] > 0 and st[4] > 0 and st[5] == 0
    g = harness.load_golden(os.path.join(GOLD, "nat_exhaust.npz"))
    nat = g["st_nat_stats_map"]
    assert nat[3] > 0 and nat[7] > 0, "port exhaustion / drop not reached"
    g = harness.load_golden(os.path.join(GOLD, "nat_stale.npz"))
    assert g["st_nat_stats_map"][6] > 0, "sessions_expired not reached"
    g = harness.load_golden(os.path.join(GOLD, "nat.npz"))
    nat = g["st_nat_stats_map"]
    assert nat[0] > 0 and nat[1] > 0 and nat[2] > 0 and nat[8] > 0 and nat[9] > 0 and nat[10] > 0
    g = harness.load_golden(os.path.join(GOLD, "dhcp.npz"))
    d = g["st_stats_map"]
    assert all(d[i] > 0 for i in (0, 1, 2, 3, 4, 5, 7, 8, 9)) and d[6] == 0
    v = np.concatenate([g[k] for k in g if k.endswith("_verdict")])
    assert set(np.unique(v)) == {2, 3}
    g = harness.load_golden(os.path.join(GOLD, "qos.npz"))
    assert all(x > 0 for x in g["st_qos_stats_map"])
    g = harness.load_golden(os.path.join(GOLD, "pipeline.npz"))
    assert g["st_qos_stats_map"][1] > 0 and g["st_nat_stats_map"][5] > 0 and g["st_antispoof_stats"][1] > 0
