"""bench.py's JSON-line contract for the arms that need no GPU: the reference arm (the reference's eBPF C on the host
cores) and BASELINE config #1 (DHCP slow path).  One line on stdout, the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from oracle import pyoracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"]


def run_bench(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout  # exactly one line, and it is the JSON
    return json.loads(lines[0])


@pytest.mark.skipif(not (pyoracle.available("reference") or pyoracle.available("port")), reason="no oracle library built")
@pytest.mark.parametrize("workload", ["antispoof_64", "nat_ingress_64"])
def test_reference_arm_line(workload):
    j = run_bench("--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "0")
    for k in REQUIRED:
        assert k in j, k
    assert j["impl"] == "reference" and j["value"] > 0 and j["unit"] == "Mpps"
    assert j["config"]["workload"] == workload
    # both arms describe WHAT they ran with the same dict (the driver compares them); how an arm ran is under "details"
    sys.path.insert(0, ROOT)
    import bench
    assert j["config"] == bench.workload_config(workload, 1 << 22, 1) and "host_procs" in j["details"]
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0 and j["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_config1_dhcp_slow_line():
    j = run_bench("--workload", "dhcp_slow", "--steps", "1")
    for k in REQUIRED:
        assert k in j, k
    assert j["n_gpus"] == 0 and j["gpu_launches"] == 0 and j["roofline"] is None
    assert j["unit"] == "requests/s" and j["value"] > 1e4
    assert j["config"]["clients_with_lease"] == 256 and j["config"]["requests_per_step"] == 1000
