"""GPU parity: the CUDA dataplane, driven through the C ABI, against the CPU
oracle(s) and the committed golden vectors, bit for bit."""
import os

import numpy as np
import pytest

import harness
import scenarios

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gpu_results():
    cache = {}

    def get(script):
        if script not in cache:
            be = harness.GpuBackend()
            try:
                cache[script] = harness.run_script(be, scenarios.ALL_SCRIPTS[script]())
            finally:
                be.close()
        return cache[script]

    return get


@pytest.mark.parametrize("script", sorted(scenarios.ALL_SCRIPTS))
def test_gpu_matches_golden(script, gpu_results):
    gold = harness.load_golden(os.path.join(GOLD, script + ".npz"))
    harness.compare(gold, gpu_results(script), f"{script}: golden vs gpu")


@pytest.mark.parametrize("script", sorted(scenarios.ALL_SCRIPTS))
def test_gpu_matches_live_oracle(script, ora_kind, gpu_results):
    if ora_kind == "none":
        pytest.fail("no oracle library present on this box")
    be = harness.OracleBackend(ora_kind)
    res = harness.run_script(be, scenarios.ALL_SCRIPTS[script]())
    harness.compare(res, gpu_results(script), f"{script}: {ora_kind} oracle vs gpu")
