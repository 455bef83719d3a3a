"""GPU parity: the CUDA dataplane, driven through the C ABI, against the CPU
oracle(s) and the committed golden vectors, bit for bit.  Every script runs
twice: with pageable host buffers (whole-arena staging copies) and with pinned
host buffers (zero-copy header gather/scatter pipeline)."""
import os

import pytest

import harness
import scenarios

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gpu_results():
    cache = {}

    def get(script, pinned):
        if (script, pinned) not in cache:
            be = harness.GpuBackend(pinned=pinned)
            try:
                cache[(script, pinned)] = harness.run_script(be, scenarios.ALL_SCRIPTS[script]())
            finally:
                be.close()
        return cache[(script, pinned)]

    return get


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "pinned"])
@pytest.mark.parametrize("script", sorted(scenarios.ALL_SCRIPTS))
def test_gpu_matches_golden(script, pinned, gpu_results):
    gold = harness.load_golden(os.path.join(GOLD, script + ".npz"))
    harness.compare(gold, gpu_results(script, pinned), f"{script}: golden vs gpu ({'pinned' if pinned else 'pageable'})")


@pytest.mark.parametrize("script", ["pipeline", "nat", "dhcp"])
def test_gpu_matches_golden_abi_host_arena(script, gpu_results):
    """Frames in memory obtained from bng_host_alloc() (huge-page backed, cudaHostRegister'ed)."""
    gold = harness.load_golden(os.path.join(GOLD, script + ".npz"))
    harness.compare(gold, gpu_results(script, "abi"), f"{script}: golden vs gpu (bng_host_alloc arena)")


@pytest.mark.parametrize("script", sorted(scenarios.ALL_SCRIPTS))
def test_gpu_matches_live_oracle(script, ora_kind, gpu_results):
    if ora_kind == "none":
        pytest.fail("no oracle library present on this box")
    be = harness.OracleBackend(ora_kind)
    res = harness.run_script(be, scenarios.ALL_SCRIPTS[script]())
    harness.compare(res, gpu_results(script, False), f"{script}: {ora_kind} oracle vs gpu")
