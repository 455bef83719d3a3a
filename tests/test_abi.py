"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports
every symbol include/bng_b200.h declares, and refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "bng_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bng_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    fns = declared_functions()
    for must in ("bng_open", "bng_close", "bng_map_update", "bng_map_lookup", "bng_map_delete", "bng_prog_run",
                 "bng_events_drain", "bng_sync", "bng_stats_device_ptr", "bng_shard_of_mac"):
        assert must in fns


def test_library_exports_every_declared_symbol():
    from bng_b200 import dataplane
    lib = dataplane.load_library()
    missing = [f for f in declared_functions() if not hasattr(lib, f)]
    assert not missing, f"libbng_b200.so lacks {missing}"
    assert set(dataplane.EXPORTED_SYMBOLS) == set(declared_functions())
    assert lib.bng_abi_version() == 2


def test_shard_function_matches_host_mirror():
    import numpy as np
    from bng_b200 import shard_of_mac, synth
    macs = synth.sub_mac_key(np.arange(1000))
    for world in (1, 2, 4, 8):
        ours = synth.shard_of_mac(macs, world)
        assert all(shard_of_mac(int(m), world) == int(s) for m, s in zip(macs[:200], ours[:200]))
        if world > 1:
            counts = np.bincount(ours, minlength=world)
            assert counts.min() > 1000 / world * 0.7  # roughly balanced


def test_no_cpu_fallback():
    """Without a CUDA device bng_open() must fail loudly (this test is skipped on a GPU box)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from bng_b200 import Dataplane
    with pytest.raises(RuntimeError, match="no CUDA device"):
        Dataplane()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bng_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f in ("synth.py",) and "pyoracle" not in text, \
                    f"{os.path.join(dp, f)} mentions the oracle"
