"""The C++ host mirror of the reference's Go boundary types (bng_b200/host/bng_host.hpp),
exercised by tests/host/test_host.cpp, which is modelled on the reference's own Go tests."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "test_host.cpp")
BIN = os.path.join(ROOT, "tests", "host", "test_host")


HOST = os.path.join(ROOT, "bng_b200", "host")
SLOW_SRC = os.path.join(HOST, "dhcp_slow_bench.cpp")
SLOW_BIN = os.path.join(HOST, "dhcp_slow_bench")


def _stale(binary, deps):
    return not os.path.exists(binary) or any(os.path.getmtime(binary) < os.path.getmtime(d) for d in deps)


def build_host_test():
    hdrs = [os.path.join(HOST, h) for h in ("bng_host.hpp", "bng_dhcp_slow.hpp", "bng_nat_log.hpp", "bng_shard.hpp")] + \
        [os.path.join(ROOT, "include", "bng_b200.h")]
    lib = ["-L" + os.path.join(ROOT, "bng_b200"), "-lbng_b200"]
    if _stale(BIN, [SRC] + hdrs):
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", SRC, "-o", BIN] + lib + ["-Wl,-rpath,$ORIGIN/../../bng_b200"], check=True)
    if _stale(SLOW_BIN, [SLOW_SRC] + hdrs):  # BASELINE config #1: the DHCP slow path restated, CPU only
        subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", SLOW_SRC, "-o", SLOW_BIN] + lib + ["-Wl,-rpath,$ORIGIN/.."], check=True)


def test_dhcp_slow_path_bench_runs():
    """config #1 plumbing: 1 000 DISCOVERs, 256 clients with a lease; every one is answered with an OFFER and the
    replies are reproducible (FNV-1a over the first round's bytes)."""
    import json
    build_host_test()
    a = json.loads(subprocess.run([SLOW_BIN, "3"], capture_output=True, text=True, check=True).stdout)
    b = json.loads(subprocess.run([SLOW_BIN, "2"], capture_output=True, text=True, check=True).stdout)
    assert a["requests"] == 3000 and a["offers"] == 3000 and a["reply_bytes"] == 3000 * 300
    assert a["replies_fnv1a"] == b["replies_fnv1a"]


def test_host_mirror_without_device():
    build_host_test()
    r = subprocess.run([BIN, "cpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_host_mirror_on_gpu():
    build_host_test()
    r = subprocess.run([BIN, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
