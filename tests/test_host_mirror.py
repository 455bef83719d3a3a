"""The C++ host mirror of the reference's Go boundary types (bng_b200/host/bng_host.hpp),
exercised by tests/host/test_host.cpp, which is modelled on the reference's own Go tests."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "test_host.cpp")
BIN = os.path.join(ROOT, "tests", "host", "test_host")


def build_host_test():
    deps = [SRC, os.path.join(ROOT, "bng_b200", "host", "bng_host.hpp"), os.path.join(ROOT, "include", "bng_b200.h")]
    if os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", SRC, "-o", BIN, "-L" + os.path.join(ROOT, "bng_b200"), "-lbng_b200",
                    "-Wl,-rpath,$ORIGIN/../../bng_b200"], check=True)


def test_host_mirror_without_device():
    build_host_test()
    r = subprocess.run([BIN, "cpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_host_mirror_on_gpu():
    build_host_test()
    r = subprocess.run([BIN, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
