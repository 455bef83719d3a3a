"""Generates tests/golden/*.npz by running every script of tests/scenarios.py
through the REFERENCE oracle (oracle/_ref/libbng_ref.so = the reference's own
eBPF C sources compiled natively).  Needs /root/reference to (re)build that
library, so it only runs in the build container; the fixtures it writes are
committed and travel.

    python tests/golden/make_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import harness  # noqa: E402
import scenarios  # noqa: E402
from oracle import pyoracle  # noqa: E402


def main():
    pyoracle.build("ref")
    for name, fn in scenarios.ALL_SCRIPTS.items():
        be = harness.OracleBackend("reference")
        res = harness.run_script(be, fn())
        path = os.path.join(HERE, name + ".npz")
        harness.save_golden(path, res)
        print(f"{name}: {len(res)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
