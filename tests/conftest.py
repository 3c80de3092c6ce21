import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_bin():
    """Path of the CPU oracle binary (test infrastructure).  Built on demand with gcc."""
    out = os.path.join(ROOT, "oracle", "_build", "carmel_oracle")
    srcs = [os.path.join(ROOT, "oracle", f) for f in ("oracle_cli.cpp", "carmel_oracle.hpp", "gibbs_oracle.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return out


@pytest.fixture(scope="session")
def forest_oracle_bin(oracle_bin):
    """Path of the forest-em CPU oracle binary (same Makefile as carmel_oracle)."""
    out = os.path.join(ROOT, "oracle", "_build", "forest_oracle")
    srcs = [os.path.join(ROOT, "oracle", f) for f in ("forest_cli.cpp", "forest_oracle.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return out


@pytest.fixture(scope="session")
def native_lib():
    """libcarmel_b200.so, (re)built in-tree if sources changed and nvcc is present."""
    from carmel_b200 import build as _b
    return _b.build()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
