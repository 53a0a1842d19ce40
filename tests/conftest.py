import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The C restatement (oracle/liblouvain_oracle.so), built on demand."""
    from oracle.binding import Oracle
    if not os.path.exists(Oracle.path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liblouvain_oracle.so"], check=True)
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The real reference CPU path; skipped where oracle/_ref was never built."""
    from oracle.binding import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libcuvite_ref.so not built (needs /root/reference)")
    return Reference()


@pytest.fixture(scope="session")
def cases():
    from graphs import small_cases
    return small_cases()


def load_golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))


GOLDEN = ["ring6x5", "er1k", "rmat10", "rmat10_intw", "rmat9_realw", "hub30k", "empty_rows", "multi9"]


@pytest.fixture(scope="session")
def gpu():
    """A GpuGraph on cuda:0 through the C ABI; the CUDA library must load."""
    from cuvite_b200 import build
    build.build()
    from cuvite_b200.engine import GpuGraph, device_count
    assert device_count() > 0, "gpu test on a box without a CUDA device"
    g = GpuGraph()
    yield g
    g.close()
