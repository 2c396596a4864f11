import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def kflib():
    from roskfpos_b200 import lib
    lib.lib()
    return lib


def pytest_sessionfinish(session, exitstatus):
    """Leave the observed parity numbers (stable fraction, tie count, worst error per test) on disk."""
    from tests import util
    out = os.environ.get("KFPOS_PARITY_REPORT", os.path.join(ROOT, "gpurun_out", "parity_report.json"))
    util.write_parity_report(out)
