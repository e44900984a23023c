import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests", "golden") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/liboracle.so), built on demand.  Test infrastructure only."""
    from oracle_lib import load_oracle

    return load_oracle()


@pytest.fixture(scope="session")
def gpu():
    """The CUDA product library opened on cuda:0 through the C ABI."""
    from visioncortex_b200.runtime import Context

    return Context(0)
