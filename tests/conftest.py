import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def port():
    from oracle.api import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    from oracle.api import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libgsref.so not built (needs /root/reference; run oracle/build_ref.py)")
    return Ref()


@pytest.fixture(scope="session")
def best_oracle():
    """The compiled reference kernels when present, else the C port (pinned against them)."""
    from oracle.api import Port, Ref
    return Ref() if Ref.available() else Port()


@pytest.fixture(scope="session")
def c1():
    from gaussiansplattingmlx_b200.scene import make_workload
    return make_workload("C1")
