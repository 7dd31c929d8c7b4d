import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("qwen3-tts-axera-russian_b200")


@pytest.fixture(scope="session")
def backend():
    import __graft_entry__ as g
    if not os.path.exists(g.LIB) or not os.path.exists(g.SERVER):
        g.build(force=True)
    return importlib.import_module("qwen3-tts-axera-russian_b200.backend")


@pytest.fixture(scope="session")
def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "dual_npu"))
