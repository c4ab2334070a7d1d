import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure): builds oracle/libdge_oracle.so on demand."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def dge_lib():
    """The product library; built in-tree with nvcc if missing (cross-compiles without a GPU)."""
    from embedding_b200 import build
    build.build()
    from embedding_b200 import abi
    abi.lib()
    return abi


@pytest.fixture(scope="session")
def ctx(dge_lib):
    c = dge_lib.Context(0)
    yield c
    c.close()
