import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `pytest -m gpu` under gpurun)")


@pytest.fixture(scope="session")
def built():
    """Compiles libb200sched.so and the oracle once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from oracle import pyoracle

    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def engine_mod(built):
    from scheduler_plugins_b200 import engine

    return engine


@pytest.fixture()
def eng(engine_mod):
    e = engine_mod.Engine(0)
    yield e
    e.close()
