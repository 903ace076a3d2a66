"""pytest configuration: the `gpu` marker and import paths.

`-m "not gpu"` runs on a CPU-only box: oracle vs golden vectors, host logic, C-ABI symbol check.
`-m gpu` are the parity tests proper: they call the CUDA path through the C ABI and check it against the
oracle (tests/ is one of the few places allowed to import `oracle/`).
"""
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    from boss_runs_b200 import build, _lib
    build.build()
    return _lib.load()
