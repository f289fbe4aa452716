import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def assets_dirs():
    from mcray_tracing_b200 import assets
    return assets.ensure_all()


@pytest.fixture(scope="session")
def built():
    """Everything compiled (product library, oracle, and the reference probe when possible)."""
    import __graft_entry__ as g
    g.build()
    return True
