import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The CPU oracle is compiled on demand (gcc); the CUDA library must already be built
    (python -c 'import __graft_entry__ as g; g.build()') -- tests never build a fallback."""
    from oracle import oracle as O
    if not os.path.exists(O.ORACLE_SO):
        O.build()
    lib = os.path.join(ROOT, "node_speex_resampler_b200", "libspeexb200.so")
    if not os.path.exists(lib):
        subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.build()"], cwd=ROOT, check=True)
    yield


def has_gpu() -> bool:
    try:
        from node_speex_resampler_b200 import lib
        return lib().spxb_device_count() > 0
    except Exception:
        return False
