import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def lib_built():
    """libb2a.so, built in-tree (nvcc cross-compiles on the CPU box)."""
    from audio_edge_ml_pipeline_b200.build import build_lib
    return build_lib()
