import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "wire_format.npz"))


def golden_cases(g, prefix):
    """Yield (base key, parsed tokens) of every golden case whose key starts with `prefix`."""
    seen = set()
    for k in g.files:
        if k.startswith(prefix):
            base = k.rsplit("_", 1)[0]
            if base not in seen:
                seen.add(base)
                yield base, base.split("_")
