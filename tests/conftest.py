import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def golden_cases():
    return sorted(f[:-3] for f in os.listdir(GOLDEN_DIR) if f.startswith("sfno_") and f.endswith(".pt"))


@pytest.fixture(scope="session")
def load_golden():
    import torch

    cache = {}

    def _load(name):
        if name not in cache:
            cache[name] = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), map_location="cpu", weights_only=False)
        return cache[name]

    return _load
