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


def pytest_sessionstart(session):
    """The shared library is a build artefact (not in git).  If a fresh checkout has not been built yet and the CUDA
    toolchain is present, build it once (nvcc cross-compiles sm_100a without a GPU) so the ABI / host-logic tests have
    something to load; without nvcc the tests that need the library fail loudly, as the product does."""
    import shutil
    import subprocess

    lib = os.path.join(ROOT, "spherical-dyffusion_b200", "libsfno_b200.so")
    if os.path.isfile(lib) or os.environ.get("PYTEST_XDIST_WORKER"):
        return
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "spherical-dyffusion_b200", "csrc"), "-j", str(min(8, os.cpu_count() or 2))],
                       check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(autouse=True, scope="session")
def _host_logic_cold_update():
    """The product's cold-sampling update is a CUDA kernel with no CPU path.  The host-logic tests drive the window program
    with CPU stand-ins for the networks (the oracle), so for CPU tensors -- and only for those -- the three-operand update is
    evaluated with torch here, in the tests.  CUDA tensors still take the product's kernel."""
    import torch

    from spherical_dyffusion_b200.dyffusion import DYffusion

    product = DYffusion._cold

    def cold(x, nxt, cur):
        if x.is_cuda:
            return product(x, nxt, cur)
        return x + (nxt - cur)

    DYffusion._cold = staticmethod(cold)
    yield
    DYffusion._cold = staticmethod(product)
