"""pytest configuration: registers the `gpu` marker and exposes the package / oracle loaders."""
import importlib
import os
import sys

import pytest

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO_ROOT not in sys.path:
    sys.path.insert(0, REPO_ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_package():
    """The product package (its directory name is not an identifier, hence importlib)."""
    return importlib.import_module("multi-adapter-particles_b200")


def load_oracle():
    """The CPU oracle: test infrastructure only (oracle/oracle.h)."""
    return importlib.import_module("oracle.oracle_py")


@pytest.fixture(scope="session")
def mapc():
    pkg = load_package()
    # `make` is incremental: a library older than its sources must never be what the tests exercise
    if os.path.exists(os.path.join(pkg.PKG_DIR, "csrc", "Makefile")) and os.path.exists("/usr/local/cuda/bin/nvcc"):
        pkg.build()
    elif not os.path.exists(pkg.LIB_PATH):
        pkg.build()
    pkg.load()
    return pkg


@pytest.fixture(scope="session")
def oracle():
    orc = load_oracle()
    orc.load()
    return orc


@pytest.fixture(scope="session")
def gpu(mapc):
    """Fails (does not skip) when no device is usable: a GPU test must never pass on a fallback."""
    n = mapc.device_count()
    assert n >= 1, "no CUDA device visible"
    return n
