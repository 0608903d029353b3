"""pytest plumbing.  `-m "not gpu"` = oracle vs goldens, host logic, ABI surface (no GPU needed);
`-m gpu` = parity tests proper, calling the CUDA path through the C ABI and failing loudly if it is missing."""
import importlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "codex-storage-proofs-circuits_b200"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def pkg():
    importlib.import_module(PKG + ".build").ensure_built()      # no-op when the in-tree build exists
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def orc():
    """the C oracle (checker)"""
    from oracle import coracle
    coracle.build()
    return coracle


@pytest.fixture(scope="session")
def pyorc():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="session")
def vectors():
    with open(os.path.join(GOLDEN, "vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ctx(pkg):
    """GPU context through the C ABI.  No skip: a gpu-marked test without the CUDA library or a device must fail."""
    c = pkg.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def orc_mod(orc):
    return orc


@pytest.fixture(scope="session")
def pyorc_mod(pyorc):
    return pyorc
