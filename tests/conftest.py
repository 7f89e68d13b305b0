"""Shared fixtures.  `-m "not gpu"` = oracle vs golden/reference, host logic, C-ABI surface (no GPU);
`-m gpu` = the parity tests proper: CUDA path through the C ABI vs the CPU oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref_available():
    from oracle.pyoracle import Reference
    return Reference.available("k120") and Reference.available("k500")


@pytest.fixture(scope="session")
def pkg():
    from ft8b200_loader import load
    return load()


@pytest.fixture(scope="session")
def ctx(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    c = pkg.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ctx500(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    c = pkg.Context(0, max_candidates=500, max_messages=200)
    yield c
    c.close()


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
