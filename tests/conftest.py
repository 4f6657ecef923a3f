import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(HERE, "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fixture_system():
    """The reference's bundled matrix + rhs, frozen by tests/golden/make_golden.py."""
    from oracle_bindings import CSR

    z = np.load(os.path.join(HERE, "golden", "fixture_poisson_P1.npz"))
    n = len(z["rowptr"]) - 1
    return CSR(n, n, z["rowptr"], z["colindex"], z["val"]), z["b"].astype(np.float64)


@pytest.fixture(scope="session")
def oracle():
    from oracle_bindings import Oracle

    o = Oracle.get()
    o.set_threads(min(8, os.cpu_count() or 1))
    return o


def system_by_name(name, oracle, fixture_system):
    """(A, b) for a golden.json case key."""
    if name == "fixture":
        return fixture_system
    base, kind = name.rsplit("_", 1)
    if base == "poisson3d_24":
        A = oracle.gen_poisson3d(24, 24, 24)
    elif base == "poisson2d_96":
        A = oracle.gen_poisson2d(96, 96)
    elif base == "poisson3d_40":
        A = oracle.gen_poisson3d(40, 40, 40)
    else:
        raise KeyError(name)
    if kind == "ones":
        b = np.ones(A.nrow)
    else:
        b = A.to_scipy() @ np.random.default_rng(42).random(A.nrow)
    return A, b
