"""Shared fixtures.  `-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, ABI.
`-m gpu` runs on a B200: parity of the CUDA path (through the C ABI) against the oracle / goldens."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    return O


@pytest.fixture(scope="session")
def port(oracle_mod):
    """The plain-C restatement (oracle/libpcd_oracle.so); compiled on first use."""
    return oracle_mod.OracleLib()


@pytest.fixture(scope="session")
def ref(oracle_mod):
    """The reference itself (oracle/_ref); only where it was built from /root/reference."""
    if not oracle_mod.RefLib.available():
        if os.path.isdir(os.path.join(oracle_mod.REFERENCE_ROOT, "src")):
            oracle_mod.build(ref=True)
        else:
            pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle_mod.RefLib()


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLD, name + ".npz")))
        return cache[name]

    return load


@pytest.fixture(scope="session")
def pcd():
    """The product: ctypes view of libpcd_b200.so.  Built in-tree if the .so is missing (nvcc only)."""
    import poisson_caustic_design_b200 as P
    if not os.path.exists(P.LIB_PATH):
        from poisson_caustic_design_b200 import build
        build.build_cuda()
    return P


def setup_from_params(oracle_mod, params):
    p = params
    return oracle_mod.Setup(int(p[0]), int(p[1]), int(p[2]), int(p[3]), float(p[4]), float(p[5]), float(p[6]), float(p[7]))


def rel_linf(a, b, scale=None):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    s = np.max(np.abs(b)) if scale is None else scale
    return float(np.max(np.abs(a - b)) / (s if s > 0 else 1.0))
