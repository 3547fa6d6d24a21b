import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def gl24():
    return np.polynomial.legendre.leggauss(24)


@pytest.fixture(scope="session")
def arrays_small():
    import torj_jl_b200 as tj

    return tj.solovev_arrays(65, 65)


@pytest.fixture(scope="session")
def arrays_full():
    import torj_jl_b200 as tj

    return tj.solovev_arrays(257, 257)  # BASELINE.json config 3 grid


@pytest.fixture(scope="session")
def oracle_small(arrays_small):
    from oracle import torj_oracle as O

    return O.OraclePlasma(*arrays_small.values())


@pytest.fixture(scope="session")
def oracle_full(arrays_full):
    from oracle import torj_oracle as O

    return O.OraclePlasma(*arrays_full.values())


@pytest.fixture(scope="session")
def launcher():
    """reference test/tests/setup.jl:64-74 geometry on the Solov'ev box."""
    import torj_jl_b200 as tj

    x0 = np.array([2.5, 0.0, 0.4])
    N0 = tj.pol_tor_angles_2_vector(np.deg2rad(30.0), 0.0)
    return dict(x0=x0, N0=N0, f=95e9, spot=0.0174, inv_Rc=1.0 / 3.99)
