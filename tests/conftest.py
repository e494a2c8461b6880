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
def oracle():
    from oracle.loader import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "fabe13_golden_v1.npz")
    z = np.load(path)
    cases = {}
    for name in z["case_names"]:
        name = str(name)
        cases[name] = {
            "x": z[f"{name}/x"].view(np.float64),
            "ref_sin": z[f"{name}/ref_sin"].view(np.float64),
            "ref_cos": z[f"{name}/ref_cos"].view(np.float64),
            "libm_sin": z[f"{name}/libm_sin"].view(np.float64),
            "libm_cos": z[f"{name}/libm_cos"].view(np.float64),
            "q_exact": z[f"{name}/q_exact"],
            "k_ref": z[f"{name}/k_ref"],
        }
    return cases


@pytest.fixture(scope="session")
def fabe13():
    """The product binding; building the library is part of the test session (nvcc cross-compiles without a GPU)."""
    from fabe_b200.build import build_library

    build_library()
    import fabe_b200

    return fabe_b200


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("test marked gpu but no CUDA device is visible")
    return torch.device("cuda:0")
