import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return {name: np.load(os.path.join(GOLDEN, name + ".npz")) for name in
            ("frenet_rk4", "cartesian_euler", "geometry", "glue")}


@pytest.fixture(scope="session")
def oracle_params():
    from oracle import nlp
    base = nlp.Params(N=40)
    return {N: nlp.Params(N=N, cinf_A=base.cinf_A, cinf_b=base.cinf_b) for N in (10, 20, 40)}
