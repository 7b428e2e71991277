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


def _gpu_available():
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU: the gpu-marked tests are skipped, not failed (the product has no
    CPU path -- BatchSolver raises -- so they cannot run there).  On a GPU box nothing is skipped: a missing
    library then fails the tests loudly."""
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (the solver has no CPU path)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return {name: np.load(os.path.join(GOLDEN, name + ".npz")) for name in
            ("frenet_rk4", "cartesian_euler", "geometry", "glue")}


@pytest.fixture(scope="session")
def oracle_params():
    from oracle import nlp
    base = nlp.Params(N=40)
    return {N: nlp.Params(N=N, cinf_A=base.cinf_A, cinf_b=base.cinf_b) for N in (10, 20, 40)}
