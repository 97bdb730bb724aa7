import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle as orc

    orc.build()
    orc.lib()
    return orc


_CASES = {}


def make_case(name, n, dim, nq, k, flavour="sift", seed=11):
    """Small seeded dataset + oracle-built index, cached per session.  Returns dict(base, queries, centroids, oracle, arrays)."""
    key = (name, n, dim, nq, k, flavour, seed)
    if key in _CASES:
        return _CASES[key]
    from oracle import oracle as orc
    from tools import synth

    base, queries, cent = synth.make_numpy(n, dim, nq, k, flavour, seed)
    ix = orc.OracleIndex.from_arrays(base, cent, seed=seed + 100, nthreads=8)
    case = dict(base=base, queries=queries, centroids=cent, oracle=ix, arrays=ix.arrays())
    _CASES[key] = case
    return case


@pytest.fixture(scope="session")
def case_d128(oracle_lib):
    return make_case("d128", 20000, 128, 64, 64, "sift", 11)


@pytest.fixture(scope="session")
def case_d96(oracle_lib):
    # 96 -> padded to 128 (DEEP-shaped, config 3)
    return make_case("d96", 12000, 96, 48, 48, "deep", 12)


@pytest.fixture(scope="session")
def case_d960(oracle_lib):
    return make_case("d960", 6000, 960, 24, 16, "gist", 13)
