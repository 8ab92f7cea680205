import gzip
import json
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith(".gz"):
        with gzip.open(path, "rt") as f:
            return json.load(f)
    with open(path) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def enum_ff():
    return load_golden("enum_force_free_d4.json.gz")


@pytest.fixture(scope="session")
def enum_kerr():
    return load_golden("enum_kerr_magnetosphere_d3.json.gz")


@pytest.fixture(scope="session")
def resid_ff():
    return load_golden("resid_force_free.json.gz")


@pytest.fixture(scope="session")
def resid_kerr():
    return load_golden("resid_kerr_magnetosphere.json.gz")


def uniques_by_depth(g):
    return {int(d): g["depths"][d]["uniques"] for d in g["depths"]}


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
