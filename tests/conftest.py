"""Shared test plumbing.

* registers the ``gpu`` marker (tests that need a B200 and call through the C-ABI library);
* puts ``oracle/`` on sys.path — tests are one of the three places allowed to import the oracle;
* exposes the product package (its directory name has a hyphen, so it is imported by path name).
"""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG_NAME = "pytorch-rl-enhancedstablebaselines_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))

    return load


def ulp32(a, b):
    """|a-b| in units of float32 spacing at b."""
    a = np.asarray(a, np.float64)
    b32 = np.asarray(b, np.float32)
    return np.abs(a - b32.astype(np.float64)) / np.spacing(np.abs(b32)).astype(np.float64)
