"""pytest configuration: the ``gpu`` marker and shared fixtures.

``-m "not gpu"``: oracle vs golden vectors, host logic, code generation, C-ABI exports (no CUDA device needed).
``-m gpu``      : parity of the CUDA path (through the C ABI) against the oracle on a B200.
"""

from __future__ import annotations

import sys

from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu() -> bool:
    try:
        from geconpy_b200 import _lib

        return _lib.load_library().gecon_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # A gpu-marked test on a box without a device is an error of the run configuration, not a skip:
    # fail loudly (the product has no CPU fallback) unless the run deselected gpu tests with -m "not gpu".
    pass


@pytest.fixture(scope="session")
def gpu_available():
    return _has_gpu()


@pytest.fixture(scope="session")
def rng():
    return np.random.default_rng(20261017)
