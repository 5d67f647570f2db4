"""Test configuration: markers, import paths, fixture loaders."""

from __future__ import annotations

import json
import math
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "mcmc-db_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def dec(v):
    """Decode the nan/inf strings used in the JSON fixtures."""
    if isinstance(v, str):
        return {"nan": math.nan, "inf": math.inf, "-inf": -math.inf}[v]
    return v


@pytest.fixture(scope="session")
def ref_cases():
    return json.loads((GOLDEN / "ref_cases.json").read_text())


@pytest.fixture(scope="session")
def compare_cases():
    return json.loads((GOLDEN / "compare_cases.json").read_text())


@pytest.fixture(scope="session")
def packaged_subset():
    arrays = np.load(GOLDEN / "packaged_subset.npz")
    meta = json.loads((GOLDEN / "packaged_subset.json").read_text())
    return arrays, meta


def same_float(a: float, b: float) -> bool:
    """Bit-level equality that treats nan == nan."""
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    return a == b


def close(a: float, b: float, rel: float) -> bool:
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= rel * max(abs(a), abs(b), 1e-300)
