"""CPU-side checks of the boundary: the shared library builds, loads without a GPU and
exports every symbol include/b200ref.h declares (no compute calls here)."""

from __future__ import annotations

import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from mcmc_ref_b200 import _native

    return _native.load()


def test_header_symbols_exported(lib):
    header = (ROOT / "include" / "b200ref.h").read_text()
    names = set(re.findall(r"\b(b200ref_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 13
    from mcmc_ref_b200 import _native

    assert names == set(_native.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n


def test_version_and_workspace_queries(lib):
    assert lib.b200ref_version() == 100
    small = lib.b200ref_workspace_min_bytes(4, 1000)
    big = lib.b200ref_workspace_bytes(1000, 16, 10000)
    assert 0 < small < big
    # 40 B/draw of scratch (2 x key64 + 2 x idx32 + 2 x z) plus the z table
    assert big >= 1000 * 160000 * 40
    assert lib.b200ref_last_error() == b""


def test_summary_layout_matches_header():
    header = (ROOT / "include" / "b200ref.h").read_text()
    from mcmc_ref_b200 import batch

    idx = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"#define B200REF_S_([A-Z0-9_]+) (\d+)", header)}
    for i, col in enumerate(batch.SUMMARY_COLUMNS):
        assert idx[col] == i
    assert int(re.search(r"#define B200REF_SUMMARY_K (\d+)", header).group(1)) == len(batch.SUMMARY_COLUMNS)


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU the product path must fail loudly, not compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mcmc_ref_b200 import _native, diagnostics

    with pytest.raises(_native.NativeError, match="no CPU fallback"):
        diagnostics.split_rhat([[1.0, 2.0, 3.0, 4.0]] * 4)


def test_guards_match_reference_messages():
    """Guards run before any device work (diagnostics.py:24-30, :88-90)."""
    from mcmc_ref_b200 import diagnostics

    two = [[1.0, 2.0, 3.0, 4.0], [1.1, 2.1, 3.1, 4.1]]
    with pytest.raises(ValueError, match="R-hat diagnostics require at least 4 chains; got 2 chain"):
        diagnostics.split_rhat(two)
    with pytest.raises(ValueError, match="ESS diagnostics require at least 4 chains; got 2 chain"):
        diagnostics.ess_bulk(two)
    with pytest.raises(ValueError, match="ESS diagnostics require at least 4 chains"):
        diagnostics.ess_tail(two)
    with pytest.raises(ValueError, match="min_chains must be >= 1; got 0"):
        diagnostics.split_rhat(two, min_chains=0)
    import math

    assert math.isnan(diagnostics.split_rhat([[1.0, 2.0, 3.0, 4.0]], min_chains=1))
    assert math.isnan(diagnostics.ess_bulk([[1.0, 2.0, 3.0, 4.0]], min_chains=1))


def test_backend_registry_contract():
    from mcmc_ref_b200 import backends

    assert set(backends.BACKENDS) == {"b200"}
    with pytest.raises(ValueError, match="Unknown backend: numpy"):
        backends.get_backend("numpy")


def test_product_never_imports_oracle():
    pkg = ROOT / "mcmc-db_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = path.read_text()
        assert "oracle" not in text.replace("the oracle", ""), path
