"""The diagnostics-facing parts of the reference's ``mcmc_ref.convert``:
``_chains_from_table`` (convert.py:150-161), ``_compute_diagnostics`` (:134-147),
``_checks`` (:164-172), ``_enforce_checks`` (:175-178).  File-format conversion
(CSV / json-zip -> Parquet) is out of scope.

``chains_tensor_from_table`` replaces the per-parameter ``to_pylist`` + bucket + sort
of the reference with one (chain, draw) ordering shared by all parameters: an Arrow
parameter column whose rows are chain-major is already the device layout ``x[p][C][N]``.
"""

from __future__ import annotations

from collections.abc import Iterable

import numpy as np

from . import batch
from .backend_b200 import column_to_numpy


def _row_order(table) -> tuple[np.ndarray | None, int, int]:
    """Row permutation that sorts by (chain, draw); None if rows are already in that order."""
    chain = column_to_numpy_int(table, "chain")
    draw = column_to_numpy_int(table, "draw")
    chains, counts = np.unique(chain, return_counts=True)
    if len(set(counts.tolist())) > 1:
        raise ValueError("ragged chains are not supported by the b200 backend; pass equal-length chains")
    c, n = len(chains), int(counts[0]) if len(counts) else 0
    key = chain.astype(np.int64) * (int(draw.max()) + 1 if draw.size else 1) + draw.astype(np.int64)
    if np.all(key[1:] >= key[:-1]):
        return None, c, n
    return np.lexsort((draw, chain)), c, n


def column_to_numpy_int(table, name: str) -> np.ndarray:
    if hasattr(table, "column"):
        col = table.column(name)
        return np.asarray(col.to_numpy(zero_copy_only=False) if hasattr(col, "to_numpy") else col)
    return np.asarray(table[name])


def chains_tensor_from_table(table, params: Iterable[str]) -> np.ndarray:
    """[P, C, N] float64 ordered by (chain, draw) -- batched ``_chains_from_table``."""
    order, c, n = _row_order(table)
    cols = []
    for p in params:
        v = column_to_numpy(table, p)
        if order is not None:
            v = v[order]
        cols.append(v.reshape(c, n))
    return np.stack(cols) if cols else np.empty((0, c, n))


def _chains_from_table(table, param: str) -> list[list[float]]:
    return chains_tensor_from_table(table, [param])[0].tolist()


def _compute_diagnostics(table, params: Iterable[str], *, min_chains: int = 4) -> dict[str, dict[str, float]]:
    params = list(params)
    if min_chains < 1:
        raise ValueError(f"min_chains must be >= 1; got {min_chains}")
    if not params:
        return {}
    x = chains_tensor_from_table(table, params)
    c = x.shape[1]
    if c < min_chains:
        raise ValueError(f"R-hat diagnostics require at least {min_chains} chains; got {c} chain(s)")
    nan = float("nan")
    if c < 2 or x.shape[2] == 0:
        return {p: {"rhat": nan, "ess_bulk": nan, "ess_tail": nan} for p in params}
    out = batch.diagnostics_batch(x, quantiles=())
    rhat = out["rhat"].cpu().numpy()
    eb = out["ess_bulk"].cpu().numpy()
    et = out["ess_tail"].cpu().numpy()
    return {p: {"rhat": float(rhat[i]), "ess_bulk": float(eb[i]), "ess_tail": float(et[i])}
            for i, p in enumerate(params)}


def _checks(n_chains: int, n_draws: int, diag: dict[str, dict[str, float]]) -> dict[str, bool]:
    ess_ok = all(values.get("ess_bulk", 0.0) > 400 for values in diag.values())
    rhat_ok = all(values.get("rhat", 1.0) < 1.01 for values in diag.values())
    return {
        "ndraws_is_10k": n_chains * n_draws == 10_000,
        "nchains_is_gte_4": n_chains >= 4,
        "ess_above_400": ess_ok,
        "rhat_below_1_01": rhat_ok,
    }


def _enforce_checks(checks: dict[str, bool]) -> None:
    failures = [name for name, ok in checks.items() if not ok]
    if failures:
        raise ValueError(f"quality checks failed: {', '.join(failures)}")
