"""Comparison utilities -- mirror of the reference's ``mcmc_ref.compare`` (compare.py).

``ParamResult`` / ``CompareResult`` have the same fields (the CLI serialises them with
``vars()``, cli.py:317-321); relative errors and verdicts come from the compare kernel,
the failure strings are formatted exactly as the reference does (compare.py:45).
"""

from __future__ import annotations

from collections.abc import Mapping, Sequence
from dataclasses import dataclass

import numpy as np

from . import batch


@dataclass(frozen=True)
class ParamResult:
    ref: float
    actual: float
    rel_error: float
    passed: bool


@dataclass(frozen=True)
class CompareResult:
    passed: bool
    details: dict[str, dict[str, ParamResult]]
    failures: list[str]


def compare_stats(
    ref_stats: Mapping[str, Mapping[str, float]],
    actual_stats: Mapping[str, Mapping[str, float]],
    tolerance: float,
    metrics: Sequence[str],
) -> CompareResult:
    details: dict[str, dict[str, ParamResult]] = {}
    failures: list[str] = []
    metrics = list(metrics)
    present = [p for p in ref_stats if p in actual_stats]
    nan = float("nan")
    ref_vals = np.array([[float(ref_stats[p].get(m, nan)) for m in metrics] for p in present], dtype=np.float64)
    act_vals = np.array([[float(actual_stats[p].get(m, nan)) for m in metrics] for p in present], dtype=np.float64)
    if ref_vals.size:
        rel_t, ok_t, _ = batch.compare_batch(ref_vals, act_vals, tolerance)
        rel = rel_t.cpu().numpy().reshape(ref_vals.shape)
        ok = ok_t.cpu().numpy().reshape(ref_vals.shape)
    row = {p: i for i, p in enumerate(present)}
    for param in ref_stats:
        if param not in actual_stats:
            failures.append(f"missing param: {param}")
            continue
        i = row[param]
        param_details: dict[str, ParamResult] = {}
        for j, metric in enumerate(metrics):
            rel_error = float(rel[i, j])
            passed = bool(ok[i, j])
            if not passed:
                failures.append(f"{param}.{metric} rel_error={rel_error:.3g} > {tolerance}")
            param_details[metric] = ParamResult(
                ref=float(ref_vals[i, j]), actual=float(act_vals[i, j]), rel_error=rel_error, passed=passed
            )
        details[param] = param_details
    return CompareResult(passed=len(failures) == 0, details=details, failures=failures)


def _nonfinite_stats(row: np.ndarray) -> dict[str, float]:
    """Candidate draws holding NaN / inf (a diverged sampler -- the case the tool exists to flag).  The reference's
    pure-Python ``sum(values) / n`` gives nan or +-inf for the mean and ``(inf - inf)**2`` = nan for the spread
    (compare.py:58-64); ``compare_stats`` then reports ``rel_error=nan > tol`` and the verdict is a FAILED compare,
    not an exception.  No arithmetic on the draws is needed for that answer."""
    with np.errstate(invalid="ignore", over="ignore"):
        mean = float(np.sum(row) / row.size)
    return {"mean": mean, "std": float("nan")}


def compute_basic_stats(values: Sequence[float]) -> dict[str, float]:
    n = len(values)
    if n == 0:
        return {"mean": float("nan"), "std": float("nan")}
    row = np.asarray(values, dtype=np.float64).reshape(n)
    if not np.isfinite(row).all():
        return _nonfinite_stats(row)
    out = batch.stats_batch(row.reshape(1, n), quantiles=())
    return {"mean": float(out["mean"][0].item()), "std": float(out["std"][0].item())}


def compute_stats_from_draws(draws: Mapping[str, Sequence[float]]) -> dict[str, dict[str, float]]:
    """Per-parameter mean/std of candidate draws; parameters of equal length share one launch."""
    result: dict[str, dict[str, float]] = {}
    by_len: dict[int, list[str]] = {}
    for param, values in draws.items():
        by_len.setdefault(len(values), []).append(param)
    for n, params in by_len.items():
        if n == 0:
            for p in params:
                result[p] = {"mean": float("nan"), "std": float("nan")}
            continue
        x = np.stack([np.asarray(draws[p], dtype=np.float64) for p in params])
        finite = np.isfinite(x).all(axis=1)
        for i in np.flatnonzero(~finite):  # one bad parameter fails its own compare, not the whole batch
            result[params[i]] = _nonfinite_stats(x[i])
        good = np.flatnonzero(finite)
        if good.size:
            out = batch.stats_batch(x[good] if good.size < len(params) else x, quantiles=())
            mean = out["mean"].cpu().numpy()
            std = out["std"].cpu().numpy()
            for j, i in enumerate(good):
                result[params[i]] = {"mean": float(mean[j]), "std": float(std[j])}
    return {p: result[p] for p in draws}
