"""Rank-normalised split R-hat and ESS -- same names, signatures, guards and error
messages as the reference's ``mcmc_ref.diagnostics`` (diagnostics.py:13-73), computed
on the GPU.  The scalar forms call the batched kernels with P = 1.

Documented divergences from the reference:
  * non-finite draws raise ``ValueError`` (the reference's result on NaN depends on
    the sort order of NaN and is not meaningful);
  * chains must all have the same length (the packaged data layout is rectangular);
    ragged input raises ``ValueError``.
"""

from __future__ import annotations

import math
from collections.abc import Sequence

import numpy as np

from . import batch


def _validate_min_chains(min_chains: int) -> None:
    if min_chains < 1:
        raise ValueError(f"min_chains must be >= 1; got {min_chains}")


def _as_array(chains: Sequence[Sequence[float]]) -> np.ndarray | None:
    """[C, N] float64, or None if there are no draws at all."""
    lens = {len(c) for c in chains}
    if len(lens) > 1:
        raise ValueError("ragged chains are not supported by the b200 backend; pass equal-length chains")
    n = lens.pop() if lens else 0
    if n == 0:
        return None
    return np.asarray(chains, dtype=np.float64).reshape(len(chains), n)


def _guard(chains, min_chains: int, what: str) -> bool:
    _validate_min_chains(min_chains)
    if len(chains) < min_chains:
        raise ValueError(
            f"{what} diagnostics require at least {min_chains} chains; got {len(chains)} chain(s)"
        )
    return len(chains) < 2


def _diag(chains) -> dict[str, float] | None:
    x = _as_array(chains)
    if x is None:
        return None
    out = batch.diagnostics_batch(x[None, :, :], quantiles=())
    return {k: float(out[k][0].item()) for k in ("rhat", "ess_bulk", "ess_tail")}


def split_rhat(chains: Sequence[Sequence[float]], *, min_chains: int = 4) -> float:
    """Rank-normalized split R-hat with folded variant (returns max of both)."""
    if _guard(chains, min_chains, "R-hat"):
        return float("nan")
    d = _diag(chains)
    return math.nan if d is None else d["rhat"]


def ess_bulk(chains: Sequence[Sequence[float]], *, min_chains: int = 4) -> float:
    if _guard(chains, min_chains, "ESS"):
        return float("nan")
    d = _diag(chains)
    return math.nan if d is None else d["ess_bulk"]


def ess_tail(chains: Sequence[Sequence[float]], *, min_chains: int = 4) -> float:
    if _guard(chains, min_chains, "ESS"):
        return float("nan")
    d = _diag(chains)
    return math.nan if d is None else d["ess_tail"]


def diagnostics(chains: Sequence[Sequence[float]], *, min_chains: int = 4) -> dict[str, float]:
    """All three in one GPU pass (the reference computes them with six sorts)."""
    if _guard(chains, min_chains, "R-hat"):
        return {"rhat": math.nan, "ess_bulk": math.nan, "ess_tail": math.nan}
    d = _diag(chains)
    return d or {"rhat": math.nan, "ess_bulk": math.nan, "ess_tail": math.nan}
