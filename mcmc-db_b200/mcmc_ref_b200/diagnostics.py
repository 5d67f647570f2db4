"""Rank-normalised split R-hat and ESS -- same names, signatures, guards and error
messages as the reference's ``mcmc_ref.diagnostics`` (diagnostics.py:13-73), computed
on the GPU.  The scalar forms call the batched kernels with P = 1.

Ragged chains follow the reference's rules: ranks and the fold median are taken over ALL
draws of all chains (``_rank_normalize`` / ``_fold_chains``, diagnostics.py:93-133), each chain is
split at its own midpoint (:76-85), and the chain statistics truncate every (split) chain to the
shortest one (``_rhat`` :140, ``_ess`` :158).

Documented divergence from the reference: non-finite draws raise ``ValueError`` (the reference's
result on NaN depends on the sort order of NaN and is not meaningful).
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
    """[C, N] float64, or None if there are no draws at all (equal-length chains only)."""
    n = len(chains[0]) if len(chains) else 0
    if n == 0:
        return None
    return np.asarray(chains, dtype=np.float64).reshape(len(chains), n)


def _diag_ragged(chains) -> dict[str, float]:
    """Chains of different lengths, by the reference's rules (see the module docstring): one pooled ranking of all
    draws on the device, then the chain kernels on the truncated layouts the reference's `_rhat` / `_ess` see."""
    nan = math.nan
    lens = [len(c) for c in chains]
    flat = np.concatenate([np.asarray(c, dtype=np.float64).reshape(-1) for c in chains])
    if flat.size == 0:
        return {"rhat": nan, "ess_bulk": nan, "ess_tail": nan}
    r = batch.ranks_batch(flat.reshape(1, 1, -1))
    zb = r["z"].reshape(-1)
    zf = r["z_folded"].reshape(-1)
    off = np.concatenate([[0], np.cumsum(lens)])
    out = {"rhat": nan, "ess_bulk": nan, "ess_tail": nan}
    # ESS: every chain cut to the shortest one (diagnostics.py:158-162); a chain without draws makes n = 0 -> nan
    n = min(lens)
    if n >= 2:
        rows = np.concatenate([np.arange(o, o + n) for o in off[:-1]])
        idx = batch.to_index(rows)
        st = batch.chain_stats_batch(zb[idx].reshape(1, len(lens), n), zf[idx].reshape(1, len(lens), n))
        out["ess_bulk"] = float(st["ess_bulk"][0].item())
        out["ess_tail"] = float(st["ess_tail"][0].item())
    # R-hat: halves of each chain at ITS midpoint, chains with fewer than two draws dropped (:76-85), every half cut
    # to the shortest half (:140).  Laid out as pseudo-chains [first half | second half] of length 2 h so that the
    # kernel's own split recovers exactly those 2 m sequences.
    halves = [(o, ln // 2) for o, ln in zip(off[:-1], lens) if ln // 2 > 0]
    if len(halves) >= 1:
        h = min(hc for _, hc in halves)
        if h >= 2:
            rows = np.concatenate([np.concatenate([np.arange(o, o + h), np.arange(o + hc, o + hc + h)]) for o, hc in halves])
            idx = batch.to_index(rows)
            st = batch.chain_stats_batch(zb[idx].reshape(1, len(halves), 2 * h), zf[idx].reshape(1, len(halves), 2 * h))
            out["rhat"] = float(st["rhat"][0].item())
    return out


def _guard(chains, min_chains: int, what: str) -> bool:
    _validate_min_chains(min_chains)
    if len(chains) < min_chains:
        raise ValueError(
            f"{what} diagnostics require at least {min_chains} chains; got {len(chains)} chain(s)"
        )
    return len(chains) < 2


def _diag(chains) -> dict[str, float] | None:
    if len({len(c) for c in chains}) > 1:
        return _diag_ragged(chains)
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
