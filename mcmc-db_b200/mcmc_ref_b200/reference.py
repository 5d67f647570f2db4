"""Python API for the compare / diagnostics path -- mirror of ``mcmc_ref.reference``
(reference.py:30-48, :76-122) with ``backend="b200"`` as the default.

Data access goes through a ``store`` object with the reference's ``DataStore``
interface (``open_draws(model, params=...)`` returning something with ``.read_all()``,
``read_meta(model)``).  ``store.PackagedStore`` reads the reference's packaged layout
(``draws/<model>.draws.parquet``, ``meta/<model>.meta.json``) unchanged.
"""

from __future__ import annotations

from collections.abc import Mapping, Sequence

from .backends import get_backend
from .compare import compare_stats, compute_stats_from_draws
from .convert import _compute_diagnostics
from .store import PackagedStore


def list_models(store=None) -> list[str]:
    """reference.py:15-17"""
    return (store or PackagedStore()).list_models()


def draws(
    model: str,
    params: Sequence[str] | None = None,
    chains: Sequence[int] | None = None,
    return_: str = "arrow",
    store=None,
):
    """Reference draws of a model (reference.py:51-73).  ``return_``: ``"arrow"`` (table / reader),
    ``"numpy"`` ([rows, P], parameter fastest, like ``Draws.to_numpy`` draws.py:20-29), ``"list"`` (row dicts),
    and -- new -- ``"device"``: the CUDA tensor ``[P, C, N]`` the batched kernels consume, rows verified /
    ordered by (chain, draw) on the device (``ingest.table_to_device``)."""
    store = store or PackagedStore()
    reader = store.open_draws(model, params=params, chains=chains)
    if return_ == "arrow" and params is not None:
        return reader
    table = reader.read_all() if hasattr(reader, "read_all") else reader
    names = list(params) if params is not None else [c for c in table.column_names if c not in {"chain", "draw"}]
    if return_ == "arrow":
        return table
    if return_ == "numpy":
        import numpy as np

        return np.stack([table.column(p).to_numpy(zero_copy_only=False) for p in names], axis=-1)
    if return_ == "list":
        return table.to_pylist()
    if return_ == "device":
        from .ingest import table_to_device

        return table_to_device(table, names)[0]
    raise ValueError(f"Unknown return type: {return_}")


def stan_data(model: str, store=None) -> dict:
    """reference.py:20-22"""
    return (store or PackagedStore()).read_stan_data(model)


def model_code(model: str, store=None) -> str:
    """reference.py:25-27"""
    return (store or PackagedStore()).read_stan_code(model)


def log_density(model: str, draws: Mapping[str, Sequence[float]] | None = None, store=None):
    """EXTENSION (no reference counterpart): likelihood-block log-density of a packaged GLM-style model
    over draws, on the tensor cores -- see ``models.packaged_lp``."""
    from .models import packaged_lp

    return packaged_lp(model, draws=draws, store=store)


def stats(
    model: str,
    params: Sequence[str] | None = None,
    backend: str = "b200",
    quantile_mode: str = "exact",
    store=None,
) -> dict[str, dict[str, float]]:
    store = store or PackagedStore()
    reader = store.open_draws(model, params=params)
    table = reader.read_all() if hasattr(reader, "read_all") else reader
    if params is None:
        params = [c for c in table.column_names if c not in {"chain", "draw"}]
    backend_impl = get_backend(backend)
    return backend_impl.stats(table, params, quantile_mode=quantile_mode)


def diagnostics_for_model(
    model: str,
    params: Sequence[str] | None = None,
    store=None,
    recompute: bool = False,
) -> dict[str, dict[str, float]]:
    store = store or PackagedStore()
    if not recompute:
        try:
            meta = store.read_meta(model)
        except FileNotFoundError:
            meta = {}
        diag = meta.get("diagnostics")
        if isinstance(diag, dict) and diag:
            if params is None:
                return diag
            return {p: diag[p] for p in params if p in diag}
    reader = store.open_draws(model, params=params)
    table = reader.read_all() if hasattr(reader, "read_all") else reader
    if params is None:
        params = [c for c in table.column_names if c not in {"chain", "draw"}]
    return _compute_diagnostics(table, params)


def compare(
    model: str,
    actual: Mapping[str, Sequence[float]],
    tolerance: float = 0.15,
    metrics: Sequence[str] = ("mean", "std"),
    backend: str = "b200",
    store=None,
):
    ref_stats = stats(model, params=list(actual.keys()), backend=backend, store=store)
    actual_stats = compute_stats_from_draws(actual)
    return compare_stats(ref_stats, actual_stats, tolerance=tolerance, metrics=metrics)
