"""Minimal reader of the reference's packaged data layout (store.py:21-168).

Resolution order as in the reference: packaged root (``mcmc_ref_data`` / ``mcmc_ref``
package data if installed, or an explicit ``packaged_root``) first, then the local
root (``MCMC_REF_LOCAL_ROOT`` or ``~/.mcmc-ref``).  Only what the compare /
diagnostics path needs: ``open_draws`` with column projection and chain filter,
``read_meta``, ``list_models``, and ``read_stan_data`` / ``read_stan_code`` for the log-density extension.
"""

from __future__ import annotations

import json
import os
from collections.abc import Sequence
from pathlib import Path


def _default_local_root() -> Path:
    env = os.environ.get("MCMC_REF_LOCAL_ROOT")
    return Path(env) if env else Path.home() / ".mcmc-ref"


def _default_packaged_root() -> Path | None:
    from importlib import resources

    for package_name in ("mcmc_ref_data", "mcmc_ref"):
        try:
            data = Path(str(resources.files(package_name).joinpath("data")))
        except Exception:
            continue
        if (data / "draws").exists() or (data / "meta").exists():
            return data
    return None


class PackagedStore:
    def __init__(self, local_root: Path | None = None, packaged_root: Path | None = None) -> None:
        roots = [packaged_root or _default_packaged_root(), local_root or _default_local_root()]
        self._roots = [Path(r) for r in roots if r is not None and Path(r).exists()]

    def list_models(self) -> list[str]:
        names: set[str] = set()
        for root in self._roots:
            for path in (root / "draws").glob("*.draws.parquet"):
                names.add(path.stem.replace(".draws", ""))
        return sorted(names)

    def _resolve(self, model: str, subdir: str, suffix: str) -> Path | None:
        for root in self._roots:
            path = root / subdir / f"{model}{suffix}"
            if path.exists():
                return path
        return None

    def resolve_draws_path(self, model: str) -> Path:
        path = self._resolve(model, "draws", ".draws.parquet")
        if path is None:
            raise FileNotFoundError(f"draws not found for model: {model}")
        return path

    def read_meta(self, model: str) -> dict:
        path = self._resolve(model, "meta", ".meta.json")
        if path is None:
            raise FileNotFoundError(f"metadata not found for model: {model}")
        return json.loads(path.read_text())

    def read_stan_data(self, model: str) -> dict:
        """``stan_data/<model>.data.json`` (store.py:57-66)."""
        path = self._resolve(model, "stan_data", ".data.json")
        if path is None:
            raise FileNotFoundError(f"stan data not found for model: {model}")
        return json.loads(path.read_text())

    def read_stan_code(self, model: str) -> str:
        """``stan_code/<model>.stan``, else ``stan_models/<model>.stan`` (store.py:68-77, :110-119)."""
        for sub in ("stan_code", "stan_models"):
            path = self._resolve(model, sub, ".stan")
            if path is not None:
                return path.read_text()
        raise FileNotFoundError(f"stan code not found for model: {model}")

    def open_draws(self, model: str, params: Sequence[str] | None = None, chains: Sequence[int] | None = None,
                   batch_size: int = 1024):
        import pyarrow.dataset as ds

        dataset = ds.dataset(self.resolve_draws_path(model), format="parquet")
        available = dataset.schema.names
        selected = [c for c in available if c not in {"chain", "draw"}] if params is None else list(params)
        filt = ds.field("chain").isin(list(chains)) if chains is not None else None
        scanner = dataset.scanner(columns=["chain", "draw", *selected], filter=filt, batch_size=batch_size)
        return scanner.to_reader()
