"""The reference's ``mcmc_ref.convert`` quality gate on the B200 path:
``convert_file`` (convert.py:26-67) with ``_ensure_chain_draw`` (:105-120),
``_count_chains_draws`` (:123-131), ``_chains_from_table`` (:150-161),
``_compute_diagnostics`` (:134-147), ``_checks`` (:164-172), ``_enforce_checks`` (:175-178).
The diagnostics of all parameters come from ONE batched device pass; the written
``meta.json`` has the reference's keys and (to rounding) its values.

``chains_tensor_from_table`` replaces the per-parameter ``to_pylist`` + bucket + sort
of the reference with one (chain, draw) ordering shared by all parameters: an Arrow
parameter column whose rows are chain-major is already the device layout ``x[p][C][N]``.
"""

from __future__ import annotations

from collections.abc import Iterable

import numpy as np

import json
import zipfile
from dataclasses import dataclass
from datetime import date
from pathlib import Path

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
    from . import ingest

    x, c, _n = ingest.table_to_device(table, params)  # device-side order check / gather
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


@dataclass(frozen=True)
class ConvertResult:
    draws_path: Path | None
    meta_path: Path | None
    meta: dict


def _ensure_chain_draw(table):
    """Add the missing ``chain`` (all 0) / ``draw`` (row number) columns (convert.py:105-120)."""
    import pyarrow as pa

    names = set(table.column_names)
    rows = table.num_rows
    if "draw" not in names:
        table = table.append_column("draw", pa.array(np.arange(rows, dtype=np.int32)))
    if "chain" not in names:
        table = table.append_column("chain", pa.array(np.zeros(rows, dtype=np.int32)))
    return table


def _count_chains_draws(table) -> tuple[int, int]:
    """(number of chains, shortest chain length) (convert.py:123-131)."""
    chain = column_to_numpy_int(table, "chain")
    if chain.size == 0:
        return 0, 0
    _, counts = np.unique(chain, return_counts=True)
    return int(counts.size), int(counts.min())


def _read_json_zip(path: Path):
    """Chain-list json-zip: ``[ {param: [draws...]}, ... ]``, one dict per chain (convert.py:78-102)."""
    import pyarrow as pa

    with zipfile.ZipFile(path) as zf:
        payload = json.loads(zf.read(zf.namelist()[0]))
    if not isinstance(payload, list) or not payload:
        raise ValueError("json-zip payload must be a non-empty list of chains")
    params = sorted(payload[0].keys())
    n = len(next(iter(payload[0].values())))
    cols = {"chain": np.repeat(np.arange(len(payload)), n), "draw": np.tile(np.arange(n), len(payload))}
    for name in params:
        cols[name] = np.concatenate([np.asarray(ch[name][:n], dtype=np.float64) for ch in payload])
    return pa.table(cols)


def _read_input(path: Path):
    import pyarrow.csv as pacsv
    import pyarrow.parquet as pq

    if path.suffix == ".csv":
        return pacsv.read_csv(path)
    if path.suffix == ".parquet":
        return pq.read_table(path)
    if path.suffixes[-2:] == [".json", ".zip"]:
        return _read_json_zip(path)
    raise ValueError(f"Unsupported input format: {path}")


def convert_table(table, name: str, out_draws_dir: Path | None = None, out_meta_dir: Path | None = None,
                  force: bool = False, source: str = "converted") -> ConvertResult:
    """The quality gate of ``convert_file`` on an in-memory table; writes Parquet + meta when dirs are given."""
    table = _ensure_chain_draw(table)
    params = [c for c in table.column_names if c not in {"chain", "draw"}]
    n_chains, n_draws = _count_chains_draws(table)
    diag = _compute_diagnostics(table, params, min_chains=1 if force else 4)
    checks = _checks(n_chains, n_draws, diag)
    if not force:
        _enforce_checks(checks)
    meta = {
        "model": name,
        "parameters": params,
        "n_chains": n_chains,
        "n_draws_per_chain": n_draws,
        "diagnostics": diag,
        "generated_date": date.today().isoformat(),
        "checks": checks,
        "source": source,
    }
    draws_path = meta_path = None
    if out_draws_dir is not None and out_meta_dir is not None:
        import pyarrow.parquet as pq

        draws_path = Path(out_draws_dir) / f"{name}.draws.parquet"
        meta_path = Path(out_meta_dir) / f"{name}.meta.json"
        pq.write_table(table, draws_path)
        meta_path.write_text(json.dumps(meta, indent=2, sort_keys=True))
    return ConvertResult(draws_path=draws_path, meta_path=meta_path, meta=meta)


def convert_file(input_path: Path, name: str, out_draws_dir: Path, out_meta_dir: Path, force: bool = False,
                 source: str = "converted") -> ConvertResult:
    return convert_table(_read_input(Path(input_path)), name, Path(out_draws_dir), Path(out_meta_dir),
                         force=force, source=source)
