"""``B200Backend.stats`` -- GPU counterpart of ``NumpyBackend.stats`` (backends_numpy.py:17-49).

Accepts what the reference's backends accept: an Arrow ``Table``, anything with
``.read_all()``, or a mapping of column name -> array.  Returns the same dict:
``{param: {"mean", "std", f"q{int(q*100)}"...}}``.  Quantiles are numpy-linear, the
selected order statistics are bit-exact.
"""

from __future__ import annotations

from collections.abc import Iterable

import numpy as np

from . import batch


def column_to_numpy(table, param: str) -> np.ndarray:
    if hasattr(table, "column"):
        col = table.column(param)
        if hasattr(col, "to_numpy"):
            return np.asarray(col.to_numpy(zero_copy_only=False), dtype=np.float64)
        return np.asarray(col, dtype=np.float64)
    return np.asarray(table[param], dtype=np.float64)


class B200Backend:
    name = "b200"

    def __init__(self) -> None:
        from . import _native

        _native.load()  # fail loudly at construction, as the reference's backends do on ImportError

    def stats(
        self,
        table,
        params: Iterable[str],
        quantiles: Iterable[float] = (0.05, 0.5, 0.95),
        quantile_mode: str = "exact",
    ) -> dict[str, dict[str, float]]:
        if hasattr(table, "read_all"):
            table = table.read_all()
        params = list(params)
        qs = list(quantiles)
        results: dict[str, dict[str, float]] = {}
        if not params:
            return results
        cols = [column_to_numpy(table, p) for p in params]
        # all columns of one table have the same number of rows -> one [P, M] launch
        x = np.stack(cols)
        out = batch.stats_batch(x, quantiles=qs)
        mean = out["mean"].cpu().numpy()
        std = out["std"].cpu().numpy()
        quant = out["quantiles"].cpu().numpy()
        for i, p in enumerate(params):
            entry = {"mean": float(mean[i]), "std": float(std[i])}
            for j, q in enumerate(qs):
                entry[f"q{int(q * 100)}"] = float(quant[i, j])
            results[p] = entry
        return results
