"""Backend registry -- the reference's plugin seam (backends.py:14-55).

Same ``Backend`` protocol, ``BackendSpec`` and ``get_backend`` contract; the one
registered backend is ``"b200"``.  The reference's CPU backends ("arrow", "numpy")
are deliberately not registered here: this package has no CPU path.  To plug the GPU
backend into the reference itself, add the entry shown in INTEGRATION.md to the
reference's own ``BACKENDS`` dict.
"""

from __future__ import annotations

from collections.abc import Callable, Iterable
from dataclasses import dataclass
from typing import Any, Protocol


class Backend(Protocol):
    name: str

    def stats(
        self,
        table: Any,
        params: Iterable[str],
        quantiles: Iterable[float] = (0.05, 0.5, 0.95),
        quantile_mode: str = "exact",
    ) -> dict[str, dict[str, float]]:
        """Compute per-parameter stats from an Arrow Table or batches."""


@dataclass(frozen=True)
class BackendSpec:
    name: str
    loader: Callable[[], Backend]


def _load_b200() -> Backend:
    from .backend_b200 import B200Backend  # lazy import

    return B200Backend()


BACKENDS: dict[str, BackendSpec] = {
    "b200": BackendSpec(name="b200", loader=_load_b200),
}


def get_backend(name: str) -> Backend:
    spec = BACKENDS.get(name)
    if spec is None:
        raise ValueError(f"Unknown backend: {name}")
    return spec.loader()
