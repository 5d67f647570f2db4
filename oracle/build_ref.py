"""Build `oracle/_ref/`: the UNMODIFIED reference package, for timing and cross-checking only.

TEST / BENCH INFRASTRUCTURE.  The reference (`mcmc-ref`) is pure Python, so "building" it is
making its package importable next to this file: `python oracle/build_ref.py` copies
`/root/reference/src/mcmc_ref/*.py` (no data, no scripts) to `oracle/_ref/mcmc_ref/`.  The
directory is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it
travels to the GPU box with the snapshot, where `bench.py --impl reference` times the reference's
own `split_rhat` / `ess_bulk` / `ess_tail` / `NumpyBackend.stats` / `compare_stats` on the host
cores (`"kind": "reference"`).  Nothing under `mcmc-db_b200/` imports it.
"""

from __future__ import annotations

import shutil
import sys
from pathlib import Path

SRC = Path("/root/reference/src/mcmc_ref")
DST = Path(__file__).resolve().parent / "_ref" / "mcmc_ref"
# the hot path and what it imports; the corpus-generation modules need cmdstanpy and stay behind
MODULES = ("__init__.py", "diagnostics.py", "compare.py", "backends.py", "backends_numpy.py", "backends_arrow.py",
           "reference.py", "store.py", "draws.py", "convert.py", "pairs.py", "provenance.py", "generate.py",
           "model_inventory.py", "cli.py")


def build(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref/mcmc_ref is in place (copied now or earlier)."""
    if not SRC.exists():
        if verbose:
            print(f"oracle/build_ref: {SRC} is not present (GPU box): using the prebuilt copy"
                  if DST.exists() else f"oracle/build_ref: {SRC} is not present and there is no prebuilt copy")
        return DST.exists()
    DST.mkdir(parents=True, exist_ok=True)
    for name in MODULES:
        if (SRC / name).exists():
            shutil.copyfile(SRC / name, DST / name)
    if verbose:
        print(f"oracle/build_ref: copied {len(list(DST.glob('*.py')))} modules to {DST}")
    return True


def import_reference():
    """The reference package from oracle/_ref (None when it was never built)."""
    root = str(DST.parent)
    if not (DST / "diagnostics.py").exists():
        return None
    if root not in sys.path:
        sys.path.insert(0, root)
    import mcmc_ref  # noqa: F401
    from mcmc_ref import compare, diagnostics  # noqa: F401
    from mcmc_ref.backends_numpy import NumpyBackend  # noqa: F401

    return sys.modules["mcmc_ref"]


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
