"""Golden fixture of BASELINE config 1, produced by running the UNMODIFIED reference.

    python tests/golden/make_c1_golden.py        (build container: /root/reference present)

Scenario: candidate = chains 0-3 of the packaged model
`eight_schools-eight_schools_noncentered` (10 params x 4 chains x 1000 draws), reference = all 10
packaged chains.  Records what `mcmc_ref.reference.compare` returns (numpy and arrow backends,
reference.py:107-122), the candidate's `split_rhat / ess_bulk / ess_tail` (diagnostics.py:13-73) and
`NumpyBackend.stats` of the candidate (backends_numpy.py:17-49).  The draws themselves travel in
`packaged_subset.npz`; this file only stores the reference's answers -> `c1_compare.json`.
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import pyarrow as pa

REF = Path("/root/reference")
sys.path.insert(0, str(REF / "src"))
sys.path.insert(0, str(REF / "packages/mcmc-ref-data/src"))

from mcmc_ref import diagnostics as ref_diag  # noqa: E402
from mcmc_ref import reference as ref_api  # noqa: E402
from mcmc_ref.backends_numpy import NumpyBackend  # noqa: E402

HERE = Path(__file__).resolve().parent
MODEL = "eight_schools-eight_schools_noncentered"


def main() -> None:
    table = ref_api.draws(MODEL, return_="arrow")
    table = table.read_all() if hasattr(table, "read_all") else table
    params = [c for c in table.column_names if c not in {"chain", "draw"}]
    chain = np.asarray(table.column("chain").to_numpy())
    keep = np.isin(chain, [0, 1, 2, 3])
    actual = {p: np.asarray(table.column(p).to_numpy())[keep].tolist() for p in params}
    out = {"model": MODEL, "params": params, "candidate_chains": [0, 1, 2, 3], "tolerance": 0.15, "compare": {}}
    for backend in ("numpy", "arrow"):
        res = ref_api.compare(MODEL, actual, tolerance=0.15, backend=backend)
        out["compare"][backend] = {
            "passed": res.passed, "failures": res.failures,
            "details": {p: {m: vars(r) for m, r in d.items()} for p, d in res.details.items()},
        }
    # tight tolerance: the same inputs must FAIL with the reference's failure strings
    res = ref_api.compare(MODEL, actual, tolerance=0.005, backend="numpy")
    out["compare_tight"] = {"tolerance": 0.005, "passed": res.passed, "failures": res.failures}
    diag = {}
    for p in params:
        col = np.asarray(table.column(p).to_numpy())
        chains = [col[chain == c].tolist() for c in range(4)]
        diag[p] = {"rhat": ref_diag.split_rhat(chains), "ess_bulk": ref_diag.ess_bulk(chains),
                   "ess_tail": ref_diag.ess_tail(chains)}
    out["candidate_diagnostics"] = diag
    cand_table = pa.table({p: np.asarray(actual[p]) for p in params})
    out["candidate_numpy_stats"] = NumpyBackend().stats(cand_table, params)
    (HERE / "c1_compare.json").write_text(json.dumps(out, indent=1))
    print("wrote c1_compare.json:", out["compare"]["numpy"]["passed"], len(out["compare_tight"]["failures"]), "tight failures")


if __name__ == "__main__":
    main()
