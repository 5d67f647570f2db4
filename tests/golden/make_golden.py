"""Generate the golden fixtures by running the UNMODIFIED reference.

Run in the build container (where ``/root/reference`` is mounted):

    python tests/golden/make_golden.py

Writes, next to this file:
  ref_cases.json          small hand-made / seeded inputs with the reference's own
                          split_rhat / ess_bulk / ess_tail / _rank_normalize (z),
                          folded z, and the average ranks recovered from the
                          reference's code path.
  compare_cases.json      compare_stats / compute_basic_stats inputs and outputs
                          (verdicts, rel errors, failure strings).
  packaged_subset.npz     draws [P, C, N] of a few packaged models (chain-major)
  packaged_subset.json    their shipped meta.json diagnostics (the reference's
                          published golden values) + NumpyBackend.stats output.
  packaged_all_meta.json  shipped diagnostics of all packaged models that have
                          draws here (used only when /root/reference is present).

Nothing here runs on the GPU box; the fixtures travel, the reference does not.
"""

from __future__ import annotations

import json
import math
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
sys.path.insert(0, str(REF / "src"))

from mcmc_ref import compare as ref_compare  # noqa: E402
from mcmc_ref import diagnostics as ref_diag  # noqa: E402
from mcmc_ref.backends_numpy import NumpyBackend  # noqa: E402
from mcmc_ref.convert import _chains_from_table  # noqa: E402

HERE = Path(__file__).resolve().parent
DATA = REF / "packages/mcmc-ref-data/src/mcmc_ref_data/data"


def _enc(v: float):
    if isinstance(v, float):
        if math.isnan(v):
            return "nan"
        if math.isinf(v):
            return "inf" if v > 0 else "-inf"
    return v


def _ref_ranks(chains):
    """Average ranks via the reference's own loop: swap inv_cdf for identity."""

    class _Ident:
        def inv_cdf(self, p):
            return p

    saved = ref_diag.NormalDist
    ref_diag.NormalDist = _Ident
    try:
        pz = ref_diag._rank_normalize(chains)
    finally:
        ref_diag.NormalDist = saved
    n = sum(len(c) for c in chains)
    out = []
    for row in pz:
        r = [round((p * n + 0.5) * 2) / 2 for p in row]
        out.append(r)
    return out


def _case(name, chains, min_chains=4):
    chains = [[float(v) for v in c] for c in chains]
    rec = {"name": name, "chains": chains, "min_chains": min_chains}
    try:
        rec["rhat"] = _enc(ref_diag.split_rhat(chains, min_chains=min_chains))
        rec["ess_bulk"] = _enc(ref_diag.ess_bulk(chains, min_chains=min_chains))
        rec["ess_tail"] = _enc(ref_diag.ess_tail(chains, min_chains=min_chains))
    except ValueError as exc:
        rec["error"] = str(exc)
        return rec
    if sum(len(c) for c in chains):
        rec["z"] = ref_diag._rank_normalize(chains)
        folded = ref_diag._fold_chains(chains)
        rec["folded"] = folded
        rec["z_folded"] = ref_diag._rank_normalize(folded)
        rec["avg_rank"] = _ref_ranks(chains)
        rec["avg_rank_folded"] = _ref_ranks(folded)
    return rec


def _ar1(rng, c, n, phi, mu=0.0, sigma=1.0):
    e = rng.standard_normal((c, n))
    x = np.empty((c, n))
    x[:, 0] = e[:, 0]
    s = math.sqrt(1 - phi * phi)
    for i in range(1, n):
        x[:, i] = phi * x[:, i - 1] + s * e[:, i]
    return mu + sigma * x


def make_ref_cases():
    rng = np.random.Generator(np.random.Philox(key=4711))
    cases = []
    a = [[1.0, 2.0, 3.0, 4.0], [1.1, 2.1, 3.1, 4.1], [0.9, 1.9, 2.9, 3.9], [1.05, 2.05, 3.05, 4.05]]
    cases.append(_case("unit_ess_positive", a))
    cases.append(_case("unit_identical", [[1.0] * 4] * 4))
    cases.append(_case("unit_scale_diff", [[0.0] * 4, [10.0] * 4, [0.0] * 4, [10.0] * 4]))
    cases.append(_case("unit_two_chains_error", [[1.0, 2.0, 3.0, 4.0], [1.1, 2.1, 3.1, 4.1]]))
    cases.append(_case("unit_single_chain_override", [[1.0, 2.0, 3.0, 4.0]], min_chains=1))
    cases.append(_case("cli_4x2", [[1.0, 2.0], [1.5, 2.5], [1.0, 2.0], [1.5, 2.5]]))
    cases.append(_case("two_chains_min1", _ar1(rng, 2, 40, 0.3).tolist(), min_chains=1))
    cases.append(_case("normal_4x50", rng.standard_normal((4, 50)).tolist()))
    cases.append(_case("normal_6x257_odd", (3.0 + 0.5 * rng.standard_normal((6, 257))).tolist()))
    cases.append(_case("ties_4x64", np.round(rng.standard_normal((4, 64)), 1).tolist()))
    cases.append(_case("ties_heavy_5x40", rng.integers(0, 4, (5, 40)).astype(float).tolist()))
    cases.append(_case("signed_zero", [[0.0, -0.0, 1.0, -1.0, 0.0, 2.0], [-0.0, 0.0, -2.0, 0.5, -0.5, 0.0],
                                       [1.0, -1.0, 0.0, -0.0, 3.0, -3.0], [0.25, -0.25, -0.0, 0.0, 0.0, 1.5]]))
    cases.append(_case("ar95_4x300", _ar1(rng, 4, 300, 0.95).tolist()))
    cases.append(_case("ar90_8x500", _ar1(rng, 8, 500, 0.9, mu=-3.0, sigma=2.0).tolist()))
    cases.append(_case("ar50_16x125", _ar1(rng, 16, 125, 0.5, mu=7.0, sigma=0.25).tolist()))
    cases.append(_case("cauchy_4x101", rng.standard_cauchy((4, 101)).tolist()))
    x = _ar1(rng, 4, 120, 0.2)
    x[1] += 1.5
    x[3] += 1.5
    cases.append(_case("shifted_chains_4x120", x.tolist()))
    x = rng.standard_normal((4, 30))
    x[2] = 0.75
    cases.append(_case("one_constant_chain", x.tolist()))
    cases.append(_case("n3_half1", rng.standard_normal((4, 3)).tolist()))
    cases.append(_case("n1", rng.standard_normal((4, 1)).tolist()))
    cases.append(_case("n5", rng.standard_normal((4, 5)).tolist()))
    cases.append(_case("antithetic_4x80", _ar1(rng, 4, 80, -0.6).tolist()))
    cases.append(_case("huge_range", (10.0 ** rng.uniform(-200, 200, (4, 32)) * rng.choice([-1, 1], (4, 32))).tolist()))
    cases.append(_case("exp_skew_10x100", rng.exponential(2.0, (10, 100)).tolist()))
    cases.append(_case("constant_chains_4x30", [[0.0] * 30, [1.0] * 30, [0.0] * 30, [1.0] * 30]))
    cases.append(_case("constant_chains_4x1500", [[0.0] * 1500, [0.0] * 1500, [1.0] * 1500, [1.0] * 1500]))
    cases.append(_case("constant_chains_5x7", [[0.3] * 7, [0.1] * 7, [0.7] * 7, [0.1] * 7, [0.3] * 7]))
    (HERE / "ref_cases.json").write_text(json.dumps(cases))
    return len(cases)


def make_compare_cases():
    rng = np.random.Generator(np.random.Philox(key=4712))
    out = {"basic": [], "compare": []}
    for vals in ([1.0, 2.0, 1.5, 2.5], [], [3.0], rng.standard_normal(257).tolist(),
                 (1e6 + rng.standard_normal(64)).tolist()):
        st = ref_compare.compute_basic_stats(vals)
        out["basic"].append({"values": vals, "mean": _enc(st["mean"]), "std": _enc(st["std"])})
    ref = {"mu": {"mean": 1.0, "std": 1.0}, "tau": {"mean": 0.0, "std": 2.0},
           "theta[1]": {"mean": -3.5, "std": 0.1}, "nanref": {"mean": float("nan"), "std": 1.0},
           "gone": {"mean": 1.0, "std": 1.0}, "edge": {"mean": 2.0, "std": 4.0}}
    act = {"mu": {"mean": 1.05, "std": 0.95}, "tau": {"mean": 1e-13, "std": 2.5},
           "theta[1]": {"mean": -3.5, "std": 0.2}, "nanref": {"mean": 1.0, "std": 1.0},
           "edge": {"mean": 2.3, "std": 4.6}}
    for tol, metrics in ((0.1, ["mean", "std"]), (0.15, ["mean", "std"]), (0.15, ["mean"]),
                         (0.15, ["mean", "std", "q50"])):
        res = ref_compare.compare_stats(ref, act, tolerance=tol, metrics=metrics)
        out["compare"].append({
            "ref": {k: {m: _enc(v) for m, v in d.items()} for k, d in ref.items()},
            "act": act, "tolerance": tol, "metrics": metrics, "passed": res.passed,
            "failures": res.failures,
            "details": {p: {m: [_enc(r.ref), _enc(r.actual), _enc(r.rel_error), r.passed]
                            for m, r in d.items()} for p, d in res.details.items()},
        })
    (HERE / "compare_cases.json").write_text(json.dumps(out))


SUBSET = ["eight_schools-eight_schools_noncentered", "wells_data-wells_dist", "sblri-blr",
          "kidiq-kidscore_momiq", "arK-arK", "garch-garch11"]


def make_packaged():
    import pyarrow.parquet as pq

    arrays = {}
    meta_out = {}
    all_meta = {}
    backend = NumpyBackend()
    for path in sorted((DATA / "draws").glob("*.draws.parquet")):
        model = path.name.replace(".draws.parquet", "")
        meta = json.loads((DATA / "meta" / f"{model}.meta.json").read_text())
        diag = meta.get("diagnostics") or {}
        all_meta[model] = {"n_chains": meta.get("n_chains"), "n_draws_per_chain": meta.get("n_draws_per_chain"),
                           "diagnostics": diag, "checks": meta.get("checks")}
        if model not in SUBSET:
            continue
        table = pq.read_table(path)
        params = [c for c in table.column_names if c not in {"chain", "draw"}]
        cols = []
        for p in params:
            chains = _chains_from_table(table, p)
            cols.append(np.array(chains, dtype=np.float64))
        arr = np.stack(cols)  # [P, C, N]
        arrays[model] = arr
        stats = backend.stats(table, params)
        meta_out[model] = {"params": params, "diagnostics": {p: diag[p] for p in params},
                           "numpy_stats": stats, "checks": meta.get("checks"),
                           "n_chains": meta.get("n_chains"), "n_draws_per_chain": meta.get("n_draws_per_chain")}
    np.savez_compressed(HERE / "packaged_subset.npz", **arrays)
    (HERE / "packaged_subset.json").write_text(json.dumps(meta_out))
    (HERE / "packaged_all_meta.json").write_text(json.dumps(all_meta))
    return {m: a.shape for m, a in arrays.items()}


if __name__ == "__main__":
    print("ref cases:", make_ref_cases())
    make_compare_cases()
    print("packaged:", make_packaged())
