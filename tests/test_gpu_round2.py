"""GPU parity tests added in round 2 (all through the Python mirror -> ctypes -> C-ABI):

* BASELINE config 1 as a compare SCENARIO: chains 0-3 of the packaged eight_schools model against all ten packaged
  chains through `reference.compare` (fixture `c1_compare.json`, produced by the unmodified reference).
* the reference's arrow backend as a cross-check of `B200Backend.stats` (tests/unit/test_backends_consistency.py:17-28).
* value parity of the long-chain paths (wide partition, binned score scatter, oversize-bucket list overflow).
* error behaviour: quantile levels outside [0, 1], non-finite candidate draws in a compare, 4096 chains.
"""

from __future__ import annotations

import json
import math

import numpy as np
import pytest

from conftest import GOLDEN, close
from oracle import mcmc_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.fixture(scope="module")
def B():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import __graft_entry__ as g

    g.build()
    from mcmc_ref_b200 import batch

    return batch


def _write_packaged(root, model, draws, params):
    """The reference's packaged layout (draws/<model>.draws.parquet: chain, draw int64 + one double column per param)."""
    import pyarrow as pa
    import pyarrow.parquet as pq

    p, c, n = draws.shape
    cols = {"chain": pa.array(np.repeat(np.arange(c), n), type=pa.int64()),
            "draw": pa.array(np.tile(np.arange(n), c), type=pa.int64())}
    for i, name in enumerate(params):
        cols[name] = pa.array(draws[i].reshape(-1), type=pa.float64())
    (root / "draws").mkdir(parents=True, exist_ok=True)
    pq.write_table(pa.table(cols), root / "draws" / f"{model}.draws.parquet")


def test_c1_compare_scenario(B, packaged_subset, tmp_path):
    """BASELINE configs[0]: 10 params x 4 chains x 1000 candidate draws vs the packaged reference posterior."""
    from mcmc_ref_b200 import reference
    from mcmc_ref_b200.store import PackagedStore

    gold = json.loads((GOLDEN / "c1_compare.json").read_text())
    arrays, meta = packaged_subset
    model = gold["model"]
    params = meta[model]["params"]
    assert params == gold["params"]
    draws = arrays[model]  # [10 params, 10 chains, 1000 draws]
    _write_packaged(tmp_path, model, draws, params)
    store = PackagedStore(local_root=tmp_path / "none", packaged_root=tmp_path)
    actual = {p: draws[i, :4].reshape(-1).tolist() for i, p in enumerate(params)}
    res = reference.compare(model, actual, tolerance=gold["tolerance"], store=store)
    want = gold["compare"]["numpy"]
    assert res.passed is want["passed"] is True and res.failures == want["failures"]
    for p in params:
        for m in ("mean", "std"):
            g, w = res.details[p][m], want["details"][p][m]
            assert g.passed is w["passed"]
            for k in ("ref", "actual", "rel_error"):
                assert close(getattr(g, k), w[k], 1e-12), (p, m, k, getattr(g, k), w[k])
    # the arrow backend of the reference gives the same verdicts (its quantiles differ by an ulp, mean / std do not)
    assert gold["compare"]["arrow"]["passed"] is True
    # same inputs at a tolerance they cannot meet: the reference's failure strings, in its order
    tight = reference.compare(model, actual, tolerance=gold["compare_tight"]["tolerance"], store=store)
    assert tight.passed is False and tight.failures == gold["compare_tight"]["failures"]
    # the candidate's own diagnostics and numpy statistics
    d = B.diagnostics_batch(draws[:, :4])
    for i, p in enumerate(params):
        for k in ("rhat", "ess_bulk", "ess_tail"):
            assert close(float(d[k][i].item()), gold["candidate_diagnostics"][p][k], TOL), (p, k)
        ns = gold["candidate_numpy_stats"][p]
        assert close(float(d["mean"][i].item()), ns["mean"], 1e-12) and close(float(d["std"][i].item()), ns["std"], 1e-12)
        for j, q in enumerate(("q5", "q50", "q95")):
            assert float(d["quantiles"][i, j].item()) == ns[q], (p, q)


def test_arrow_backend_cross_check(B, packaged_subset):
    """`ArrowBackend.stats` (backends_arrow.py:36-51) vs the b200 backend on the same table: mean / std to 1e-10
    relative as the reference's own consistency test asks (test_backends_consistency.py:17-28); the linear quantiles
    of pyarrow differ from numpy's `_lerp` by at most one ulp."""
    import pyarrow as pa
    import pyarrow.compute as pc

    from mcmc_ref_b200.backends import get_backend

    arrays, meta = packaged_subset
    rng = np.random.default_rng(17)
    tables = {m: (arrays[m].reshape(arrays[m].shape[0], -1), meta[m]["params"]) for m in arrays.files}
    tables["synthetic_100"] = (rng.standard_normal((1, 100)), ["x"])  # the reference test's own size
    be = get_backend("b200")
    for name, (x, params) in tables.items():
        table = pa.table({p: x[i] for i, p in enumerate(params)})
        got = be.stats(table, params)
        for i, p in enumerate(params):
            col = table.column(p)
            a_mean = pc.mean(col).as_py()
            a_std = pc.stddev(col, ddof=0).as_py()
            a_q = pc.quantile(col, q=[0.05, 0.5, 0.95], interpolation="linear", skip_nulls=True).to_pylist()
            assert abs(got[p]["mean"] - a_mean) <= 1e-10 * max(abs(a_mean), 1e-300), (name, p)
            assert abs(got[p]["std"] - a_std) <= 1e-10 * max(abs(a_std), 1e-300), (name, p)
            for key, aq in zip(("q5", "q50", "q95"), a_q):
                ulp = math.ulp(max(abs(aq), abs(got[p][key])))
                assert abs(got[p][key] - aq) <= ulp, (name, p, key, got[p][key], aq)


def _check_vs_oracle(B, x, label, lags=True):
    d = B.diagnostics_batch(x[None])
    want = O.diagnostics_param(x)
    for k in ("rhat", "ess_bulk", "ess_tail"):
        assert close(float(d[k][0].item()), want[k], TOL), (label, k, float(d[k][0].item()), want[k])
    if lags:
        assert int(d["lag_bulk"][0].item()) == want["lag_bulk"] and int(d["lag_tail"][0].item()) == want["lag_tail"], label
    ns = O.numpy_stats(x.reshape(-1))
    for j, q in enumerate(("q5", "q50", "q95")):
        assert float(d["quantiles"][0, j].item()) == ns[q], (label, q)
    assert close(float(d["mean"][0].item()), ns["mean"], 1e-12), label


def test_long_chain_binned_scatter_vs_oracle(B):
    """M > 3M pooled draws in diagnostics mode: wide partition + two-step (binned) score scatter.  Folded scores,
    tail R-hat and tail ESS of that path against the oracle -- continuous, and with ties (values rounded)."""
    rng = np.random.default_rng(2024)
    c, n = 4, 1_000_000
    from scipy.signal import lfilter

    phi = 0.5
    x = lfilter([math.sqrt(1 - phi * phi)], [1.0, -phi], rng.standard_normal((c, n)), axis=1)  # AR(1) chains
    _check_vs_oracle(B, x, "binned_continuous")
    _check_vs_oracle(B, np.round(x, 3), "binned_tied")


def test_oversize_list_overflow_falls_back_cleanly(B):
    """More oversize sub-buckets than the list holds (600 values repeated ~5000 times on the wide path): k_finish_big
    must read only the stored entries and the parameter must come out of the LSD fallback with exact results."""
    rng = np.random.default_rng(7)
    x = rng.integers(0, 600, (4, 750_000)).astype(np.float64)
    _check_vs_oracle(B, x, "600_values_x_5000")


def test_average_bucket_over_capacity(B):
    """~17M pooled draws in one parameter: the average sub-bucket exceeds the on-chip capacity."""
    rng = np.random.default_rng(8)
    x = rng.standard_normal((4, 4_300_000))
    d = B.diagnostics_batch(x[None])
    flat = np.sort(x.reshape(-1))
    ns = O.numpy_stats(x.reshape(-1))
    for j, q in enumerate(("q5", "q50", "q95")):
        assert float(d["quantiles"][0, j].item()) == ns[q], q
    m = flat.size
    assert 0.999 < float(d["rhat"][0].item()) < 1.001
    assert 0.5 * m < float(d["ess_bulk"][0].item()) <= m and 0.5 * m < float(d["ess_tail"][0].item()) <= m
    # rank-normalised R-hat of the bulk against the oracle on the same data (one pooled sort on the host)
    want = O.diagnostics_param(x)
    for k in ("rhat", "ess_bulk", "ess_tail"):
        assert close(float(d[k][0].item()), want[k], TOL), (k, float(d[k][0].item()), want[k])


def test_4096_chains(B):
    """The advertised chain limit: k_decide0 keeps 3 doubles per chain in (opted-in) dynamic shared memory."""
    rng = np.random.default_rng(9)
    x = rng.standard_normal((4096, 8))
    d = B.diagnostics_batch(x[None])
    want = O.diagnostics_param(x)
    for k in ("rhat", "ess_bulk", "ess_tail"):
        assert close(float(d[k][0].item()), want[k], TOL), (k, float(d[k][0].item()), want[k])
    with pytest.raises(ValueError, match="chains"):
        B.diagnostics_batch(rng.standard_normal((1, 4097, 2)))


def test_quantile_levels_validated(B):
    x = np.random.default_rng(1).standard_normal((2, 500))
    for bad in (-0.1, 1.5, float("nan")):
        with pytest.raises(ValueError, match=r"range \[0, 1\]"):
            B.stats_batch(x, quantiles=(0.5, bad))
        with pytest.raises(ValueError, match=r"range \[0, 1\]"):
            B.diagnostics_batch(x.reshape(2, 4, 125), quantiles=(bad,))
    out = B.stats_batch(x, quantiles=(0.0, 1.0))
    assert np.array_equal(out["quantiles"].cpu().numpy(), np.stack([x.min(axis=1), x.max(axis=1)], axis=1))


def test_compare_with_nonfinite_candidate_is_a_failed_verdict(B):
    """A diverged sampler (NaN / inf draws) must FAIL the compare with the reference's failure strings
    (compare.py:38-51, :58-64: mean nan or +-inf, std nan, `rel_error=nan > tol`), not raise -- and one bad parameter
    must not take the others with it."""
    from mcmc_ref_b200.compare import compare_stats, compute_basic_stats, compute_stats_from_draws

    rng = np.random.default_rng(3)
    good = rng.standard_normal(400).tolist()
    bad_nan = list(good)
    bad_nan[7] = float("nan")
    bad_inf = list(good)
    bad_inf[9] = float("inf")
    act = compute_stats_from_draws({"a": good, "b": bad_nan, "c": bad_inf})
    assert math.isfinite(act["a"]["mean"]) and math.isnan(act["b"]["mean"]) and math.isnan(act["b"]["std"])
    assert act["c"]["mean"] == math.inf and math.isnan(act["c"]["std"])
    ref = {k: {"mean": float(np.mean(good)), "std": float(np.std(good))} for k in "abc"}
    res = compare_stats(ref, act, 0.15, ("mean", "std"))
    assert res.passed is False
    assert res.failures == ["b.mean rel_error=nan > 0.15", "b.std rel_error=nan > 0.15",
                            "c.mean rel_error=inf > 0.15", "c.std rel_error=nan > 0.15"]
    assert res.details["a"]["mean"].passed and res.details["a"]["std"].passed
    s = compute_basic_stats(bad_nan)
    assert math.isnan(s["mean"]) and math.isnan(s["std"])


@pytest.mark.parametrize("m", [160_000, 262_144])
def test_z_table_bit_exact_everywhere(B, m):
    """All 2M - 1 entries of the device score table against `statistics.NormalDist().inv_cdf` (diagnostics.py:124-131),
    tails included (0 ulp): a permutation visits every odd code k = 2t + 1, pairs of equal values every even one."""
    from statistics import NormalDist

    nd = NormalDist()
    rng = np.random.default_rng(m)
    perm = rng.permutation(m).astype(np.float64)
    pairs = np.repeat(np.arange(m // 2, dtype=np.float64), 2)[rng.permutation(m)]
    r = B.ranks_batch(np.stack([perm, pairs]).reshape(2, 1, m), folded=False)
    z = r["z"].cpu().numpy().reshape(2, m)
    rank = r["avg_rank"].cpu().numpy().reshape(2, m)
    assert np.array_equal(rank[0], perm + 1.0) and np.array_equal(rank[1], 2.0 * pairs + 1.5)
    want_odd = np.array([nd.inv_cdf((t + 0.5) / m) for t in range(m)])          # untied draw at sorted position t
    want_even = np.array([nd.inv_cdf((2 * j + 1.0) / m) for j in range(m // 2)])  # run [2j, 2j + 2): k = 4j + 2
    assert np.array_equal(z[0], want_odd[perm.astype(np.int64)])
    assert np.array_equal(z[1], want_even[pairs.astype(np.int64)])


def test_numpy_layout_rows_param_fastest(B):
    """`Draws.to_numpy()` layout ([C * N, P], parameter fastest, draws.py:20-29) through the transposing load:
    the same bits as the canonical [P, C, N] tensor -- odd P, ragged tile edges, and a workspace that forces
    several parameter blocks."""
    rng = np.random.default_rng(23)
    for p, c, n in ((37, 4, 1001), (5, 6, 333), (64, 4, 2048)):
        x = rng.standard_normal((p, c, n)) * rng.uniform(0.5, 2.0, (p, 1, 1)) + rng.integers(-3, 3, (p, 1, 1))
        x[p // 2] = np.round(x[p // 2], 1)  # a tied parameter
        rows = np.ascontiguousarray(x.reshape(p, c * n).T)  # [C * N, P]
        want = B.diagnostics_batch(x)
        got = B.diagnostics_rows(rows, n_chains=c)
        for k, v in want.items():
            assert np.array_equal(v.cpu().numpy(), got[k].cpu().numpy(), equal_nan=True), (p, c, n, k)
    # a workspace that holds only a few parameters: several transposed blocks, same bits
    from mcmc_ref_b200 import _native

    lib = _native.load()
    x = rng.standard_normal((29, 4, 700))
    rows = np.ascontiguousarray(x.reshape(29, -1).T)
    want = B.diagnostics_batch(x)
    got = B.diagnostics_rows(rows, n_chains=4, workspace=B.Workspace(), ws_cap_bytes=int(lib.b200ref_rows_workspace_bytes(5, 4, 700)))
    for k, v in want.items():
        assert np.array_equal(v.cpu().numpy(), got[k].cpu().numpy(), equal_nan=True), ("blocks", k)
    with pytest.raises(ValueError, match="equal chains"):
        B.diagnostics_rows(np.zeros((10, 3)), n_chains=4)


def test_very_sticky_chains_reach_the_two_tile_rounds(B):
    """Autocorrelation that stays positive past lag 1041: the persistent rounds kernel switches from its single-tile
    to its two-tile variant in mid-flight (shared mbarrier phase, same work list).  Lags compared for equality."""
    from scipy.signal import lfilter

    rng = np.random.default_rng(123)
    c, n, phi = 4, 60_000, 0.9995
    x = lfilter([math.sqrt(1 - phi * phi)], [1.0, -phi], rng.standard_normal((c, n)), axis=1)
    d = B.diagnostics_batch(np.stack([x, rng.standard_normal((c, n))]))
    want = O.diagnostics_param(x)
    assert want["lag_bulk"] > 1100, want["lag_bulk"]  # really past the single-tile rounds
    for k in ("rhat", "ess_bulk", "ess_tail"):
        assert close(float(d[k][0].item()), want[k], TOL), (k, float(d[k][0].item()), want[k])
    assert int(d["lag_bulk"][0].item()) == want["lag_bulk"] and int(d["lag_tail"][0].item()) == want["lag_tail"]
    assert float(d["ess_bulk"][1].item()) > 0.5 * c * n  # the well-mixed parameter of the same batch is untouched


def test_heavy_ties_stay_on_the_fast_path(B):
    """Zero-inflated and discrete columns (normal in real posteriors): a value repeated far more often than any
    on-chip bucket holds gets a coarse bucket of its own (duplicate splitters -> [v, nextafter(v))) and is
    streamed, not sorted.  Exact ranks / scores / diagnostics, and no single-CTA LSD fallback: the pass over
    such columns costs about what it costs over continuous ones."""
    import time

    import torch

    rng = np.random.default_rng(41)
    c, n = 16, 10_000
    cont = rng.standard_normal((c, n))
    half_zero = np.where(rng.random((c, n)) < 0.5, 0.0, cont)
    spike = np.where(rng.random((c, n)) < 0.3, 1.25, cont) + np.where(rng.random((c, n)) < 0.1, -0.0, 0.0)
    for name, x in (("half_zero", half_zero), ("spike_030", spike)):
        m = x.size
        r = B.ranks_batch(x[None])
        rank, z = O.rank_normalize(x)
        rank_f, z_f = O.rank_normalize(O.fold_chains(x))
        assert np.array_equal(r["avg_rank"][0].cpu().numpy(), rank), name
        assert np.array_equal(r["avg_rank_folded"][0].cpu().numpy(), rank_f), name
        assert np.array_equal(r["z"][0].cpu().numpy(), z) and np.array_equal(r["z_folded"][0].cpu().numpy(), z_f), name
        _check_vs_oracle(B, x, name)
    # 21 equally likely values at the config-5 size: each holds more draws than the largest on-chip bucket
    disc = rng.integers(0, 21, (64, 4096)).astype(np.float64) * 0.5 - 3.0
    _check_vs_oracle(B, disc, "discrete_21")
    # cost: 64 zero-inflated columns against 64 continuous ones (same shape), best of 3
    xs = {"cont": rng.standard_normal((64, c, n)), "half_zero": None}
    xs["half_zero"] = np.where(rng.random((64, c, n)) < 0.5, 0.0, xs["cont"])
    ms = {}
    for name, x in xs.items():
        xt = B.to_device(x)
        B.diagnostics_batch(xt)
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            B.diagnostics_batch(xt)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        ms[name] = best * 1e3
    print(f"heavy ties: continuous {ms['cont']:.2f} ms, 50 % zeros {ms['half_zero']:.2f} ms")
    assert ms["half_zero"] < 3.0 * ms["cont"], ms


def test_slab_overflow_takes_the_lsd_path(B):
    """Single-pass partition: a repeated value that the strided sample never sees (it sits on every position the sample
    does not read) lands in an ordinary coarse bucket, whose slab it overflows; the parameter must be listed for the LSD
    path and come out exact.  A neighbouring well-behaved parameter in the same batch must be untouched."""
    rng = np.random.default_rng(91)
    c, n = 8, 8192
    m = c * n
    nb1 = 1
    while nb1 < 256 and m > nb1 * 768:
        nb1 *= 2
    s = min(nb1 * 16, m)
    sampled = np.array([((2 * j + 1) * m) // (2 * s) for j in range(s)])
    evil = np.full(m, 0.125)
    evil[sampled] = rng.standard_normal(s)
    good = rng.standard_normal(m)
    x = np.stack([good.reshape(c, n), evil.reshape(c, n), good.reshape(c, n) * 2.0 + 1.0])
    d = B.diagnostics_batch(x)
    for i in range(3):
        want = O.diagnostics_param(x[i])
        for k in ("rhat", "ess_bulk", "ess_tail"):
            assert close(float(d[k][i].item()), want[k], TOL), (i, k, float(d[k][i].item()), want[k])
        ns = O.numpy_stats(x[i].reshape(-1))
        for j, q in enumerate(("q5", "q50", "q95")):
            assert float(d["quantiles"][i, j].item()) == ns[q], (i, q)
    r = B.ranks_batch(x[1][None])
    rank, z = O.rank_normalize(x[1])
    assert np.array_equal(r["avg_rank"][0].cpu().numpy(), rank)
    assert np.array_equal(r["z"][0].cpu().numpy(), z)


def test_many_discrete_values_in_slabs(B):
    """Columns rounded to a grid (every 101st parameter of the synthetic workload): dozens of pure buckets whose slabs
    are sized from the sample, next to ordinary buckets that hold a few grid points each."""
    rng = np.random.default_rng(92)
    c, n = 16, 10_000
    for scale, name in ((0.25, "grid_150_values"), (4.0, "grid_2400_values")):
        x = np.round(rng.standard_normal((c, n)) * scale, 2)
        r = B.ranks_batch(x[None])
        rank, z = O.rank_normalize(x)
        rank_f, z_f = O.rank_normalize(O.fold_chains(x))
        assert np.array_equal(r["avg_rank"][0].cpu().numpy(), rank), name
        assert np.array_equal(r["avg_rank_folded"][0].cpu().numpy(), rank_f), name
        assert np.array_equal(r["z"][0].cpu().numpy(), z) and np.array_equal(r["z_folded"][0].cpu().numpy(), z_f), name
        _check_vs_oracle(B, x, name)


@pytest.mark.parametrize("n", [8, 16, 24, 40, 1000, 1016, 2048, 2056, 4096, 6152, 16392, 40_000])
def test_round0_tile_variant_chain_lengths(B, n):
    """Chain lengths that are multiples of 8 take the tile variant of round 0 (k_moments0t: 2-D TMA into swizzled
    landing buffers, both kinds per thread): tile and unit boundaries, a split point inside a row (n = 1000, 1016,
    6152), chains shorter than the lag window, sticky chains that go on into the later rounds -- against the oracle,
    truncation lags compared for equality."""
    from scipy.signal import lfilter

    rng = np.random.default_rng(1000 + n)
    c = 5
    for phi in (0.0, 0.9):
        x = lfilter([math.sqrt(1 - phi * phi)], [1.0, -phi], rng.standard_normal((c, n)), axis=1)
        x[2] += 0.3  # unequal chain means: between-chain variance in R-hat and in var_hat
        _check_vs_oracle(B, x, f"n={n} phi={phi}")
    if n >= 16:
        x = rng.standard_normal((c, n))
        x[:, : n // 2] = 1.5  # constant first halves: zero within-half variance must come out exactly
        x[1] = -2.0           # and one constant chain
        _check_vs_oracle(B, x, f"n={n} constant halves")



def test_sticky_prediction_cannot_change_results(B):
    """Round 0 covers 48 lags instead of 16 for the parameters whose FIRST chain looks sticky (k_predict_sticky); an
    unfinished item then enters the later rounds at lag 49 (list 1) instead of lag 17 (list 0).  The guess may be
    wrong in both directions and the two lists may be empty or not in any combination: R-hat, ESS and the truncation
    lags must match the oracle in every case, alone and batched together, also with the Geyer truncation."""
    from scipy.signal import lfilter

    rng = np.random.default_rng(4949)
    c, n = 6, 4096

    def ar(phi, rows):
        return lfilter([math.sqrt(1 - phi * phi)], [1.0, -phi], rng.standard_normal((rows, n)), axis=1)

    all_sticky = ar(0.93, c)                                   # guessed sticky, needs more than 48 lags: list 1
    first_only = np.concatenate([ar(0.95, 1), ar(0.0, c - 1)])  # guessed sticky, the pooled walk ends within a few lags
    all_but_first = np.concatenate([ar(0.0, 1), ar(0.9, c - 1)])  # not guessed: list 0, rounds from lag 17
    mild = ar(0.5, c)                                          # ends inside round 0 either way
    very = ar(0.99, c)                                         # several later rounds
    cases = {"all_sticky": all_sticky, "first_only": first_only, "all_but_first": all_but_first, "mild": mild, "very": very}
    want = {k: O.diagnostics_param(v) for k, v in cases.items()}

    def geyer_want(x):
        _, z = O.rank_normalize(x)
        _, zf = O.rank_normalize(O.fold_chains(x))
        (wb, lb), (wt, lt) = O.ess_geyer(z, return_lag=True), O.ess_geyer(zf, return_lag=True)
        return {"ess_bulk": wb, "ess_tail": wt, "lag_bulk": lb, "lag_tail": lt}

    def check(names, geyer=False):
        x = np.stack([cases[k] for k in names])
        d = B.diagnostics_batch(x, ess_truncation="geyer") if geyer else B.diagnostics_batch(x)
        for i, k in enumerate(names):
            w = geyer_want(cases[k]) if geyer else want[k]
            for key in ("ess_bulk", "ess_tail") if geyer else ("rhat", "ess_bulk", "ess_tail"):
                assert close(float(d[key][i].item()), w[key], TOL), (names, k, key, float(d[key][i].item()), w[key])
            assert int(d["lag_bulk"][i].item()) == w["lag_bulk"] and int(d["lag_tail"][i].item()) == w["lag_tail"], (names, k)

    for k in cases:
        check([k])                                   # one list empty, the other not (or both)
    check(list(cases))                               # both lists in one launch
    check(["all_but_first", "mild"])                 # only list 0
    check(["all_sticky", "very", "first_only"])      # only list 1
    check(list(cases), geyer=True)



def test_randomised_shapes_against_the_oracle(B):
    """`tools/fuzz_parity.py`: 1 .. 9 chains, chain lengths around the tile / sub-tile / split boundaries, ties,
    constant stretches, sticky / anti-correlated / heavy-tailed chains, mixed batches -- R-hat, ESS, truncation lags and
    quantiles against the oracle."""
    import importlib.util
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("fuzz_parity", Path(__file__).resolve().parent.parent / "tools" / "fuzz_parity.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(60, 7) == 0


def test_ragged_chains_follow_the_reference(B):
    """Chains of different lengths (fixture from the unmodified reference, `make_ragged_golden.py`): pooled ranks over
    all draws, each chain split at its own midpoint, every (split) chain truncated to the shortest one."""
    from conftest import dec
    from mcmc_ref_b200 import diagnostics as D

    cases = json.loads((GOLDEN / "ragged_cases.json").read_text())
    assert len(cases) >= 9
    for case in cases:
        chains = case["chains"]
        got = D.diagnostics(chains)
        for k in ("rhat", "ess_bulk", "ess_tail"):
            assert close(got[k], dec(case[k]), TOL), (case["name"], k, got[k], case[k])
        assert close(D.split_rhat(chains), dec(case["rhat"]), TOL), case["name"]
        assert close(D.ess_bulk(chains), dec(case["ess_bulk"]), TOL) and close(D.ess_tail(chains), dec(case["ess_tail"]), TOL)
    # the chain-statistics entry point on its own: scores in, R-hat / ESS out == the fused pass on equal chains
    rng = np.random.default_rng(77)
    x = rng.standard_normal((3, 4, 500)) + rng.standard_normal((3, 4, 1)) * 0.1
    r = B.ranks_batch(x)
    st = B.chain_stats_batch(r["z"], r["z_folded"])
    d = B.diagnostics_batch(x)
    for k in ("rhat", "rhat_bulk", "rhat_tail", "ess_bulk", "ess_tail", "lag_bulk", "lag_tail"):
        assert np.array_equal(st[k].cpu().numpy(), d[k].cpu().numpy(), equal_nan=True), k
