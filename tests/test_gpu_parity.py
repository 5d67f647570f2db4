"""GPU parity tests: the CUDA path (through the Python mirror -> ctypes -> C-ABI) against
the CPU oracle and the committed golden fixtures.

Exact class (bit-exact): average ranks, selected order statistics, numpy-linear
quantiles, medians, verdict booleans, failure strings, and z in ALL AS241 branches (the tail
branches go through log(): the device restates glibc's log, so Z_TAIL_ULP = 0).
Tolerance class: R-hat / ESS / mean / std -- BASELINE.json states 1e-6 relative; the
tests hold the kernels to TOL = 1e-9.
"""

from __future__ import annotations

import math

import numpy as np
import pytest

from conftest import close, dec, same_float
from oracle import mcmc_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-9
Z_TAIL_ULP = 0


@pytest.fixture(scope="module")
def B():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import __graft_entry__ as g

    g.build()
    from mcmc_ref_b200 import batch

    return batch


def ulp_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ai = a.view(np.int64).copy()
    bi = b.view(np.int64).copy()
    ai = np.where(ai < 0, np.int64(-(2**63)) - ai, ai)
    bi = np.where(bi < 0, np.int64(-(2**63)) - bi, bi)
    return np.abs(ai - bi)


def check_z(got: np.ndarray, want: np.ndarray, ranks: np.ndarray, m: int, label=""):
    """Central branch bit-exact; tails within Z_TAIL_ULP.  Returns fraction of bit-equal tail entries."""
    p = (ranks - 0.5) / m
    central = np.abs(p - 0.5) <= 0.425
    assert np.array_equal(got[central], want[central]), f"{label}: central-branch z not bit-exact"
    tail = ~central
    if tail.any():
        d = ulp_diff(got[tail], want[tail])
        assert d.max() <= Z_TAIL_ULP, f"{label}: tail z off by {d.max()} ulp"
        return float((d == 0).mean())
    return 1.0


def assert_close_arr(got, want, rel, label):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    for i, (g, w) in enumerate(zip(got.ravel(), want.ravel())):
        assert close(float(g), float(w), rel), (label, i, g, w)


# ---------------------------------------------------------------- golden: reference cases

def test_ref_cases_ranks_and_z(B, ref_cases):
    fracs = []
    for case in ref_cases:
        if "z" not in case:
            continue
        x = np.array(case["chains"], dtype=np.float64)
        out = B.ranks_batch(x[None])
        m = x.size
        rank = out["avg_rank"][0].cpu().numpy()
        rank_f = out["avg_rank_folded"][0].cpu().numpy()
        assert np.array_equal(rank, np.array(case["avg_rank"])), case["name"]
        assert np.array_equal(rank_f, np.array(case["avg_rank_folded"])), case["name"]
        fracs.append(check_z(out["z"][0].cpu().numpy(), np.array(case["z"]), rank, m, case["name"]))
        fracs.append(check_z(out["z_folded"][0].cpu().numpy(), np.array(case["z_folded"]), rank_f, m, case["name"]))
    assert len(fracs) >= 40


def test_ref_cases_scalar_api(B, ref_cases):
    from mcmc_ref_b200 import diagnostics as D

    for case in ref_cases:
        ch, kw = case["chains"], {"min_chains": case["min_chains"]}
        if "error" in case:
            with pytest.raises(ValueError) as ei:
                D.split_rhat(ch, **kw)
            assert str(ei.value) == case["error"]
            continue
        got = (D.split_rhat(ch, **kw), D.ess_bulk(ch, **kw), D.ess_tail(ch, **kw))
        want = (dec(case["rhat"]), dec(case["ess_bulk"]), dec(case["ess_tail"]))
        for g, w, nm in zip(got, want, ("rhat", "ess_bulk", "ess_tail")):
            assert close(g, w, TOL), (case["name"], nm, g, w)
        all3 = D.diagnostics(ch, **kw)
        assert same_float(all3["rhat"], got[0]) and same_float(all3["ess_bulk"], got[1])


def test_reference_unit_tests_through_new_names(B):
    """The reference's own tests/unit/test_diagnostics.py:10-58 and test_compare.py:6-24."""
    from mcmc_ref_b200.compare import compare_stats
    from mcmc_ref_b200.diagnostics import ess_bulk, split_rhat

    assert 0.99 <= split_rhat([[1.0] * 4] * 4) <= 1.01
    assert ess_bulk([[1.0, 2.0, 3.0, 4.0], [1.1, 2.1, 3.1, 4.1], [0.9, 1.9, 2.9, 3.9], [1.05, 2.05, 3.05, 4.05]]) > 0
    assert split_rhat([[0.0] * 4, [10.0] * 4, [0.0] * 4, [10.0] * 4]) > 1.1
    with pytest.raises(ValueError, match="at least 4 chains"):
        split_rhat([[1.0, 2.0, 3.0, 4.0], [1.1, 2.1, 3.1, 4.1]])
    assert math.isnan(split_rhat([[1.0, 2.0, 3.0, 4.0]], min_chains=1))
    with pytest.raises(ValueError, match="at least 4 chains"):
        ess_bulk([[1.0, 2.0, 3.0, 4.0], [1.1, 2.1, 3.1, 4.1]])
    ref = {"mu": {"mean": 1.0, "std": 1.0}}
    res = compare_stats(ref, {"mu": {"mean": 1.05, "std": 0.95}}, tolerance=0.1, metrics=["mean", "std"])
    assert res.passed is True and res.failures == [] and res.details["mu"]["mean"].passed is True
    res = compare_stats(ref, {"mu": {"mean": 2.0, "std": 1.0}}, tolerance=0.1, metrics=["mean", "std"])
    assert res.passed is False and res.failures


# ---------------------------------------------------------------- golden: packaged models

def test_packaged_subset_diagnostics_vs_shipped_meta(B, packaged_subset):
    arrays, meta = packaged_subset
    n = 0
    for model, info in meta.items():
        out = B.diagnostics_batch(arrays[model])
        for k in ("rhat", "ess_bulk", "ess_tail"):
            got = out[k].cpu().numpy()
            for i, p in enumerate(info["params"]):
                assert close(float(got[i]), info["diagnostics"][p][k], TOL), (model, p, k)
                n += 1
        from mcmc_ref_b200.convert import _checks

        diag = {p: {k: float(out[k][i].item()) for k in ("rhat", "ess_bulk", "ess_tail")}
                for i, p in enumerate(info["params"])}
        assert _checks(info["n_chains"], info["n_draws_per_chain"], diag) == info["checks"]
        ess_ok, rhat_ok = B.checks_batch(out["rhat"], out["ess_bulk"])
        assert ess_ok == info["checks"]["ess_above_400"] and rhat_ok == info["checks"]["rhat_below_1_01"]
    assert n == 96


def test_packaged_subset_stats_backend(B, packaged_subset):
    """B200Backend.stats vs NumpyBackend.stats golden (quantiles bit-exact)."""
    from mcmc_ref_b200.backends import get_backend

    arrays, meta = packaged_subset
    be = get_backend("b200")
    for model, info in meta.items():
        x = arrays[model]
        table = {p: x[i].reshape(-1) for i, p in enumerate(info["params"])}
        got = be.stats(table, info["params"])
        for p in info["params"]:
            want = info["numpy_stats"][p]
            assert set(got[p]) == set(want) == {"mean", "std", "q5", "q50", "q95"}
            for k in ("q5", "q50", "q95"):
                assert got[p][k] == want[k], (model, p, k)
            assert close(got[p]["mean"], want["mean"], 1e-13), (model, p)
            assert close(got[p]["std"], want["std"], 1e-12), (model, p)
    other = be.stats({"a": np.arange(101.0)}, ["a"], quantiles=(0.0, 0.25, 1.0))
    assert other["a"] == {"mean": 50.0, "std": float(np.std(np.arange(101.0))), "q0": 0.0, "q25": 25.0, "q100": 100.0}


def test_compare_cases(B, compare_cases):
    from mcmc_ref_b200.compare import compare_stats, compute_basic_stats, compute_stats_from_draws

    for c in compare_cases["basic"]:
        st = compute_basic_stats(c["values"])
        assert close(st["mean"], dec(c["mean"]), 1e-13)
        assert close(st["std"], dec(c["std"]), 1e-9 if len(c["values"]) == 64 else 1e-12)
    for c in compare_cases["compare"]:
        ref = {k: {m: dec(v) for m, v in d.items()} for k, d in c["ref"].items()}
        res = compare_stats(ref, c["act"], c["tolerance"], c["metrics"])
        assert res.passed == c["passed"]
        assert res.failures == c["failures"]
        for p, d in c["details"].items():
            for m, (r, a, rel, ok) in d.items():
                g = res.details[p][m]
                assert same_float(g.ref, dec(r)) and same_float(g.actual, dec(a))
                assert same_float(g.rel_error, dec(rel)) and g.passed == ok
    st = compute_stats_from_draws({"a": [1.0, 2.0, 1.5, 2.5], "b": [], "c": [3.0, 3.0]})
    assert st["a"] == {"mean": 1.75, "std": 0.5590169943749475}
    assert math.isnan(st["b"]["mean"]) and st["c"] == {"mean": 3.0, "std": 0.0}


# ---------------------------------------------------------------- seeded inputs vs oracle

def _ar1(rng, c, n, phi):
    e = rng.standard_normal((c, n))
    x = np.empty((c, n))
    x[:, 0] = e[:, 0]
    s = math.sqrt(1 - phi * phi)
    for i in range(1, n):
        x[:, i] = phi * x[:, i - 1] + s * e[:, i]
    return x


def _workload(rng, p, c, n):
    x = np.empty((p, c, n))
    for i in range(p):
        phi = (0.0, 0.5, 0.9)[i % 3]
        x[i] = (i % 17 - 8) + 2.0 ** (i % 5 - 2) * _ar1(rng, c, n, phi)
        if i % 4 == 3:
            x[i] = np.round(x[i], 2)
        if i % 5 == 4:
            x[i, 1::2] += 0.4 * 2.0 ** (i % 5 - 2)
    return x


# the last two shapes exceed ~330k pooled draws per parameter: the wide (sub-bucketed) ordering path
@pytest.mark.parametrize("shape", [(6, 16, 10000), (4, 64, 4096), (5, 4, 1001), (3, 7, 2049), (9, 10, 1000),
                                   (4, 4, 100001), (2, 8, 262145)])
def test_seeded_against_oracle(B, shape):
    p, c, n = shape
    rng = np.random.default_rng(p * 1000 + c)
    x = _workload(rng, p, c, n)
    m = c * n
    r = B.ranks_batch(x)
    d = B.diagnostics_batch(x)
    fr = []
    for i in range(p):
        rank, z = O.rank_normalize(x[i])
        folded = O.fold_chains(x[i])
        rank_f, z_f = O.rank_normalize(folded)
        assert np.array_equal(r["avg_rank"][i].cpu().numpy(), rank), i
        assert np.array_equal(r["avg_rank_folded"][i].cpu().numpy(), rank_f), i
        fr.append(check_z(r["z"][i].cpu().numpy(), z, rank, m, f"param {i}"))
        fr.append(check_z(r["z_folded"][i].cpu().numpy(), z_f, rank_f, m, f"param {i} folded"))
        want = O.diagnostics_param(x[i])
        for k in ("rhat", "rhat_bulk", "rhat_tail", "ess_bulk", "ess_tail"):
            assert close(float(d[k][i].item()), want[k], TOL), (i, k, float(d[k][i].item()), want[k])
        assert int(d["lag_bulk"][i].item()) == want["lag_bulk"], i
        assert int(d["lag_tail"][i].item()) == want["lag_tail"], i
        st = O.numpy_stats(x[i].reshape(-1))
        bs = O.compute_basic_stats(x[i].reshape(-1))
        assert close(float(d["mean"][i].item()), bs["mean"], 1e-13)
        assert close(float(d["std"][i].item()), bs["std"], 1e-11)
        for j, k in enumerate(("q5", "q50", "q95")):
            assert float(d["quantiles"][i, j].item()) == st[k], (i, k)
        assert float(d["median"][i].item()) == O.pooled_median(x[i].reshape(-1))
        s = np.sort(x[i].reshape(-1))
        for j, q in enumerate((0.05, 0.5, 0.95)):
            _, lo, a, b = O.quantile_lerp(s, q)
            assert float(d["order_stats"][i, j, 0].item()) == a and float(d["order_stats"][i, j, 1].item()) == b
    print(f"shape {shape}: mean fraction of tail z bit-equal to the stdlib = {np.mean(fr):.4f}")


def test_geyer_truncation_extension(B):
    """EXTENSION: Geyer initial-monotone ESS (oracle-defined); the default stays the reference's rule."""
    rng = np.random.default_rng(7)
    c, n = 4, 6001
    x = np.stack([_ar1(rng, c, n, 0.0), 3.0 + _ar1(rng, c, n, 0.5), _ar1(rng, c, n, 0.95), -_ar1(rng, c, n, -0.6),
                  np.round(_ar1(rng, c, n, 0.9), 1)])
    d = B.diagnostics_batch(x, ess_truncation="geyer")
    ref = B.diagnostics_batch(x)
    for i in range(x.shape[0]):
        _, z = O.rank_normalize(x[i])
        _, zf = O.rank_normalize(O.fold_chains(x[i]))
        wb, lb = O.ess_geyer(z, return_lag=True)
        wt, lt = O.ess_geyer(zf, return_lag=True)
        assert close(float(d["ess_bulk"][i].item()), wb, TOL), (i, float(d["ess_bulk"][i].item()), wb)
        assert close(float(d["ess_tail"][i].item()), wt, TOL), (i, float(d["ess_tail"][i].item()), wt)
        assert int(d["lag_bulk"][i].item()) == lb and int(d["lag_tail"][i].item()) == lt
        assert close(float(d["rhat"][i].item()), float(ref["rhat"][i].item()), 0.0)  # R-hat is untouched
        want = O.diagnostics_param(x[i])
        assert close(float(ref["ess_bulk"][i].item()), want["ess_bulk"], TOL)       # default = reference rule
    assert int(d["lag_bulk"][2].item()) > 48  # the sticky chain's walk crossed several rounds
    with pytest.raises(ValueError, match="ess_truncation"):
        B.diagnostics_batch(x, ess_truncation="fft")


def test_long_sticky_chains_many_rounds(B):
    """Chains far longer than any on-chip buffer, and autocorrelation that needs hundreds of lags."""
    rng = np.random.default_rng(99)
    c, n = 4, 50001
    x = np.stack([_ar1(rng, c, n, 0.99), 5.0 + 0.1 * _ar1(rng, c, n, 0.9), _ar1(rng, c, n, 0.0)])
    d = B.diagnostics_batch(x)
    for i in range(3):
        want = O.diagnostics_param(x[i])
        for k in ("rhat", "ess_bulk", "ess_tail"):
            assert close(float(d[k][i].item()), want[k], TOL), (i, k, float(d[k][i].item()), want[k])
        assert int(d["lag_bulk"][i].item()) == want["lag_bulk"] and int(d["lag_tail"][i].item()) == want["lag_tail"]
    assert int(d["lag_bulk"][0].item()) > 100  # really went through many 16-lag rounds


def test_edge_cases(B):
    rng = np.random.default_rng(5)
    cases = {
        "constant": np.full((4, 300), 2.5),
        "two_values_long_runs": np.repeat(np.array([[0.0, 1.0]]), 3000, axis=0).T.reshape(4, 1500).copy(),
        "few_distinct": rng.integers(0, 3, (8, 1000)).astype(float),
        "signed_zero": np.where(rng.random((4, 700)) < 0.5, 0.0, -0.0) * 1.0,
        "n1": rng.standard_normal((4, 1)),
        "n2": rng.standard_normal((4, 2)),
        "n3": rng.standard_normal((4, 3)),
        "n5_c2": rng.standard_normal((2, 5)),
        "tile_edge_2048": np.round(rng.standard_normal((2, 1024)), 1),
        "tile_edge_2049": np.round(rng.standard_normal((1, 2049)), 1),
        "tile_edge_4097": rng.standard_normal((1, 4097)),
        "huge_magnitudes": 10.0 ** rng.uniform(-300, 300, (4, 500)) * rng.choice([-1, 1], (4, 500)),
        "denormals": rng.standard_normal((4, 200)) * 1e-310,
        "all_negative_tight": -1e6 + 1e-6 * rng.standard_normal((4, 900)),
    }
    for name, x in cases.items():
        m = x.size
        r = B.ranks_batch(x[None])
        rank, z = O.rank_normalize(x)
        rank_f, z_f = O.rank_normalize(O.fold_chains(x))
        assert np.array_equal(r["avg_rank"][0].cpu().numpy(), rank), name
        assert np.array_equal(r["avg_rank_folded"][0].cpu().numpy(), rank_f), name
        check_z(r["z"][0].cpu().numpy(), z, rank, m, name)
        check_z(r["z_folded"][0].cpu().numpy(), z_f, rank_f, m, name)
        d = B.diagnostics_batch(x[None])
        want = O.diagnostics_param(x)
        for k in ("rhat", "ess_bulk", "ess_tail"):
            assert close(float(d[k][0].item()), want[k], TOL), (name, k, float(d[k][0].item()), want[k])
        assert float(d["quantiles"][0, 1].item()) == O.numpy_stats(x.reshape(-1))["q50"], name


def test_nonfinite_rejected(B):
    x = np.random.default_rng(0).standard_normal((2, 4, 100))
    for bad in (np.nan, np.inf, -np.inf):
        y = x.copy()
        y[1, 2, 50] = bad
        with pytest.raises(ValueError, match="non-finite"):
            B.diagnostics_batch(y)
        with pytest.raises(ValueError, match="non-finite"):
            B.stats_batch(y.reshape(2, -1))
    B.diagnostics_batch(x)  # the error state does not stick


def test_batching_is_invisible(B):
    """A workspace that holds only a few parameters gives the same bits as one big batch."""
    from mcmc_ref_b200 import _native

    rng = np.random.default_rng(11)
    x = _workload(rng, 13, 4, 600)
    lib = _native.load()
    full = B.diagnostics_batch(x)
    cap = int(lib.b200ref_workspace_bytes(3, 4, 600))
    small = B.diagnostics_batch(x, workspace=B.Workspace(), ws_cap_bytes=cap)
    for k, v in full.items():
        assert np.array_equal(v.cpu().numpy(), small[k].cpu().numpy(), equal_nan=True), k


# ---------------------------------------------------------------- fused pass

def test_full_pass_device_and_host(B):
    import torch

    p, c, n, mref = 40, 8, 512, 1000
    cand = B.synth_fill(p, c, n, kind=0)
    ref = B.synth_fill(p, 10, 100, kind=1).reshape(p, mref)
    summary, passed = B.full_pass(cand, ref, tolerance=0.15)
    got = summary.cpu().numpy()
    want = O.full_pass(cand.cpu().numpy(), ref.cpu().numpy(), tolerance=0.15)
    cols = B.SUMMARY_COLUMNS
    for name in ("rhat", "ess_bulk", "ess_tail", "mean", "std", "ref_mean", "ref_std", "rel_mean", "rel_std"):
        assert_close_arr(got[:, cols.index(name)], want[name], TOL, name)
    for name in ("q5", "q50", "q95", "ref_q5", "ref_q50", "ref_q95"):
        assert np.array_equal(got[:, cols.index(name)], want[name]), name
    flags = got[:, cols.index("flags")].astype(int)
    assert np.array_equal((flags & 1) != 0, want["pass_mean"])
    assert np.array_equal((flags & 2) != 0, want["pass_std"])
    assert passed == want["passed"]
    # extension columns (MCSE of the mean, z-score check): oracle-defined, tolerance-class
    for name in ("mcse_mean", "ref_se_mean", "z_mean"):
        assert_close_arr(got[:, cols.index(name)], want[name], TOL, name)
    assert np.array_equal(got[:, cols.index("z_pass")] != 0, want["z_pass"])
    s3, _ = B.full_pass(cand, ref, tolerance=0.15, z_crit=1.0)
    want3 = O.full_pass(cand.cpu().numpy(), ref.cpu().numpy(), tolerance=0.15, z_crit=1.0)
    assert np.array_equal(s3.cpu().numpy()[:, cols.index("z_pass")] != 0, want3["z_pass"])
    assert want3["z_pass"].any() and not want3["z_pass"].all()  # both verdicts occur at |z| <= 1
    ch = cand.cpu().pin_memory()
    rh = ref.cpu().pin_memory()
    s_host, passed_h = B.full_pass_host(ch, rh, tolerance=0.15)
    assert passed_h == passed
    assert np.array_equal(s_host.numpy(), got, equal_nan=True)
    # slabs smaller than P: same bits
    lib = B._native.load()
    slab = 16
    small_ws = int(lib.b200ref_workspace_bytes(slab, c, n)) + 2 * slab * (c * n + mref + len(cols)) * 8 + slab * 13 * 8 + 8192
    s_host2, _ = B.full_pass_host(ch, rh, tolerance=0.15, workspace=B.Workspace(), ws_bytes=small_ws)
    assert np.array_equal(s_host2.numpy(), got, equal_nan=True)
    torch.cuda.synchronize()


def test_synth_is_tiling_independent(B):
    a = B.synth_fill(10, 4, 300, p0=0)
    b = B.synth_fill(4, 4, 300, p0=6)
    assert np.array_equal(a[6:].cpu().numpy(), b.cpu().numpy())
    x = a.cpu().numpy()
    assert abs(x[1].mean() - (1 % 17 - 8)) < 0.5  # mu_p = (p mod 17) - 8
    r = B.synth_fill(10, 1, 5000, kind=1).cpu().numpy()
    assert abs(r[3].std() - 2.0 ** (3 % 5 - 2)) < 0.1 * 2.0 ** (3 % 5 - 2)


# ---------------------------------------------------------------- full-size properties

def test_full_size_properties_c2(B):
    """BASELINE config 2 shape (a 128-parameter slice): size-independent properties."""
    import torch

    p, c, n = 128, 16, 10000
    m = c * n
    x = B.synth_fill(p, c, n, kind=0)
    r = B.ranks_batch(x)
    rank = r["avg_rank"].reshape(p, m)
    # sum of average ranks is M(M+1)/2 exactly, ties or not
    assert torch.equal(rank.sum(dim=1), torch.full((p,), m * (m + 1) / 2, dtype=torch.float64, device=rank.device))
    # sortedness: ordering the draws by rank gives a non-decreasing sequence
    xs = x.reshape(p, m)
    order = torch.argsort(rank, dim=1, stable=True)
    sorted_x = torch.gather(xs, 1, order)
    assert bool((sorted_x[:, 1:] >= sorted_x[:, :-1]).all())
    # idempotence: ranks of the ranks are the ranks
    r2 = B.ranks_batch(r["avg_rank"], folded=False)
    assert torch.equal(r2["avg_rank"], r["avg_rank"])
    # z is an odd function of the centred rank up to the rounding of p: z(rank) ~ -z(M + 1 - rank)
    z = r["z"].reshape(p, m)
    zs = torch.gather(z, 1, order)
    no_ties = (rank == rank.round()).all(dim=1)
    assert int(no_ties.sum()) > p // 2
    sel = zs[no_ties]
    assert torch.allclose(sel, -sel.flip(1), rtol=0.0, atol=1e-10)
    assert bool((sel[:, 1:] > sel[:, :-1]).all())  # strictly increasing in rank
    # folded ranks: same sum; median element has the smallest folded rank
    rf = r["avg_rank_folded"].reshape(p, m)
    assert torch.equal(rf.sum(dim=1), rank.sum(dim=1))
    d = B.diagnostics_batch(x)
    rhat = d["rhat"].cpu().numpy()
    # every 97th parameter has shifted chains -> fails rhat < 1.01; the others pass
    bad = np.array([(i % 97) == 96 for i in range(p)])
    assert (rhat[bad] > 1.01).all() and (rhat[~bad] < 1.01).all()
    ess = d["ess_bulk"].cpu().numpy()
    assert (ess > 0).all() and (ess <= m).all()
    # oracle spot-check on 3 parameters of the full-size run
    for i in (0, 1, 2, 100):
        want = O.diagnostics_param(x[i].cpu().numpy())
        for k in ("rhat", "ess_bulk", "ess_tail"):
            assert close(float(d[k][i].item()), want[k], TOL), (i, k)


@pytest.mark.parametrize("shape", [(96, 64, 4096), (2, 8, 1_000_000)], ids=["c5-slice", "c4-slice"])
def test_full_size_properties_c5_c4(B, shape):
    """BASELINE configs 5 (wide posterior) and 4 (long chains) at their real C x N: size-independent properties.

    * sum of the average ranks = M(M+1)/2 exactly;
    * a strictly increasing transform of the draws (exp) leaves ranks, bulk scores and bulk ESS / R-hat bit-identical;
    * reversing the chain order leaves pooled ranks untouched and permutes nothing else;
    * quantiles are ordered and lie inside [min, max]; 0 < ESS <= M.
    """
    import torch

    p, c, n = shape
    m = c * n
    x = B.synth_fill(p, c, n, kind=0)
    x = (x - x.mean(dim=(1, 2), keepdim=True)) / x.std(dim=(1, 2), keepdim=True)  # keep exp() finite and injective
    d = B.diagnostics_batch(x)
    r = B.ranks_batch(x, folded=False)
    rank = r["avg_rank"].reshape(p, m)
    assert torch.equal(rank.sum(dim=1), torch.full((p,), m * (m + 1) / 2, dtype=torch.float64, device=rank.device))
    y = torch.exp(x)
    ry = B.ranks_batch(y, folded=False)
    untied = (rank == rank.round()).all(dim=1)  # rounding exp() may merge neighbours of a tied parameter
    assert int(untied.sum()) >= max(1, p // 2)
    assert torch.equal(ry["avg_rank"][untied], r["avg_rank"][untied])
    assert torch.equal(ry["z"][untied], r["z"][untied])
    dy = B.diagnostics_batch(y)
    assert torch.equal(dy["rhat_bulk"][untied], d["rhat_bulk"][untied])
    assert torch.equal(dy["ess_bulk"][untied], d["ess_bulk"][untied])
    rr = B.ranks_batch(x.flip(1), folded=False)
    assert torch.equal(rr["avg_rank"], r["avg_rank"].flip(1))
    q = d["quantiles"]
    xs = x.reshape(p, m)
    assert bool((q[:, 0] <= q[:, 1]).all() and (q[:, 1] <= q[:, 2]).all())
    assert bool((q[:, 0] >= xs.min(dim=1).values).all() and (q[:, 2] <= xs.max(dim=1).values).all())
    assert torch.equal(q[:, 1], d["median"]) or m % 2 == 0  # odd M: q50 is the middle order statistic
    for k in ("ess_bulk", "ess_tail"):
        assert bool((d[k] > 0).all() and (d[k] <= m).all())
    if p > 2:  # oracle spot-check at the full C x N of config 5
        for i in (0, 5):
            want = O.diagnostics_param(x[i].cpu().numpy())
            for k in ("rhat", "ess_bulk", "ess_tail"):
                assert close(float(d[k][i].item()), want[k], TOL), (i, k)


def test_two_devices_from_one_process(B):
    """Helper streams and kernel attributes are kept per device: one process may drive several GPUs."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(21)
    x = _workload(rng, 5, 4, 3000)
    ref = rng.standard_normal((5, 1000))
    outs = []
    for dev in (0, 1, 0):
        with torch.cuda.device(dev):
            xd = torch.from_numpy(x).to(f"cuda:{dev}")
            rd = torch.from_numpy(ref).to(f"cuda:{dev}")
            summary, _ = B.full_pass(xd, rd, workspace=B.Workspace())
            outs.append(summary.cpu().numpy())
    assert np.array_equal(outs[0], outs[1], equal_nan=True) and np.array_equal(outs[0], outs[2], equal_nan=True)


# ---------------------------------------------------------------- ingestion, convert gate, CLI

def _table_from(x: np.ndarray, names):
    import pyarrow as pa

    p, c, n = x.shape
    cols = {"chain": np.repeat(np.arange(c, dtype=np.int32), n), "draw": np.tile(np.arange(n, dtype=np.int32), c)}
    for i, name in enumerate(names):
        cols[name] = x[i].reshape(-1)
    return pa.table(cols)


def test_ingest_orders_rows_on_device(B):
    from mcmc_ref_b200 import ingest

    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, 4, 257))
    names = ["a", "b[1]", "b[2]"]
    table = _table_from(x, names)
    got, c, n = ingest.table_to_device(table, names)
    assert (c, n) == (4, 257) and np.array_equal(got.cpu().numpy(), x)
    shuffled = table.take(rng.permutation(table.num_rows))
    got2, _, _ = ingest.table_to_device(shuffled, ["b[2]", "a"])
    assert np.array_equal(got2.cpu().numpy(), x[[2, 0]])
    import torch

    dev = got.device
    ch = torch.from_numpy(np.array([0, 0, 1, 1], dtype=np.int64)).to(dev)
    assert ingest.rows_in_order(ch, torch.tensor([0, 1, 0, 1], device=dev)) == (True, False)
    assert ingest.rows_in_order(ch, torch.tensor([1, 0, 0, 1], device=dev)) == (False, False)
    assert ingest.rows_in_order(ch, torch.tensor([0, 0, 0, 1], device=dev)) == (True, True)


def test_convert_gate_reproduces_shipped_meta(B, packaged_subset, tmp_path):
    """convert_table on the packaged draws gives the meta.json the reference shipped (convert.py:44-66)."""
    from mcmc_ref_b200 import convert

    arrays, meta = packaged_subset
    for model, info in meta.items():
        table = _table_from(arrays[model], info["params"])
        res = convert.convert_table(table, model, tmp_path, tmp_path, force=True)
        m = res.meta
        assert m["parameters"] == info["params"] and m["n_chains"] == info["n_chains"]
        assert m["n_draws_per_chain"] == info["n_draws_per_chain"] and m["checks"] == info["checks"]
        assert set(m) == {"model", "parameters", "n_chains", "n_draws_per_chain", "diagnostics", "generated_date",
                          "checks", "source"}
        for p in info["params"]:
            for k in ("rhat", "ess_bulk", "ess_tail"):
                assert close(m["diagnostics"][p][k], info["diagnostics"][p][k], TOL), (model, p, k)
        import json

        assert json.loads(res.meta_path.read_text())["diagnostics"].keys() == m["diagnostics"].keys()
        assert res.draws_path.exists()
    few = _table_from(np.random.default_rng(0).standard_normal((1, 2, 50)), ["mu"])
    with pytest.raises(ValueError, match="at least 4 chains"):
        convert.convert_table(few, "few")
    bad = _table_from(np.random.default_rng(0).standard_normal((1, 4, 50)), ["mu"])
    with pytest.raises(ValueError, match="quality checks failed: ndraws_is_10k"):
        convert.convert_table(bad, "bad")


def test_cli_hot_path_commands(B, tmp_path, monkeypatch):
    """The reference's CLI tests (tests/unit/test_cli.py:41-86) through the b200 commands."""
    import json

    import pyarrow as pa
    import pyarrow.parquet as pq
    from click.testing import CliRunner

    from mcmc_ref_b200.cli import main

    (tmp_path / "draws").mkdir()
    (tmp_path / "meta").mkdir()
    pq.write_table(pa.table({
        "chain": pa.array([0, 0, 1, 1, 2, 2, 3, 3], type=pa.int32()),
        "draw": pa.array([0, 1, 0, 1, 0, 1, 0, 1], type=pa.int32()),
        "mu": pa.array([1.0, 2.0, 1.5, 2.5, 1.0, 2.0, 1.5, 2.5], type=pa.float64()),
    }), tmp_path / "draws" / "example.draws.parquet")
    (tmp_path / "meta" / "example.meta.json").write_text("{}")
    monkeypatch.setenv("MCMC_REF_LOCAL_ROOT", str(tmp_path))
    runner = CliRunner()
    res = runner.invoke(main, ["stats", "example", "--format", "json", "--include-diagnostics", "--quantile-mode", "exact"])
    assert res.exit_code == 0, res.output
    out = json.loads(res.output)
    assert out["mu"]["mean"] == 1.75 and out["mu"]["std"] == 0.5590169943749475
    assert math.isnan(out["mu"]["rhat"]) and out["mu"]["ess_bulk"] == 8.0  # the reference's values for this fixture
    actual = tmp_path / "actual.csv"
    actual.write_text("chain,draw,mu\n0,0,1.0\n0,1,2.0\n1,0,1.5\n1,1,2.5\n")
    res = runner.invoke(main, ["compare", "example", "--actual", str(actual)])
    assert res.exit_code == 0 and "passed" in res.output
    actual.write_text("chain,draw,mu\n0,0,10.0\n0,1,20.0\n")
    res = runner.invoke(main, ["compare", "example", "--actual", str(actual), "--format", "json"])
    assert res.exit_code == 2
    body = json.loads(res.output)
    assert body["passed"] is False and body["failures"][0].startswith("mu.mean rel_error=")
    res = runner.invoke(main, ["diagnostics", "example", "--format", "csv", "--recompute"])
    assert res.exit_code == 0 and res.output.splitlines()[0] == "param,rhat,ess_bulk,ess_tail"
