"""CPU oracle for the mcmc-ref draw-comparison hot path.  TEST INFRASTRUCTURE ONLY.

This file is a vectorised numpy restatement of the reference's algorithm; it is
the *checker* for the CUDA path and the timed `cpu_baseline` ("port") of
``bench.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU-baseline / ``--impl reference`` legs may import it.  The product package
(``mcmc-db_b200/mcmc_ref_b200``) never imports anything from ``oracle/``.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function
here against (a) golden vectors produced by importing the unmodified reference
(``tests/golden/make_golden.py``, run in the build container where
``/root/reference`` exists), (b) the per-parameter ``{rhat, ess_bulk, ess_tail}``
values the reference ships in ``meta/*.meta.json`` for a committed subset of its
packaged models, and -- when ``/root/reference`` is present -- all 57 packaged
models (460 params, 1380 values).

Each function cites the reference lines it follows (paths relative to the
reference checkout, ``src/mcmc_ref/...``).

Arithmetic that lives outside the reference (named, with the versions used):
  * ``statistics.NormalDist().inv_cdf`` -- CPython stdlib (3.12.3 here), Wichura
    AS241 / PPND16.  The oracle CALLS the same stdlib function for every
    distinct p (so z is the reference's own value bit for bit) and also carries
    an explicit restatement (``inv_cdf_as241``) of the published algorithm in
    the stdlib's operation order, used to pin the CUDA table kernel's formula.
  * ``statistics.median`` -- restated in ``pooled_median``.
  * ``numpy.mean / numpy.std / numpy.quantile`` (numpy 2.3.5 here; the reference
    pins 2.2.6/2.4.2) -- ``numpy_stats`` calls them directly; ``quantile_lerp``
    restates numpy's ``_lerp`` so the order statistics that were selected are
    visible to the tests.
"""

from __future__ import annotations

import math
from statistics import NormalDist

import numpy as np

# --------------------------------------------------------------------------
# AS241 (stdlib statistics._normal_dist_inv_cdf), explicit restatement
# --------------------------------------------------------------------------

_A = (2.5090809287301226727e3, 3.3430575583588128105e4, 6.7265770927008700853e4,
      4.5921953931549871457e4, 1.3731693765509461125e4, 1.9715909503065514427e3,
      1.3314166789178437745e2, 3.3871328727963666080e0)
_B = (5.2264952788528545610e3, 2.8729085735721942674e4, 3.9307895800092710610e4,
      2.1213794301586595867e4, 5.3941960214247511077e3, 6.8718700749205790830e2,
      4.2313330701600911252e1, 1.0)
_C = (7.74545014278341407640e-4, 2.27238449892691845833e-2, 2.41780725177450611770e-1,
      1.27045825245236838258e0, 3.64784832476320460504e0, 5.76949722146069140550e0,
      4.63033784615654529590e0, 1.42343711074968357734e0)
_D = (1.05075007164441684324e-9, 5.47593808499534494600e-4, 1.51986665636164571966e-2,
      1.48103976427480074590e-1, 6.89767334985100004550e-1, 1.67638483018380384940e0,
      2.05319162663775882187e0, 1.0)
_E = (2.01033439929228813265e-7, 2.71155556874348757815e-5, 1.24266094738807843860e-3,
      2.65321895265761230930e-2, 2.96560571828504891230e-1, 1.78482653991729133580e0,
      5.46378491116411436990e0, 6.65790464350110377720e0)
_F = (2.04426310338993978564e-15, 1.42151175831644588870e-7, 1.84631831751005468180e-5,
      7.86869131145613259100e-4, 1.48753612908506148525e-2, 1.36929880922735805310e-1,
      5.99832206555887937690e-1, 1.0)


def _horner(coef, r):
    acc = coef[0] * r + coef[1]
    for c in coef[2:]:
        acc = acc * r + c
    return acc


def inv_cdf_as241(p: np.ndarray) -> np.ndarray:
    """Explicit AS241 in the stdlib's operation order (no fused multiply-add).

    Follows CPython ``Lib/statistics.py`` ``_normal_dist_inv_cdf`` (called from
    the reference at ``diagnostics.py:124,131``) with mu=0, sigma=1.  The tails
    go through ``math.log`` (the platform libm, as CPython does).
    """
    p = np.asarray(p, dtype=np.float64)
    q = p - 0.5
    out = np.empty_like(p)
    central = np.abs(q) <= 0.425
    qc = q[central]
    r = 0.180625 - qc * qc
    num = _horner(_A, r) * qc
    den = _horner(_B, r)
    out[central] = 0.0 + (num / den) * 1.0
    tail = ~central
    if tail.any():
        qt = q[tail]
        pt = p[tail]
        rr = np.where(qt <= 0.0, pt, 1.0 - pt)
        rr = np.sqrt(-np.array([math.log(v) for v in rr], dtype=np.float64))
        x = np.empty_like(rr)
        mid = rr <= 5.0
        r1 = rr[mid] - 1.6
        x[mid] = _horner(_C, r1) / _horner(_D, r1)
        r2 = rr[~mid] - 5.0
        x[~mid] = _horner(_E, r2) / _horner(_F, r2)
        x = np.where(qt < 0.0, -x, x)
        out[tail] = 0.0 + x * 1.0
    return out


_NORM = NormalDist()
_ZTAB_CACHE: dict[int, np.ndarray] = {}


def z_table(m: int) -> np.ndarray:
    """T[k] = inv_cdf(((k / 2) / m)) for k = 1 .. 2m-1 (T[0] unused, = nan).

    With i/j the 0-based inclusive/exclusive bounds of a tie run in the pooled
    ascending order, the reference's ``avg_rank = (i + 1 + j) / 2.0``
    (``diagnostics.py:118``) and ``p = (rank - 0.5) / n`` (``:130``) give
    ``p = ((i + j) / 2) / n`` exactly, so k = i + j indexes every value z can take.
    Entries are produced by the stdlib function itself.
    """
    tab = _ZTAB_CACHE.get(m)
    if tab is None:
        k = np.arange(2 * m, dtype=np.float64)
        p = (k * 0.5) / float(m)
        tab = np.full(2 * m, np.nan)
        inv = _NORM.inv_cdf
        tab[1:] = [inv(v) for v in p[1:].tolist()]
        if len(_ZTAB_CACHE) > 8:
            _ZTAB_CACHE.clear()
        _ZTAB_CACHE[m] = tab
    return tab


# --------------------------------------------------------------------------
# ranks / fold / split
# --------------------------------------------------------------------------

def rank_codes(flat: np.ndarray) -> np.ndarray:
    """k = i + j per element (int64), ties by ``==`` -- ``diagnostics.py:110-122``.

    ``avg_rank = (k + 1) / 2``.  ``-0.0 == 0.0`` ties as in Python.
    """
    flat = np.asarray(flat, dtype=np.float64)
    n = flat.size
    order = np.argsort(flat, kind="stable")
    s = flat[order]
    new_run = np.empty(n, dtype=bool)
    new_run[0] = True
    np.not_equal(s[1:], s[:-1], out=new_run[1:])
    starts = np.flatnonzero(new_run)
    ends = np.append(starts[1:], n)
    run_id = np.cumsum(new_run) - 1
    k_sorted = (starts + ends)[run_id]
    k = np.empty(n, dtype=np.int64)
    k[order] = k_sorted
    return k


def rank_normalize(chains: np.ndarray):
    """``_rank_normalize`` (``diagnostics.py:101-133``) on a [C, N] array.

    Returns (avg_rank[C, N], z[C, N]).
    """
    chains = np.asarray(chains, dtype=np.float64)
    m = chains.size
    if m == 0:
        return chains.copy(), chains.copy()
    k = rank_codes(chains.reshape(-1))
    z = z_table(m)[k]
    return ((k + 1) / 2.0).reshape(chains.shape), z.reshape(chains.shape)


def pooled_median(flat: np.ndarray) -> float:
    """``statistics.median`` as used by ``_fold_chains`` (``diagnostics.py:97``)."""
    s = np.sort(np.asarray(flat, dtype=np.float64))
    n = s.size
    if n % 2 == 1:
        return float(s[n // 2])
    return float((s[n // 2 - 1] + s[n // 2]) / 2)


def fold_chains(chains: np.ndarray) -> np.ndarray:
    """``_fold_chains`` (``diagnostics.py:93-98``): ``abs(v - median(all))``."""
    chains = np.asarray(chains, dtype=np.float64)
    if chains.size == 0:
        return chains.copy()
    med = pooled_median(chains.reshape(-1))
    return np.abs(chains - med)


def split_chains(chains: np.ndarray) -> np.ndarray:
    """``_split_chains`` (``diagnostics.py:76-85``): [C, N] -> [2C, N//2]."""
    chains = np.asarray(chains, dtype=np.float64)
    c, n = chains.shape
    half = n // 2
    if half == 0:
        return np.empty((0, 0))
    return chains[:, : 2 * half].reshape(c * 2, half)


# --------------------------------------------------------------------------
# R-hat / ESS
# --------------------------------------------------------------------------

def _variance_rows(a: np.ndarray) -> np.ndarray:
    """``_variance`` (``diagnostics.py:196-201``) per row: two-pass, ddof=1."""
    n = a.shape[1]
    if n < 2:
        return np.zeros(a.shape[0])
    mean = a.sum(axis=1) / n
    var = ((a - mean[:, None]) ** 2).sum(axis=1) / (n - 1)
    # A constant row has variance exactly 0 in the reference (Python's compensated sum()
    # returns n*c exactly, so mean == c); numpy's pairwise mean can be 1 ulp off and would
    # leave ~1e-33 here, turning the reference's R-hat = inf into ~1e16.
    var[a.max(axis=1) == a.min(axis=1)] = 0.0
    return var


def rhat(chains: np.ndarray) -> float:
    """``_rhat`` (``diagnostics.py:136-151``) on equal-length rows."""
    m = chains.shape[0]
    if m < 2:
        return math.nan
    n = chains.shape[1]
    if n < 2:
        return math.nan
    means = chains.sum(axis=1) / n
    mean_total = means.sum() / m
    var_between = n * ((means - mean_total) ** 2).sum() / (m - 1)
    var_within = _variance_rows(chains).sum() / m
    var_hat = (n - 1) / n * var_within + var_between / n
    if var_within == 0:
        return 1.0 if var_between == 0 else math.inf
    return math.sqrt(var_hat / var_within)


def ess(chains: np.ndarray, return_lag: bool = False):
    """``_ess`` + ``_autocorr`` (``diagnostics.py:154-193``).

    Unsplit chains, autocovariance normalised by (n - lag), sum of rho_t
    truncated at the first negative rho_t (that rho excluded).
    """
    m = chains.shape[0]
    if m == 0:
        return (math.nan, 0) if return_lag else math.nan
    n = chains.shape[1]
    if n < 2:
        return (math.nan, 0) if return_lag else math.nan
    means = chains.sum(axis=1) / n
    mean_total = means.sum() / m
    var_between = n * ((means - mean_total) ** 2).sum() / (m - 1) if m > 1 else 0.0
    var_within = _variance_rows(chains).sum() / m
    var_hat = (n - 1) / n * var_within + var_between / n
    if var_hat == 0:
        return (float(m * n), 0) if return_lag else float(m * n)
    y = chains - means[:, None]
    rho_sum = 0.0
    lag = 1
    while lag < n:
        cov = (y[:, : n - lag] * y[:, lag:]).sum(axis=1) / (n - lag)
        rho = cov.sum() / (m * var_hat)
        if rho < 0:
            break
        rho_sum += rho
        lag += 1
    val = m * n / (1 + 2 * rho_sum)
    return (val, lag) if return_lag else val


def ess_geyer(chains: np.ndarray, return_lag: bool = False):
    """EXTENSION (north-star "Geyer initial-monotone truncation"; the reference truncates at the
    first negative rho_t, ``diagnostics.py:172-176`` -- parity unpinned, this function IS the definition).

    Same rho_t as ``ess`` (unsplit chains, (n - t)-normalised, same var_hat) with rho_0 := 1.
    Pairs P_k = rho_{2k} + rho_{2k+1}; the sum stops before the first P_k < 0 (initial positive
    sequence) and every P_k is capped by its predecessor (initial monotone sequence); only complete
    pairs (2k + 1 <= n - 1) count.  tau = max(-1 + 2 * sum_k P_k, 1 / log10(m * n)) (the lower bound
    keeps antithetic chains finite, as in Stan), ESS = m * n / tau.
    Returned lag = 2k of the first excluded pair (n if the pairs ran out).
    """
    m = chains.shape[0]
    if m == 0:
        return (math.nan, 0) if return_lag else math.nan
    n = chains.shape[1]
    if n < 2:
        return (math.nan, 0) if return_lag else math.nan
    means = chains.sum(axis=1) / n
    mean_total = means.sum() / m
    var_between = n * ((means - mean_total) ** 2).sum() / (m - 1) if m > 1 else 0.0
    var_within = _variance_rows(chains).sum() / m
    var_hat = (n - 1) / n * var_within + var_between / n
    if var_hat == 0:
        return (float(m * n), 0) if return_lag else float(m * n)
    y = chains - means[:, None]

    def rho(t):
        cov = (y[:, : n - t] * y[:, t:]).sum(axis=1) / (n - t)
        return cov.sum() / (m * var_hat)

    total, prev, k = 0.0, math.inf, 0
    stop = n
    while 2 * k + 1 <= n - 1:
        pair = (1.0 if k == 0 else rho(2 * k)) + rho(2 * k + 1)
        if pair < 0:
            stop = 2 * k
            break
        pair = min(pair, prev)
        total += pair
        prev = pair
        k += 1
    tau = max(-1.0 + 2.0 * total, 1.0 / math.log10(m * n))
    val = m * n / tau
    return (val, stop) if return_lag else val


def _guard(chains, min_chains: int, what: str):
    """Guards shared by the three public functions (``diagnostics.py:24-30``)."""
    if min_chains < 1:
        raise ValueError(f"min_chains must be >= 1; got {min_chains}")
    if len(chains) < min_chains:
        raise ValueError(
            f"{what} diagnostics require at least {min_chains} chains; got {len(chains)} chain(s)"
        )
    return len(chains) < 2


def split_rhat(chains, *, min_chains: int = 4) -> float:
    """``split_rhat`` (``diagnostics.py:13-40``)."""
    if _guard(chains, min_chains, "R-hat"):
        return math.nan
    x = np.asarray(chains, dtype=np.float64)
    _, z = rank_normalize(x)
    rb = rhat(split_chains(z))
    _, zf = rank_normalize(fold_chains(x))
    rt = rhat(split_chains(zf))
    return max(rb, rt)


def ess_bulk(chains, *, min_chains: int = 4) -> float:
    """``ess_bulk`` (``diagnostics.py:43-56``)."""
    if _guard(chains, min_chains, "ESS"):
        return math.nan
    _, z = rank_normalize(np.asarray(chains, dtype=np.float64))
    return ess(z)


def ess_tail(chains, *, min_chains: int = 4) -> float:
    """``ess_tail`` (``diagnostics.py:59-73``): ESS of the folded rank-normalised draws."""
    if _guard(chains, min_chains, "ESS"):
        return math.nan
    _, zf = rank_normalize(fold_chains(np.asarray(chains, dtype=np.float64)))
    return ess(zf)


def diagnostics_param(x: np.ndarray) -> dict:
    """All three diagnostics of one [C, N] parameter, sharing the two rank passes."""
    x = np.asarray(x, dtype=np.float64)
    _, z = rank_normalize(x)
    _, zf = rank_normalize(fold_chains(x))
    rb = rhat(split_chains(z))
    rt = rhat(split_chains(zf))
    eb, lag_b = ess(z, return_lag=True)
    et, lag_t = ess(zf, return_lag=True)
    return {"rhat": max(rb, rt), "rhat_bulk": rb, "rhat_tail": rt,
            "ess_bulk": eb, "ess_tail": et, "lag_bulk": lag_b, "lag_tail": lag_t}


def diagnostics_batch(x: np.ndarray) -> dict:
    """[P, C, N] -> dict of [P] arrays (rhat, ess_bulk, ess_tail)."""
    x = np.asarray(x, dtype=np.float64)
    rows = [diagnostics_param(x[p]) for p in range(x.shape[0])]
    return {k: np.array([r[k] for r in rows]) for k in rows[0]} if rows else {}


# --------------------------------------------------------------------------
# stats backend / compare / checks
# --------------------------------------------------------------------------

def quantile_lerp(sorted_vals: np.ndarray, q: float):
    """numpy's linear (H&F 7) quantile, ``np.quantile`` as called at ``backends_numpy.py:44``.

    Returns (value, lo_index, s[lo], s[hi]).  Follows numpy
    ``_function_base_impl._quantile`` / ``_lerp``: virtual index (n-1)*q,
    ``a + (b-a)*g`` for g < 0.5 else ``b - (b-a)*(1-g)``.
    """
    n = sorted_vals.size
    vi = (n - 1) * q
    lo = int(math.floor(vi))
    g = vi - lo
    hi = min(lo + 1, n - 1)
    a = float(sorted_vals[lo])
    b = float(sorted_vals[hi])
    diff = b - a
    val = a + diff * g if g < 0.5 else b - diff * (1 - g)
    return val, lo, a, b


def numpy_stats(flat: np.ndarray, quantiles=(0.05, 0.5, 0.95)) -> dict:
    """``NumpyBackend.stats`` for one column (``backends_numpy.py:40-47``)."""
    data = np.asarray(flat, dtype=np.float64)
    entry = {"mean": float(np.mean(data)), "std": float(np.std(data, ddof=0))}
    q_vals = np.quantile(data, list(quantiles))
    for q, v in zip(quantiles, q_vals):
        entry[f"q{int(q * 100)}"] = float(v)
    return entry


def compute_basic_stats(values) -> dict:
    """``compute_basic_stats`` (``compare.py:58-64``)."""
    v = np.asarray(values, dtype=np.float64)
    n = v.size
    if n == 0:
        return {"mean": math.nan, "std": math.nan}
    mean = v.sum() / n
    var = ((v - mean) ** 2).sum() / n
    return {"mean": float(mean), "std": float(var ** 0.5)}


def compare_arrays(ref: np.ndarray, act: np.ndarray, tolerance: float):
    """Vector form of the inner loop of ``compare_stats`` (``compare.py:38-43``).

    rel = |a - r| / max(|r|, 1e-12); passed = rel <= tol (NaN fails).
    """
    ref = np.asarray(ref, dtype=np.float64)
    act = np.asarray(act, dtype=np.float64)
    denom = np.maximum(np.abs(ref), 1e-12)
    # Python's max(abs(nan), 1e-12) returns nan (first arg kept when compare is False)
    denom = np.where(np.isnan(ref), np.nan, denom)
    with np.errstate(invalid="ignore"):
        rel = np.abs(act - ref) / denom
        passed = rel <= tolerance
    return rel, passed


def compare_stats(ref_stats, actual_stats, tolerance, metrics):
    """``compare_stats`` (``compare.py:24-55``) -> (passed, details, failures) plain tuples."""
    details = {}
    failures = []
    for param, stats in ref_stats.items():
        if param not in actual_stats:
            failures.append(f"missing param: {param}")
            continue
        pd = {}
        for metric in metrics:
            r = float(stats.get(metric, math.nan))
            a = float(actual_stats[param].get(metric, math.nan))
            denom = max(abs(r), 1e-12)
            rel = abs(a - r) / denom
            ok = rel <= tolerance
            if not ok:
                failures.append(f"{param}.{metric} rel_error={rel:.3g} > {tolerance}")
            pd[metric] = (r, a, rel, ok)
        details[param] = pd
    return len(failures) == 0, details, failures


def checks(n_chains: int, n_draws: int, diag: dict) -> dict:
    """``_checks`` (``convert.py:164-172``)."""
    return {
        "ndraws_is_10k": n_chains * n_draws == 10_000,
        "nchains_is_gte_4": n_chains >= 4,
        "ess_above_400": all(v.get("ess_bulk", 0.0) > 400 for v in diag.values()),
        "rhat_below_1_01": all(v.get("rhat", 1.0) < 1.01 for v in diag.values()),
    }


def zscore_extension(mean, std, ess_bulk, ref_mean, ref_std, mref: int, z_crit: float = 4.0) -> dict:
    """EXTENSION (no reference code: ``compare.py`` is relative-error only; parity unpinned).

    mcse_mean = std / sqrt(ess_bulk); ref_se_mean = ref_std / sqrt(Mref) (the packaged reference
    draws are thinned, ``provenance.py:17-22``: taken as independent);
    z_mean = (mean - ref_mean) / sqrt(mcse_mean**2 + ref_se_mean**2); z_pass = |z_mean| <= z_crit
    (NaN fails).  Every operation is one IEEE rounding, in this order.
    """
    with np.errstate(all="ignore"):
        se_c = np.asarray(std, dtype=np.float64) / np.sqrt(np.asarray(ess_bulk, dtype=np.float64))
        se_r = np.asarray(ref_std, dtype=np.float64) / np.sqrt(np.float64(mref))
        z = (np.asarray(mean, dtype=np.float64) - np.asarray(ref_mean, dtype=np.float64)) / np.sqrt(se_c * se_c + se_r * se_r)
        return {"mcse_mean": se_c, "ref_se_mean": se_r, "z_mean": z, "z_pass": np.abs(z) <= z_crit}


def full_pass(cand: np.ndarray, ref: np.ndarray, tolerance: float = 0.15,
              quantiles=(0.05, 0.5, 0.95), z_crit: float = 4.0) -> dict:
    """The benchmarked unit of work for P params:

    candidate [P, C, N]: mean/std (compare.py:58-64), quantiles + rhat/ess_bulk/ess_tail;
    reference [P, Mref]: NumpyBackend.stats; compare on mean and std at ``tolerance``.
    Returns dict of [P]-arrays.
    """
    cand = np.asarray(cand, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    p = cand.shape[0]
    out = {k: np.empty(p) for k in
           ("rhat", "ess_bulk", "ess_tail", "mean", "std", "ref_mean", "ref_std",
            "rel_mean", "rel_std")}
    qk = [f"q{int(q * 100)}" for q in quantiles]
    for k in qk:
        out[k] = np.empty(p)
        out["ref_" + k] = np.empty(p)
    out["pass_mean"] = np.empty(p, dtype=bool)
    out["pass_std"] = np.empty(p, dtype=bool)
    for i in range(p):
        d = diagnostics_param(cand[i])
        out["rhat"][i], out["ess_bulk"][i], out["ess_tail"][i] = d["rhat"], d["ess_bulk"], d["ess_tail"]
        flat = cand[i].reshape(-1)
        bs = compute_basic_stats(flat)
        out["mean"][i], out["std"][i] = bs["mean"], bs["std"]
        cs = numpy_stats(flat, quantiles)
        rs = numpy_stats(ref[i], quantiles)
        for k in qk:
            out[k][i] = cs[k]
            out["ref_" + k][i] = rs[k]
        out["ref_mean"][i], out["ref_std"][i] = rs["mean"], rs["std"]
    out["rel_mean"], out["pass_mean"] = compare_arrays(out["ref_mean"], out["mean"], tolerance)
    out["rel_std"], out["pass_std"] = compare_arrays(out["ref_std"], out["std"], tolerance)
    out["passed"] = bool(out["pass_mean"].all() and out["pass_std"].all())
    out.update(zscore_extension(out["mean"], out["std"], out["ess_bulk"], out["ref_mean"], out["ref_std"],
                                ref.shape[1], z_crit))
    return out
