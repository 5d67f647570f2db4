"""Small-shape walk through every kernel of the library, for `compute-sanitizer` (memcheck / racecheck / synccheck).

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

Shapes are tiny (the tools slow kernels down 10-100x) but reach every path: distribution sort + on-chip finish, the
oversize-bucket kernel, pure buckets (heavy ties), the LSD fallback, the wide (long-chain) partition with the binned
score scatter, the fold / merge, round 0 and the persistent later rounds of the ESS kernel (sticky chains), the
statistics-only pass, compare / checks / summary, ingestion (row order, gather, transposing load) and the tcgen05
GLM kernel.  Results are checked against the CPU oracle so a silent corruption also fails the run.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "mcmc-db_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from mcmc_ref_b200 import batch as B  # noqa: E402
from mcmc_ref_b200 import logdensity  # noqa: E402
from oracle import glm_oracle  # noqa: E402
from oracle import mcmc_oracle as O  # noqa: E402


def ar1(rng, c, n, phi):
    from scipy.signal import lfilter

    e = rng.standard_normal((c, n))
    return lfilter([np.sqrt(1 - phi * phi)], [1.0, -phi], e, axis=1)


def check(x, label, long=False):
    d = B.diagnostics_batch(x[None])
    want = O.diagnostics_param(x)
    for k in ("rhat", "ess_bulk", "ess_tail"):
        g, w = float(d[k][0].item()), want[k]
        assert (np.isnan(g) and np.isnan(w)) or abs(g - w) <= 1e-9 * max(abs(w), 1e-300) or g == w, (label, k, g, w)
    assert float(d["quantiles"][0, 1].item()) == O.numpy_stats(x.reshape(-1))["q50"], label
    if not long:
        r = B.ranks_batch(x[None])
        rank, z = O.rank_normalize(x)
        assert np.array_equal(r["avg_rank"][0].cpu().numpy(), rank) and np.array_equal(r["z"][0].cpu().numpy(), z), label
    print("ok", label, flush=True)


def main() -> None:
    assert torch.cuda.is_available()
    rng = np.random.default_rng(4711)
    check(rng.standard_normal((4, 1000)), "continuous 4x1000")
    check(np.round(rng.standard_normal((4, 1500)), 1), "tied 4x1500")
    check(ar1(rng, 4, 3000, 0.97), "sticky 4x3000 (persistent rounds)")
    check(np.where(rng.random((4, 5000)) < 0.5, 0.0, rng.standard_normal((4, 5000))), "zero-inflated (pure bucket)")
    check(rng.integers(0, 3, (8, 2500)).astype(float), "3 values x ~6700 (oversize / fallback)")
    check(np.full((4, 300), 2.5), "constant")
    check(rng.standard_normal((2, 5)), "n5_c2")
    check(rng.standard_normal((2, 200_000)), "wide path 2x200k", long=True)
    cand = B.synth_fill(12, 4, 700, kind=0)
    ref = B.synth_fill(12, 1, 500, kind=1).reshape(12, 500)
    summary, ok = B.full_pass(cand, ref)
    want = O.full_pass(cand.cpu().numpy(), ref.cpu().numpy(), tolerance=0.15)
    assert ok == want["passed"]
    np.testing.assert_allclose(summary.cpu().numpy()[:, 0], want["rhat"], rtol=1e-9)
    sh, okh = B.full_pass_host(cand.cpu().pin_memory(), ref.cpu().pin_memory())
    assert np.array_equal(sh.numpy(), summary.cpu().numpy(), equal_nan=True)
    print("ok full pass (device + host)", flush=True)
    x = rng.standard_normal((7, 3, 501))
    rows = np.ascontiguousarray(x.reshape(7, -1).T)
    a, b = B.diagnostics_batch(x), B.diagnostics_rows(rows, n_chains=3)
    assert all(np.array_equal(a[k].cpu().numpy(), b[k].cpu().numpy(), equal_nan=True) for k in a)
    print("ok rows layout", flush=True)
    import pyarrow as pa

    from mcmc_ref_b200.ingest import table_to_device

    perm = rng.permutation(3 * 501)
    chain, draw = np.repeat(np.arange(3), 501)[perm], np.tile(np.arange(501), 3)[perm]
    t = pa.table({"chain": chain, "draw": draw, "a": x[0].reshape(-1)[perm], "b": x[1].reshape(-1)[perm]})
    xd, c, n = table_to_device(t, ["a", "b"])
    assert (c, n) == (3, 501) and np.array_equal(xd.cpu().numpy(), x[:2])
    print("ok ingestion", flush=True)
    X = rng.standard_normal((700, 40)) / 6.0
    beta = rng.standard_normal((200, 40))
    y = (rng.random(700) < 0.5).astype(np.float64)
    lp = logdensity.logistic_lp(X, y, beta).cpu().numpy()
    np.testing.assert_allclose(lp, glm_oracle.logistic_lp(X, y, beta), rtol=1e-6)
    sig = 0.5 + rng.random(200)
    lpn = logdensity.glm_lp("normal", X, X @ beta[0] + 0.3 * rng.standard_normal(700), beta, sigma=sig).cpu().numpy()
    assert np.isfinite(lpn).all()
    print("ok glm", flush=True)
    torch.cuda.synchronize()
    print("sanitize_smoke: all paths ran")


if __name__ == "__main__":
    main()
