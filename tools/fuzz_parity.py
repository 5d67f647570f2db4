"""Randomised parity sweep of the diagnostics path against the oracle: odd shapes (1 .. 9 chains, chain lengths around
the tile / sub-tile / split boundaries), ties, constant stretches, sticky and anti-correlated chains, offsets between
chains, heavy tails, several parameters of different character per batch.  Prints the first mismatch and exits 1.

    python tools/fuzz_parity.py [cases] [seed]
"""
import math
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "mcmc-db_b200")]
import numpy as np
from scipy.signal import lfilter

from mcmc_ref_b200 import batch as B
from oracle import mcmc_oracle as O

TOL = 1e-9


def close(a, b, tol):
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= tol * max(1.0, abs(a), abs(b))


def make_param(rng, c, n):
    kind = rng.integers(0, 9)
    phi = [0.0, 0.5, 0.9, 0.97, -0.6, 0.0, 0.9, 0.3, 0.0][kind]
    e = rng.standard_normal((c, n))
    if kind == 7:
        e = rng.standard_cauchy((c, n))
    x = lfilter([math.sqrt(1 - phi * phi)], [1.0, -phi], e, axis=1) if phi else e
    if kind == 5:
        x = np.round(x, 1)                      # heavy ties
    if kind == 6:
        x = np.round(x, 2) + 0.25 * (np.arange(c) % 2)[:, None]  # ties + offsets between chains
    if kind == 8 and n >= 4:
        x[:, : n // 2] = 1.0                    # constant first halves
        x[0] = 2.0
    if rng.integers(0, 4) == 0:
        x = x * 10.0 ** rng.integers(-3, 4) + rng.integers(-5, 6)
    return x


def run(cases: int, seed: int) -> int:
    rng = np.random.default_rng(seed)
    lengths = [4, 5, 8, 9, 15, 16, 17, 24, 31, 32, 33, 40, 47, 48, 49, 50, 56, 63, 64, 65, 96, 127, 128, 129, 136, 200, 255, 256,
               257, 264, 511, 512, 1000, 1016, 1024, 2040, 2047, 2048, 2049, 2056, 4088, 4096, 4104, 6152, 8192, 10000]
    bad = 0
    for k in range(cases):
        c = int(rng.integers(1, 10))
        n = int(lengths[rng.integers(0, len(lengths))])
        p = int(rng.integers(1, 5))
        x = np.stack([make_param(rng, c, n) for _ in range(p)])
        d = B.diagnostics_batch(x)
        if k % 3 == 0 and c * n <= 40000:
            # the integer / table-lookup class bit for bit: average ranks and normal scores, bulk and folded
            r = B.ranks_batch(x)
            for i in range(p):
                rank, z = O.rank_normalize(x[i])
                rank_f, z_f = O.rank_normalize(O.fold_chains(x[i]))
                same = (np.array_equal(r["avg_rank"][i].cpu().numpy(), rank) and np.array_equal(r["z"][i].cpu().numpy(), z)
                        and np.array_equal(r["avg_rank_folded"][i].cpu().numpy(), rank_f)
                        and np.array_equal(r["z_folded"][i].cpu().numpy(), z_f))
                if not same:
                    bad += 1
                    print(f"RANK MISMATCH case {k} param {i}: C={c} N={n} P={p}", flush=True)
        for i in range(p):
            want = O.diagnostics_param(x[i])
            got = {key: float(d[key][i].item()) for key in ("rhat", "ess_bulk", "ess_tail")}
            ok = all(close(got[key], want[key], TOL) for key in got)
            # (no walk at all -- zero variance, one draw: the oracle reports lag 0 there, the device N)
            # A correlation that is exactly zero in exact arithmetic (tiny symmetric inputs, e.g. one chain of 5 draws)
            # gets its sign from the summation order: the walk may then stop one lag earlier or later with the SAME
            # sum -- accepted only if the ESS agrees to 1e-12.
            for lk, ek in (("lag_bulk", "ess_bulk"), ("lag_tail", "ess_tail")):
                same = want[lk] == 0 or int(d[lk][i].item()) == want[lk]
                ok = ok and (same or close(got[ek], want[ek], 1e-12))
            ns = O.numpy_stats(x[i].reshape(-1))
            ok = ok and all(float(d["quantiles"][i, j].item()) == ns[q] for j, q in enumerate(("q5", "q50", "q95")))
            if not ok:
                bad += 1
                print(f"MISMATCH case {k} param {i}: C={c} N={n} P={p} got {got} lags "
                      f"{int(d['lag_bulk'][i].item())}/{int(d['lag_tail'][i].item())} want {want}", flush=True)
                if bad >= 5:
                    return bad
    print(f"fuzz: {cases} cases, seed {seed}: {'OK' if not bad else str(bad) + ' mismatches'}")
    return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 120, int(sys.argv[2]) if len(sys.argv) > 2 else 20261019) else 0)
