"""Multi-core driver for the CPU oracle (TEST / BENCH INFRASTRUCTURE ONLY).

Used by bench.py's `cpu_baseline` leg and by `bench.py --impl reference`: fans the
per-parameter oracle pass (oracle/mcmc_oracle.py: full diagnostics + stats + compare
inputs) out over host cores, the way BASELINE.md section 4 prescribes for the reference
itself (parallel over parameters).  Inputs travel through POSIX shared memory so the
workers can be spawned (no fork after CUDA initialisation).
"""

from __future__ import annotations

import multiprocessing as mp
import os
import time
from multiprocessing import shared_memory

import numpy as np

KEYS = ("rhat", "ess_bulk", "ess_tail", "mean", "std", "q5", "q50", "q95",
        "ref_mean", "ref_std", "ref_q5", "ref_q50", "ref_q95")


def _worker(args):
    cand_name, cand_shape, ref_name, ref_shape, lo, hi = args
    from oracle import mcmc_oracle as O

    sc = shared_memory.SharedMemory(name=cand_name)
    sr = shared_memory.SharedMemory(name=ref_name)
    try:
        cand = np.ndarray(cand_shape, dtype=np.float64, buffer=sc.buf)
        ref = np.ndarray(ref_shape, dtype=np.float64, buffer=sr.buf)
        rows = np.empty((hi - lo, len(KEYS)))
        for j, p in enumerate(range(lo, hi)):
            d = O.diagnostics_param(cand[p])
            flat = cand[p].reshape(-1)
            bs = O.compute_basic_stats(flat)
            cs = O.numpy_stats(flat)
            rs = O.numpy_stats(ref[p])
            rows[j] = (d["rhat"], d["ess_bulk"], d["ess_tail"], bs["mean"], bs["std"], cs["q5"], cs["q50"], cs["q95"],
                       rs["mean"], rs["std"], rs["q5"], rs["q50"], rs["q95"])
        return lo, rows
    finally:
        sc.close()
        sr.close()


def _ref_worker(args):
    """One parameter block through the REFERENCE ITSELF (oracle/_ref/mcmc_ref, the unmodified package):
    split_rhat + ess_bulk + ess_tail (diagnostics.py:13-73), NumpyBackend.stats of candidate and reference
    draws (backends_numpy.py:17-49), compute_stats_from_draws + compare_stats (compare.py:24-68) --
    BASELINE.md section 4, step 2."""
    cand_name, cand_shape, ref_name, ref_shape, lo, hi, tol = args
    import pyarrow as pa

    from oracle.build_ref import import_reference

    if import_reference() is None:
        raise RuntimeError("oracle/_ref is not built (python oracle/build_ref.py)")
    from mcmc_ref import compare as rc
    from mcmc_ref import diagnostics as rd
    from mcmc_ref.backends_numpy import NumpyBackend

    sc = shared_memory.SharedMemory(name=cand_name)
    sr = shared_memory.SharedMemory(name=ref_name)
    try:
        cand = np.ndarray(cand_shape, dtype=np.float64, buffer=sc.buf)
        ref = np.ndarray(ref_shape, dtype=np.float64, buffer=sr.buf)
        rows = np.empty((hi - lo, len(KEYS)))
        ok = True
        be = NumpyBackend()
        for j, p in enumerate(range(lo, hi)):
            chains = cand[p].tolist()
            flat = [v for ch in chains for v in ch]
            rh, eb, et = rd.split_rhat(chains), rd.ess_bulk(chains), rd.ess_tail(chains)
            cs = be.stats(pa.table({"x": cand[p].reshape(-1)}), ["x"])["x"]
            rs = be.stats(pa.table({"x": ref[p]}), ["x"])
            act = rc.compute_stats_from_draws({"x": flat})
            res = rc.compare_stats(rs, act, tol, ("mean", "std"))
            ok = ok and res.passed
            rows[j] = (rh, eb, et, act["x"]["mean"], act["x"]["std"], cs["q5"], cs["q50"], cs["q95"],
                       rs["x"]["mean"], rs["x"]["std"], rs["x"]["q5"], rs["x"]["q50"], rs["x"]["q95"])
        return lo, rows, ok
    finally:
        sc.close()
        sr.close()


class OraclePool:
    """Spawned worker pool + shared-memory staging of one sample."""

    def __init__(self, cores: int | None = None) -> None:
        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.cores)
        self._shm: list[shared_memory.SharedMemory] = []
        # make sure the workers have imported numpy + the oracle before anything is timed
        self.pool.map(_warm, range(self.cores))

    def stage(self, cand: np.ndarray, ref: np.ndarray):
        self.release()
        out = []
        for a in (cand, ref):
            a = np.ascontiguousarray(a, dtype=np.float64)
            shm = shared_memory.SharedMemory(create=True, size=max(a.nbytes, 8))
            np.ndarray(a.shape, dtype=np.float64, buffer=shm.buf)[...] = a
            self._shm.append(shm)
            out.append((shm.name, a.shape))
        self._staged = out
        return out

    def run(self, tolerance: float = 0.15, chunk: int = 1, impl: str = "port"):
        """One pass over the staged sample: impl "port" = oracle/mcmc_oracle.py (vectorised numpy restatement),
        "reference" = the unmodified reference package in oracle/_ref.  Returns (seconds, dict of [P] arrays)."""
        from oracle import mcmc_oracle as O

        (cn, cs), (rn, rs) = self._staged
        p = cs[0]
        t0 = time.perf_counter()
        rows = np.empty((p, len(KEYS)))
        if impl == "reference":
            tasks = [(cn, cs, rn, rs, lo, min(lo + chunk, p), tolerance) for lo in range(0, p, chunk)]
            for lo, r, _ in self.pool.map(_ref_worker, tasks, chunksize=1):
                rows[lo:lo + len(r)] = r
        else:
            tasks = [(cn, cs, rn, rs, lo, min(lo + chunk, p)) for lo in range(0, p, chunk)]
            for lo, r in self.pool.map(_worker, tasks, chunksize=1):
                rows[lo:lo + len(r)] = r
        out = {k: rows[:, i] for i, k in enumerate(KEYS)}
        out["rel_mean"], out["pass_mean"] = O.compare_arrays(out["ref_mean"], out["mean"], tolerance)
        out["rel_std"], out["pass_std"] = O.compare_arrays(out["ref_std"], out["std"], tolerance)
        out["passed"] = bool(out["pass_mean"].all() and out["pass_std"].all())
        return time.perf_counter() - t0, out

    def release(self):
        for shm in self._shm:
            shm.close()
            shm.unlink()
        self._shm = []

    def close(self):
        self.release()
        self.pool.close()
        self.pool.join()


def _warm(_):
    from oracle import mcmc_oracle  # noqa: F401

    return os.getpid()


def host_workload(p: int, c: int, n: int, mref: int, seed: int = 4711, p0: int = 0):
    """numpy version of the SURVEY.md 8(d) workload family (same recipe as csrc/synth.cuh,
    different random stream): candidate [p, c, n], reference [p, mref]."""
    from scipy.signal import lfilter

    cand = np.empty((p, c, n))
    ref = np.empty((p, mref))
    for i in range(p):
        gp = p0 + i
        rng = np.random.Generator(np.random.Philox(key=[seed, gp]))
        mu, sigma = gp % 17 - 8.0, 2.0 ** (gp % 5 - 2)
        phi = (0.0, 0.5, 0.9)[gp % 3]
        e = rng.standard_normal((c, n))
        e[:, 1:] *= np.sqrt(1 - phi * phi)
        a = lfilter([1.0], [1.0, -phi], e, axis=1) if phi else e
        x = mu + sigma * a
        if gp % 97 == 96:
            x[1::2] += 0.5 * sigma
        if gp % 101 == 100:
            x = np.round(x, 2)
        cand[i] = x
        shift = 0.3 * sigma if gp % 89 == 88 else 0.0
        ref[i] = mu + shift + sigma * rng.standard_normal(mref)
    return cand, ref
