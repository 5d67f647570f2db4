#!/usr/bin/env python
"""Benchmark of the diagnostics + compare pass (BASELINE.json metric: draw*params/s).

  python bench.py --gpus N --steps K --warmup W            # B200 arm (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the reference algorithm on host cores

A step = one full pass over the workload's parameter shard: sort -> tie ranks -> normal
scores (raw + folded) -> split R-hat, bulk/tail ESS, mean/std/quantiles of the candidate
draws; mean/std/quantiles of the reference draws; compare verdicts (tol 0.15); for N > 1
the NCCL gather of the [P, 20] summaries and the AND of the verdicts.

Workloads (per GPU, weak scaling: parameters are sharded by column):
  c5tile  12500 params x 64 chains x 4096 draws (1/8 of configs[4]; 8 GPUs = the full config)  <- DEFAULT: the
          configuration the north-star target is quoted on; the default line also carries "secondary" blocks
          with config 2 (N = 1 only) and config 3 (GLM log-density, draw-block sharded over the N GPUs)
  c2      1000 params x 16 chains x 10000 draws + 10x1000 reference draws  (BASELINE configs[1])
  c4      64 params x 8 chains x 1M draws (BASELINE configs[3], long-chain ESS stress)
  tiny    64 x 4 x 1000 (self-test)
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for _p in (str(ROOT), str(ROOT / "mcmc-db_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    "c2": dict(P=1000, C=16, N=10000, RC=10, RN=1000,
               name="c2: 1000 params x 16 chains x 10000 draws vs 10x1000 reference draws, per GPU"),
    "c5tile": dict(P=12500, C=64, N=4096, RC=10, RN=1000,
                   name="c5 shard: 12500 params x 64 chains x 4096 draws per GPU (8 GPUs = 100k params)"),
    "c4": dict(P=64, C=8, N=1_000_000, RC=10, RN=1000,
               name="c4: long-chain ESS stress, 64 params x 8 chains x 1M draws per GPU"),
    "tiny": dict(P=64, C=4, N=1000, RC=10, RN=100, name="tiny self-test 64 x 4 x 1000"),
}
METRIC = "diag+compare draw*params/s"
UNIT = "draw*params/s"
TOL = 0.15


def measured_peak_gbs() -> tuple[float, str]:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.startswith("Active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_config(w: dict, world: int) -> dict:
    """The `config` object of the JSON line: identical in the B200 arm and the reference arm."""
    return {"workload": w["name"], "params_per_gpu": w["P"], "chains": w["C"], "draws": w["N"],
            "ref_draws": w["RC"] * w["RN"], "tolerance": TOL, "sharding": f"param columns x{world}",
            "l2": f"inputs larger than L2 ({algorithmic_bytes(w) / 1e9:.2f} GB read per step per GPU)"}


def algorithmic_bytes(w: dict) -> int:
    """every fp64 draw read once + the summary row (20 doubles) per parameter."""
    return 8 * w["P"] * (w["C"] * w["N"] + w["RC"] * w["RN"]) + 8 * 20 * w["P"]


def _dist_setup():
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def diag_measure(args, wname: str, rank: int, world: int, local: int, full: bool) -> dict:
    """One diagnostics + compare workload on this rank's parameter shard.  full=False: the device-timed value
    and the roofline blocks only (secondary block of the default line)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from mcmc_ref_b200 import _native, batch, sharding

    w = WORKLOADS[wname]
    P, C, N, MREF = w["P"], w["C"], w["N"], w["RC"] * w["RN"]
    cand = batch.synth_fill(P, C, N, p0=rank * P, kind=0)
    ref = batch.synth_fill(P, w["RC"], w["RN"], p0=rank * P, kind=1).reshape(P, MREF)
    ws = batch.Workspace()
    ws_cap = int(args.ws_gb * 2**30)
    steps = args.steps if full else min(args.steps, 10)

    def step():
        # this rank's parameter block, then ONE collective: the all_gather of the [P/G, 20] summaries (NCCL).
        # The global verdict is the AND of the gathered flags column, left on the device (read after the loop).
        summary, passed = batch.full_pass(cand, ref, tolerance=TOL, workspace=ws, ws_cap_bytes=ws_cap)
        if world > 1:
            gathered = sharding.gather_summaries(summary, world * P)
            passed = sharding.verdict_from_summary(gathered)
        return summary, passed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        summary, passed = step()

    # ---- timed region (device clock on the launching stream, max over ranks) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = _native.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        summary, passed = step()
    e1.record()
    barrier()
    launches = _native.kernel_launches() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    nl = torch.tensor([launches], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
    clocks = sampler.stop() if rank == 0 else None
    passed = bool(passed)
    total_ms = float(ms.item())
    dp_per_step = world * P * C * N
    value = dp_per_step * steps / (total_ms * 1e-3)

    # ---- per-kernel device time over the same steps (events around every launch) ----
    _native.profile(True)
    _native.profile_read(reset=True)
    barrier()
    for _ in range(steps):
        step()
    barrier()
    prof = _native.profile_read(reset=True)
    _native.profile(False)
    peak, peak_src = measured_peak_gbs()
    tot_kernel_ms = sum(v[0] for v in prof.values()) or 1.0
    # the LSD fallback kernels touch only the flagged parameters (a data-dependent handful), so
    # the per-launch algorithmic bytes below would not describe them: they never head the roofline
    main_path = {k: v for k, v in prof.items() if k not in ("k_hist0", "k_radix_pass", "k_rank_codes")}
    dom = max(main_path, key=lambda k: main_path[k][0])
    dom_ms, dom_n = prof[dom]
    # algorithmic bytes of the launches of the dominant kernel: 8 B x draws each launch processes: per step a
    # kernel runs over the candidate draws (8 B x P x C x N) and, unless it is diagnostics-only, over the
    # reference draws too (8 B x P x MREF).
    cand_only = dom in ("k_fold", "k_moments", "k_decide", "k_merge_partition")
    dom_alg_bytes = steps * 8 * P * (C * N + (0 if cand_only else MREF))
    achieved = dom_alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get(wname, {}).get(dom)
        except Exception:
            traffic = None
    pass_gbs = algorithmic_bytes(w) * world * steps / (total_ms * 1e-3) / 1e9
    pass_roofline = {"bound": "hbm", "achieved": round(pass_gbs, 1), "peak": peak * world, "unit": "GB/s",
                     "frac": round(pass_gbs / (peak * world), 4),
                     "note": "WHOLE STEP: algorithmic bytes (each draw read once) / step time -- the number to judge the pass by"}
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "launches": dom_n, "avg_launch_ms": round(dom_ms / max(dom_n, 1), 4),
                "share_of_kernel_time": round(dom_ms / tot_kernel_ms, 3),
                "note": "dominant kernel only (its share of the step is share_of_kernel_time); see pass_roofline for the step",
                "kernel_ms_per_step": {k: round(v[0] / steps, 3) for k, v in prof.items() if v[1]}}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(w, world), "compare_passed": passed,
        "clocks": clocks, "gpu_launches": int(nl.item()), "pass_roofline": pass_roofline, "roofline": roofline,
    }
    line["config"]["workspace_gb_cap"] = args.ws_gb  # caller-owned scratch: its size sets the number of parameter batches
    if not full or args.no_e2e:
        del cand, ref, ws
        torch.cuda.empty_cache()
        return line

    # ---- end to end from HOST buffers through the public call (H2D + D2H inside the timed region) ----
    # the pinned buffers are allocated after the process has been bound to its GPU's NUMA node (a no-op on a box
    # that exposes a single node); the affinity is restored for the CPU baseline's process pool below
    affinity0 = os.sched_getaffinity(0)
    host_topology = sharding.bind_host_to_gpu(local)
    cand_h = cand.cpu().pin_memory()
    ref_h = ref.cpu().pin_memory()
    sum_h = torch.empty((P, len(batch.SUMMARY_COLUMNS)), dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(steps, args.e2e_steps))

    def e2e_step():
        s, ok = batch.full_pass_host(cand_h, ref_h, tolerance=TOL, summary_host=sum_h, workspace=ws)
        if world > 1:
            ok = sharding.verdict_from_summary(sharding.gather_summaries(s.cuda(non_blocking=True), world * P))
        return s, ok

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s_host, ok_host = e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    h2d = 8 * P * (C * N + MREF) * world
    e2e_sec = float(e2e_s.item())
    e2e = {"value": dp_per_step * e2e_steps / e2e_sec, "unit": UNIT,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * len(batch.SUMMARY_COLUMNS) * P * world,
           "steps": e2e_steps, "h2d_gbs_aggregate": round(h2d * e2e_steps / e2e_sec / 1e9, 1),
           "limiter": "host->device copy: the pass is PCIe-bound end to end (h2d_gbs_aggregate is the whole box's "
                      "pinned-memory H2D rate over all ranks; the device-resident `value` is what the kernels sustain)",
           "host_topology": host_topology,
           "call": "mcmc_ref_b200.batch.full_pass_host -> b200ref_full_pass_host_f64 (pinned host buffers)"}
    same_bits = bool(np.array_equal(s_host.numpy(), summary.cpu().numpy(), equal_nan=True))
    del cand_h, ref_h
    os.sched_setaffinity(0, affinity0)

    # ---- CPU baseline + parity gate on a bounded sample (rank 0, N = 1 only) ----
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle.cpu_pool import OraclePool

        pool = OraclePool()
        ns = min(P, max(pool.cores, int(args.cpu_draws_per_core * pool.cores / (C * N))))
        idx = np.linspace(0, P - 1, ns).round().astype(int)
        pool.stage(cand[idx].cpu().numpy(), ref[idx].cpu().numpy())
        secs, want = pool.run(TOL)
        pool.close()
        cpu_baseline = {"value": ns * C * N / secs, "unit": UNIT, "cores": pool.cores, "kind": "port",
                        "sample": f"{ns} of {P} params (evenly spaced) of the same device-generated workload, "
                                  f"oracle/mcmc_oracle.py over a {pool.cores}-process pool, {secs:.1f} s"}
        got = summary.cpu().numpy()[idx]
        cols = batch.SUMMARY_COLUMNS
        max_rel = 0.0
        for k in ("rhat", "ess_bulk", "ess_tail", "mean", "std", "ref_mean", "ref_std"):
            g, wv = got[:, cols.index(k)], want[k]
            max_rel = max(max_rel, float(np.max(np.abs(g - wv) / np.maximum(np.abs(wv), 1e-300))))
        exact = all(np.array_equal(got[:, cols.index(k)], want[k]) for k in
                    ("q5", "q50", "q95", "ref_q5", "ref_q50", "ref_q95"))
        flags = got[:, cols.index("flags")].astype(int)
        verdicts = bool(np.array_equal((flags & 1) != 0, want["pass_mean"]) and
                        np.array_equal((flags & 2) != 0, want["pass_std"]))
        checks = bool(np.array_equal(got[:, 0] < 1.01, want["rhat"] < 1.01) and
                      np.array_equal(got[:, 1] > 400, want["ess_bulk"] > 400))
        parity = {"params_checked": int(ns), "max_rel_err_tolerance_class": max_rel, "quantiles_bit_exact": exact,
                  "compare_verdicts_identical": verdicts, "check_verdicts_identical": checks,
                  "ok": bool(max_rel < 1e-6 and exact and verdicts and checks)}
    line.update({"e2e": e2e, "cpu_baseline": cpu_baseline, "parity": parity, "e2e_bits_equal_device_path": same_bits})
    del cand, ref, ws
    torch.cuda.empty_cache()
    return line


def glm_measure(args, rank: int, world: int, local: int, full: bool) -> dict:
    """Config 3: logistic regression N=100k, D=256, batched log-density over 64k draws (tcgen05 path), the draws
    sharded in contiguous blocks over the ranks (X, y replicated once, untimed; one all_gather of lp per step)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from mcmc_ref_b200 import _native, batch, logdensity, sharding

    n, d, s = args.glm_n, args.glm_d, args.glm_s
    g = torch.Generator(device="cuda").manual_seed(4711)  # same seed on every rank: identical X, y, beta
    X = torch.randn((n, d), dtype=torch.float64, device="cuda", generator=g) / d ** 0.5
    beta_true = torch.randn(d, dtype=torch.float64, device="cuda", generator=g)
    y = (torch.rand(n, dtype=torch.float64, device="cuda", generator=g) < torch.sigmoid(X @ beta_true)).double()
    beta = beta_true[None, :] + 0.05 * torch.randn((s, d), dtype=torch.float64, device="cuda", generator=g)
    X, y = logdensity.broadcast_design(X, y)
    lo, hi = sharding.shard_bounds(s, world, rank)
    beta_local = beta[lo:hi].contiguous()
    ws = batch.Workspace()
    steps = args.steps if full else min(args.steps, 10)

    def step():
        # the design is the same device tensor every step: its staged bf16 split is reused (the first call stages it)
        return logdensity.glm_lp_sharded("bernoulli_logit", X, y, beta_local, n_draws=s, workspace=ws, reuse_design=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        lp = step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = _native.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        lp = step()
    e1.record()
    barrier()
    launches = _native.kernel_launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    mst = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    nl = torch.tensor([launches], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(mst, op=dist.ReduceOp.MAX)
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
    ms = float(mst.item()) / steps
    _native.profile(True)
    _native.profile_read(reset=True)
    barrier()
    for _ in range(steps):
        step()
    barrier()
    prof = _native.profile_read(reset=True)
    _native.profile(False)
    k_ms, k_n = prof["k_glm_lp"]
    flops_local = 2.0 * n * d * (hi - lo)
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = float(peaks.get("bf16_tflops", 1590.0))
    achieved = flops_local / (k_ms / max(k_n, 1) * 1e-3) / 1e12
    line = {
        "metric": "glm log-density draws/s", "value": s / (ms * 1e-3), "unit": "draws/s", "n_gpus": world,
        "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16x2-split in, fp32 accumulate, fp64 sum", "data": "synthetic",
        "config": {"workload": f"c3: logistic regression N={n}, D={d}, batched log-density over {s} draws",
                   "sharding": f"draw blocks x{world} (X, y replicated; all_gather of lp)",
                   "l2": f"operands {(n + s) * d * 4 / 1e6:.0f} MB of bf16 hi/lo re-streamed per tile, larger than L2"},
        "clocks": clocks, "gpu_launches": int(nl.item()),
        "roofline": {"bound": "tensor", "kernel": "k_glm_lp<bernoulli_logit>", "achieved": round(achieved, 1), "peak": peak,
                     "unit": "TFLOP/s", "frac": round(achieved / peak, 4), "traffic": None,
                     "note": "algorithmic flops 2*N*D*S_local per GPU; the kernel issues 3 bf16 MMAs per product for fp64-class accuracy",
                     "kernel_ms_per_step": {k: round(v[0] / steps, 3) for k, v in prof.items() if v[1]}},
    }
    # parity gate: fp64 numpy oracle on a sample of the draws (rank 0)
    if rank == 0:
        from oracle import glm_oracle

        ns = min(s, args.glm_cpu_draws if full else 128)
        idx = np.linspace(0, s - 1, ns).round().astype(int)
        Xn, yn, bn = X.cpu().numpy(), y.cpu().numpy(), beta.cpu().numpy()[idx]
        t0 = time.perf_counter()
        want = glm_oracle.logistic_lp(Xn, yn, bn)
        cpu_s = time.perf_counter() - t0
        got = lp.cpu().numpy()[idx]
        max_rel = float(np.max(np.abs(got - want) / np.abs(want)))
        line["cpu_baseline"] = {"value": ns / cpu_s, "unit": "draws/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{ns} of {s} draws, numpy fp64 oracle (BLAS threads), {cpu_s:.1f} s"}
        line["parity"] = {"draws_checked": int(ns), "max_rel_err": max_rel, "ok": bool(max_rel < 1e-6)}
    if full and world == 1:
        # end to end: pinned host fp64 buffers -> device -> kernel -> lp back on the host, every step
        Xh, yh, bh = X.cpu().pin_memory(), y.cpu().pin_memory(), beta.cpu().pin_memory()
        Xd, yd, bd = torch.empty_like(X), torch.empty_like(y), torch.empty_like(beta)
        lp_h = torch.empty(s, dtype=torch.float64).pin_memory()

        def e2e_step():
            Xd.copy_(Xh, non_blocking=True)
            yd.copy_(yh, non_blocking=True)
            bd.copy_(bh, non_blocking=True)
            lp_h.copy_(logdensity.logistic_lp(Xd, yd, bd, workspace=ws), non_blocking=True)
            torch.cuda.synchronize()

        e2e_step()
        e2e_steps = max(1, min(steps, args.e2e_steps))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        assert torch.equal(lp_h, lp.cpu())
        line["e2e"] = {"value": s / e2e_s, "unit": "draws/s", "h2d_bytes_per_step": 8 * (n * d + n + s * d),
                       "d2h_bytes_per_step": 8 * s, "steps": e2e_steps,
                       "call": "logdensity.logistic_lp on device copies of pinned host X, y, beta; lp copied back"}
    del X, y, beta, beta_local, ws
    torch.cuda.empty_cache()
    return line


def gpu_arm(args) -> None:
    import torch.distributed as dist

    rank, world, local = _dist_setup()
    if args.workload == "c3":
        line = glm_measure(args, rank, world, local, full=True)
    else:
        line = diag_measure(args, args.workload, rank, world, local, full=True)
        if args.workload == "c5tile" and not args.no_secondary:
            # the other named configurations, measured in the same run so that the driver's record holds them:
            # config 3 (tensor-core GLM, draw-block sharded over the same ranks) and, at N = 1, config 2
            secondary = {"c3": glm_measure(args, rank, world, local, full=False)}
            if world == 1:
                secondary["c2"] = diag_measure(args, "c2", rank, world, local, full=False)
            line["secondary"] = secondary
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def reference_arm(args) -> None:
    """The reference's own CPU implementation of the path on the host cores: the UNMODIFIED reference package
    (oracle/_ref, built by oracle/build_ref.py; "kind": "reference") when it is present, else the numpy port.
    Rank 0 only; other ranks exit without work.  Same `config`, steps and warm-up as the B200 arm; each step is a
    bounded sample of the workload (one parameter per core by default)."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    from oracle.build_ref import import_reference
    from oracle.cpu_pool import OraclePool, host_workload

    if args.workload == "c3":
        emit({"impl": "reference", "unavailable": "the reference has no log-density code (SURVEY.md 8c): config 3 is an extension"})
        return
    kind = args.ref_kind
    if kind == "auto":
        kind = "reference" if import_reference() is not None else "port"
    elif kind == "reference" and import_reference() is None:
        raise SystemExit("oracle/_ref is not built: run python oracle/build_ref.py where /root/reference exists")
    w = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", args.gpus))
    C, N, MREF = w["C"], w["N"], w["RC"] * w["RN"]
    pool = OraclePool()
    ns = min(w["P"], args.ref_params_per_core * pool.cores * (1 if kind == "reference" else 16))
    cand, ref = host_workload(ns, C, N, MREF)
    pool.stage(cand, ref)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        pool.run(TOL, impl=kind)
    t = 0.0
    for _ in range(args.steps):
        secs, out = pool.run(TOL, impl=kind)
        t += secs
    pool.close()
    value = ns * C * N * args.steps / t
    what = ("the unmodified reference package (oracle/_ref/mcmc_ref: split_rhat + ess_bulk + ess_tail + NumpyBackend.stats + "
            "compute_stats_from_draws + compare_stats per parameter)" if kind == "reference" else "oracle port (numpy)")
    sample = (f"each step = {ns} of {w['P']} params of the workload (numpy-generated, same recipe as the device generator), "
              f"{what} over a {pool.cores}-process pool")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(w, world), "compare_passed": bool(out["passed"]),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": pool.cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_RESULT_FD = None


def guard_stdout() -> None:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner when NCCL_DEBUG is set), so the real stdout is kept aside for the result line and file
    descriptor 1 is pointed at stderr for everything else (this process and its children)."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main() -> None:
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5tile", choices=sorted(WORKLOADS) + ["c3"])
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary c2 / c3 blocks of the default line")
    ap.add_argument("--glm-n", type=int, default=100_000)
    ap.add_argument("--glm-d", type=int, default=256)
    ap.add_argument("--glm-s", type=int, default=65_536)
    ap.add_argument("--glm-cpu-draws", type=int, default=512)
    ap.add_argument("--ws-gb", type=float, default=96.0,
                    help="cap of the device workspace (the config-5 shard then runs in 4 parameter batches; 32 GB: 9 batches, +3 ms)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-draws-per-core", type=float, default=1.0e7,
                    help="cpu_baseline leg: draws each host core gets (sample = that many parameters)")
    ap.add_argument("--ref-params-per-core", type=int, default=1)
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "reference", "port"],
                    help="--impl reference: the unmodified reference from oracle/_ref, or the numpy port")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="kernel-tuning runs: device-timed value and roofline blocks only")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
