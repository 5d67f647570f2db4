#!/usr/bin/env python
"""Benchmark of the diagnostics + compare pass (BASELINE.json metric: draw*params/s).

  python bench.py --gpus N --steps K --warmup W            # B200 arm (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the reference algorithm on host cores

A step = one full pass over the workload's parameter shard: sort -> tie ranks -> normal
scores (raw + folded) -> split R-hat, bulk/tail ESS, mean/std/quantiles of the candidate
draws; mean/std/quantiles of the reference draws; compare verdicts (tol 0.15); for N > 1
the NCCL gather of the [P, 20] summaries and the AND of the verdicts.

Workloads (per GPU, weak scaling: parameters are sharded by column):
  c2      1000 params x 16 chains x 10000 draws + 10x1000 reference draws  (BASELINE configs[1])
  c5tile  12500 params x 64 chains x 4096 draws (1/8 of configs[4]; 8 GPUs = the full config)
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


def algorithmic_bytes(w: dict) -> int:
    """every fp64 draw read once + the summary row (20 doubles) per parameter."""
    return 8 * w["P"] * (w["C"] * w["N"] + w["RC"] * w["RN"]) + 8 * 20 * w["P"]


def gpu_arm(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mcmc_ref_b200 import _native, batch, sharding

    w = WORKLOADS[args.workload]
    P, C, N, MREF = w["P"], w["C"], w["N"], w["RC"] * w["RN"]
    cand = batch.synth_fill(P, C, N, p0=rank * P, kind=0)
    ref = batch.synth_fill(P, w["RC"], w["RN"], p0=rank * P, kind=1).reshape(P, MREF)
    ws = batch.Workspace()
    ws_cap = int(args.ws_gb * 2**30)

    def step():
        # this rank's parameter block, then the [P/G, 20] summary gather + verdict AND (NCCL)
        summary, passed = batch.full_pass(cand, ref, tolerance=TOL, workspace=ws, ws_cap_bytes=ws_cap)
        if world > 1:
            sharding.gather_summaries(summary, world * P)
            passed = sharding.all_passed(passed, summary.device)
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
    for _ in range(args.steps):
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
    total_ms = float(ms.item())
    dp_per_step = world * P * C * N
    value = dp_per_step * args.steps / (total_ms * 1e-3)

    # ---- per-kernel device time over the same steps (events around every launch) ----
    _native.profile(True)
    _native.profile_read(reset=True)
    barrier()
    for _ in range(args.steps):
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
    # algorithmic bytes of the launches of the dominant kernel: 8 B x draws each launch processes.
    # All our heavy kernels are launched once per (batch, stage) over candidate and reference draws.
    # per step, a kernel is launched `sides` times over the candidate draws (8 B x P x C x N each)
    # and, unless it is diagnostics-only, as often over the reference draws (8 B x P x MREF each).
    cand_only = dom in ("k_fold", "k_moments", "k_decide", "k_merge_partition")
    dom_alg_bytes = args.steps * 8 * P * (C * N + (0 if cand_only else MREF))
    achieved = dom_alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get(args.workload, {}).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "launches": dom_n, "avg_launch_ms": round(dom_ms / max(dom_n, 1), 4),
                "share_of_kernel_time": round(dom_ms / tot_kernel_ms, 3),
                "kernel_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in prof.items() if v[1]}}
    pass_gbs = algorithmic_bytes(w) * world * args.steps / (total_ms * 1e-3) / 1e9
    pass_roofline = {"bound": "hbm", "achieved": round(pass_gbs, 1), "peak": peak * world, "unit": "GB/s",
                     "frac": round(pass_gbs / (peak * world), 4),
                     "note": "whole step: algorithmic bytes (each draw read once) / step time"}

    # ---- end to end from HOST buffers through the public call (H2D + D2H inside the timed region) ----
    cand_h = cand.cpu().pin_memory()
    ref_h = ref.cpu().pin_memory()
    sum_h = torch.empty((P, len(batch.SUMMARY_COLUMNS)), dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def e2e_step():
        s, ok = batch.full_pass_host(cand_h, ref_h, tolerance=TOL, summary_host=sum_h, workspace=ws)
        if world > 1:
            sharding.gather_summaries(s.cuda(non_blocking=True), world * P)
            ok = sharding.all_passed(ok, torch.device("cuda", local))
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
    e2e = {"value": dp_per_step * e2e_steps / float(e2e_s.item()), "unit": UNIT,
           "h2d_bytes_per_step": 8 * P * (C * N + MREF) * world, "d2h_bytes_per_step": 8 * len(batch.SUMMARY_COLUMNS) * P * world,
           "steps": e2e_steps, "call": "mcmc_ref_b200.batch.full_pass_host -> b200ref_full_pass_host_f64 (pinned host buffers)"}
    same_bits = bool(np.array_equal(s_host.numpy(), summary.cpu().numpy(), equal_nan=True))

    # ---- CPU baseline + parity gate on a bounded sample (rank 0, N = 1 only) ----
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle.cpu_pool import OraclePool

        pool = OraclePool()
        ns = min(P, args.cpu_params_per_core * pool.cores)
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

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "params_per_gpu": P, "chains": C, "draws": N, "ref_draws": MREF,
                       "tolerance": TOL, "sharding": f"param columns x{world}",
                       "l2": f"inputs larger than L2 ({algorithmic_bytes(w) / 1e9:.2f} GB read per step per GPU)",
                       "compare_passed": bool(passed)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(nl.item()), "roofline": roofline,
            "pass_roofline": pass_roofline, "cpu_baseline": cpu_baseline, "parity": parity,
            "e2e_bits_equal_device_path": same_bits,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def glm_arm(args) -> None:
    """Config 3: logistic regression N=100k, D=256, batched log-density over 64k draws (tcgen05 path)."""
    import numpy as np
    import torch

    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    from mcmc_ref_b200 import _native, batch, logdensity
    from oracle import glm_oracle

    n, d, s = args.glm_n, args.glm_d, args.glm_s
    g = torch.Generator(device="cuda").manual_seed(4711)
    X = torch.randn((n, d), dtype=torch.float64, device="cuda", generator=g) / d ** 0.5
    beta_true = torch.randn(d, dtype=torch.float64, device="cuda", generator=g)
    y = (torch.rand(n, dtype=torch.float64, device="cuda", generator=g) < torch.sigmoid(X @ beta_true)).double()
    beta = beta_true[None, :] + 0.05 * torch.randn((s, d), dtype=torch.float64, device="cuda", generator=g)
    ws = batch.Workspace()
    for _ in range(max(args.warmup, 3)):
        lp = logdensity.logistic_lp(X, y, beta, workspace=ws)
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    l0 = _native.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        lp = logdensity.logistic_lp(X, y, beta, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    launches = _native.kernel_launches() - l0
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    _native.profile(True)
    _native.profile_read(reset=True)
    for _ in range(args.steps):
        logdensity.logistic_lp(X, y, beta, workspace=ws)
    prof = _native.profile_read(reset=True)
    _native.profile(False)
    k_ms, k_n = prof["k_glm_lp"]
    flops = 2.0 * n * d * s
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = float(peaks.get("bf16_tflops", 1590.0))
    achieved = flops / (k_ms / max(k_n, 1) * 1e-3) / 1e12
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
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert torch.equal(lp_h, lp.cpu())
    # CPU baseline: fp64 numpy oracle (BLAS threads) on a sample of the draws + parity gate
    ns = min(s, args.glm_cpu_draws)
    idx = np.linspace(0, s - 1, ns).round().astype(int)
    Xn, yn, bn = X.cpu().numpy(), y.cpu().numpy(), beta.cpu().numpy()[idx]
    t0 = time.perf_counter()
    want = glm_oracle.logistic_lp(Xn, yn, bn)
    cpu_s = time.perf_counter() - t0
    got = lp.cpu().numpy()[idx]
    max_rel = float(np.max(np.abs(got - want) / np.abs(want)))
    line = {
        "metric": "glm log-density draws/s", "value": s / (ms * 1e-3), "unit": "draws/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16x2-split in, fp32 accumulate, fp64 sum", "data": "synthetic",
        "config": {"workload": f"c3: logistic regression N={n}, D={d}, batched log-density over {s} draws",
                   "l2": f"operands {(n + s) * d * 4 / 1e6:.0f} MB of bf16 hi/lo re-streamed per tile, larger than L2"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": s / e2e_s, "unit": "draws/s", "h2d_bytes_per_step": 8 * (n * d + n + s * d),
                "d2h_bytes_per_step": 8 * s, "steps": e2e_steps,
                "call": "logdensity.logistic_lp on device copies of pinned host X, y, beta; lp copied back"},
        "roofline": {"bound": "tensor", "kernel": "k_glm_lp<bernoulli_logit>", "achieved": round(achieved, 1), "peak": peak,
                     "unit": "TFLOP/s", "frac": round(achieved / peak, 4), "traffic": None,
                     "note": "algorithmic flops 2*N*D*S; the kernel issues 3 bf16 MMAs per product for fp64-class accuracy",
                     "kernel_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in prof.items() if v[1]}},
        "cpu_baseline": {"value": ns / cpu_s, "unit": "draws/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{ns} of {s} draws, numpy fp64 oracle (BLAS threads), {cpu_s:.1f} s"},
        "parity": {"draws_checked": int(ns), "max_rel_err": max_rel, "ok": bool(max_rel < 1e-6)},
    }
    emit(line)


def reference_arm(args) -> None:
    """The reference algorithm on the host cores (oracle port; the Python reference itself
    cannot travel to the GPU box).  Rank 0 only; other ranks exit without work."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    from oracle.cpu_pool import OraclePool, host_workload

    w = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", args.gpus))
    C, N, MREF = w["C"], w["N"], w["RC"] * w["RN"]
    pool = OraclePool()
    ns = min(w["P"], args.ref_params_per_core * pool.cores)
    cand, ref = host_workload(ns, C, N, MREF)
    pool.stage(cand, ref)
    for _ in range(max(1, min(args.warmup, 3))):
        pool.run(TOL)
    t = 0.0
    for _ in range(args.steps):
        secs, out = pool.run(TOL)
        t += secs
    pool.close()
    value = ns * C * N * args.steps / t
    sample = (f"each step = {ns} of {w['P']} params of the workload (numpy-generated, same recipe), "
              f"oracle port over a {pool.cores}-process pool")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "params_per_gpu": w["P"], "chains": C, "draws": N, "ref_draws": MREF,
                   "tolerance": TOL, "sharding": "none (CPU)", "compare_passed": bool(out["passed"])},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": pool.cores, "kind": "port", "sample": sample},
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
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c3"])
    ap.add_argument("--glm-n", type=int, default=100_000)
    ap.add_argument("--glm-d", type=int, default=256)
    ap.add_argument("--glm-s", type=int, default=65_536)
    ap.add_argument("--glm-cpu-draws", type=int, default=512)
    ap.add_argument("--ws-gb", type=float, default=16.0, help="cap of the device workspace")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-params-per-core", type=int, default=64)
    ap.add_argument("--ref-params-per-core", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.workload == "c3":
        glm_arm(args)
    elif args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
