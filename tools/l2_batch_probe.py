"""Per-kernel time per parameter as a function of the number of parameters processed per batch (config-5 shape):
does a batch whose workspace stays in L2 run the scatter-bound kernels faster?  Prints ms per 1000 parameters."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "mcmc-db_b200")]
import torch
from mcmc_ref_b200 import _native, batch

C, N = 64, 4096
for P in [int(a) for a in sys.argv[1:]] or [4, 8, 16, 32, 64, 256, 1389]:
    cand = batch.synth_fill(P, C, N, kind=0)
    ref = batch.synth_fill(P, 10, 1000, kind=1).reshape(P, 10000)
    ws = batch.Workspace()
    reps = max(3, 512 // P)
    for _ in range(3):
        batch.full_pass(cand, ref, tolerance=0.15, workspace=ws)
    torch.cuda.synchronize()
    _native.profile(True)
    _native.profile_read(reset=True)
    for _ in range(reps):
        batch.full_pass(cand, ref, tolerance=0.15, workspace=ws)
    prof = _native.profile_read(reset=True)
    _native.profile(False)
    sc = 1000.0 / (P * reps)
    print(P, {k: round(v[0] * sc, 3) for k, v in prof.items() if v[0] * sc >= 0.02}, "total", round(sum(v[0] for v in prof.values()) * sc, 2), flush=True)
    del cand, ref, ws
    torch.cuda.empty_cache()
