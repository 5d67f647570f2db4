"""Batched log-density of GLM-style posteriordb models over draws (extension: the reference
packages the Stan text and data of these models but never evaluates them).

``logistic_lp(X, y, beta)`` evaluates, for every draw ``beta[s]``,
``sum_n y[n]*eta - log1p(exp(eta))`` with ``eta = X @ beta[s]`` -- the likelihood block of
``y ~ bernoulli_logit(X * beta)`` -- on the tcgen05 tensor cores (csrc/glm.cuh).
"""

from __future__ import annotations

import torch

from . import _native, batch

_WS = batch.Workspace()


def logistic_lp(X, y, beta, workspace: batch.Workspace | None = None) -> torch.Tensor:
    """lp[S] (float64, CUDA) for X [N, D], y [N] in {0, 1}, beta [S, D]."""
    lib = _native.load()
    xt, yt, bt = batch.to_device(X), batch.to_device(y).reshape(-1), batch.to_device(beta)
    n, d = xt.shape
    s = bt.shape[0]
    if bt.shape[1] != d or yt.numel() != n:
        raise ValueError(f"shape mismatch: X {tuple(xt.shape)}, y {tuple(yt.shape)}, beta {tuple(bt.shape)}")
    lp = torch.empty(s, dtype=torch.float64, device=xt.device)
    ws = (workspace or _WS).ensure(int(lib.b200ref_glm_workspace_bytes(n, d, s)))
    _native.check(lib.b200ref_glm_logistic_lp_f64(xt.data_ptr(), yt.data_ptr(), bt.data_ptr(), n, d, s,
                                                  lp.data_ptr(), ws.data_ptr(), ws.numel(),
                                                  batch._stream_ptr()))
    return lp


def glm_lp(family: str, X, y, beta, trials=None, sigma=None, workspace: batch.Workspace | None = None,
           reuse_design: bool = False) -> torch.Tensor:
    """lp[S] of the likelihood block of a packaged GLM-style model, eta = X @ beta[s] (Stan ``~`` semantics:
    parameter-free terms dropped):

    * ``"bernoulli_logit"``  ``y ~ bernoulli_logit(X * beta)``          (wells_data-wells_dist.stan:9-11)
    * ``"binomial_logit"``   ``y ~ binomial_logit(trials, X * beta)``   (provenance.py:567-606)
    * ``"poisson_log"``      ``y ~ poisson_log(X * beta)``              (provenance.py:609-640)
    * ``"normal"``           ``y ~ normal(X * beta, sigma)``, ``sigma`` [S] per draw (sblri-blr.stan:12-17)

    ``reuse_design=True``: the caller states that ``X`` / ``y`` / ``trials`` are the SAME device tensors as in the
    previous call with this workspace (new draws against the same design); the staged bf16 split of X is reused.
    """
    if family not in _native.GLM_FAMILIES:
        raise ValueError(f"family must be one of {sorted(_native.GLM_FAMILIES)}; got {family!r}")
    lib = _native.load()
    xt, yt, bt = batch.to_device(X), batch.to_device(y).reshape(-1), batch.to_device(beta)
    n, d = xt.shape
    s = bt.shape[0]
    if bt.shape[1] != d or yt.numel() != n:
        raise ValueError(f"shape mismatch: X {tuple(xt.shape)}, y {tuple(yt.shape)}, beta {tuple(bt.shape)}")
    tt = st = None
    if family == "binomial_logit":
        if trials is None:
            raise ValueError("binomial_logit needs trials [N]")
        tt = batch.to_device(trials).reshape(-1)
        if tt.numel() != n:
            raise ValueError(f"shape mismatch: trials {tuple(tt.shape)} for N = {n}")
    if family == "normal":
        if sigma is None:
            raise ValueError("the normal family needs sigma [S] (one per draw)")
        st = batch.to_device(sigma).reshape(-1)
        if st.numel() != s:
            raise ValueError(f"shape mismatch: sigma {tuple(st.shape)} for S = {s}")
    lp = torch.empty(s, dtype=torch.float64, device=xt.device)
    ws = (workspace or _WS).ensure(int(lib.b200ref_glm_workspace_bytes(n, d, s)))
    flags = _native.GLM_REUSE_DESIGN if reuse_design else 0
    _native.check(lib.b200ref_glm_lp_ex_f64(_native.GLM_FAMILIES[family], xt.data_ptr(), yt.data_ptr(),
                                            tt.data_ptr() if tt is not None else None, bt.data_ptr(),
                                            st.data_ptr() if st is not None else None, n, d, s, flags, lp.data_ptr(),
                                            ws.data_ptr(), ws.numel(), batch._stream_ptr()))
    return lp


def glm_lp_sharded(family: str, X, y, beta_local, n_draws: int | None = None, trials=None, sigma_local=None,
                   group=None, workspace: batch.Workspace | None = None, reuse_design: bool = False) -> torch.Tensor:
    """Draw-block sharded ``glm_lp``: every rank holds the design ``X`` / ``y`` (replicated once,
    see ``broadcast_design``) and ITS contiguous block of the draws (``sharding.shard_bounds(n_draws, world, rank)``;
    ``sigma_local`` likewise for the normal family); each evaluates its block on its own GPU and one all_gather
    of the S log-densities gives every rank ``lp[S]`` in draw order.  Single process: same as ``glm_lp``."""
    from . import sharding

    lp_local = glm_lp(family, X, y, beta_local, trials=trials, sigma=sigma_local, workspace=workspace,
                      reuse_design=reuse_design)
    return sharding.gather_draw_blocks(lp_local, n_draws, group)


def broadcast_design(X, y, src: int = 0, group=None) -> tuple[torch.Tensor, torch.Tensor]:
    """Replicate the design matrix and response from rank ``src`` to every rank (untimed setup; 205 MB for config 3).
    Every rank passes tensors of the right shape; the contents of the non-source ranks are overwritten."""
    import torch.distributed as dist

    xt, yt = batch.to_device(X), batch.to_device(y).reshape(-1)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(xt, src=src, group=group)
        dist.broadcast(yt, src=src, group=group)
    return xt, yt
