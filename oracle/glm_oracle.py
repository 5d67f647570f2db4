"""fp64 numpy oracle of the GLM log-density extension.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED against the reference: mcmc-ref has no log-density code (grep for
log_dens|logp|log_prob hits nothing executable; ``lp__`` is stripped from every stored
draw, generate_draws.py:64-75).  The specification is the packaged Stan text, e.g.
``switched ~ bernoulli_logit(beta[1] + beta[2] * dist)`` (stan_models/wells_data-wells_dist.stan:9-11):
the likelihood block contributes  sum_n y_n*eta_n - log(1 + exp(eta_n)),  eta = X beta.
Priors and the dropped-constant convention of ``~`` are outside this function.
"""

from __future__ import annotations

import numpy as np


def logistic_lp(X: np.ndarray, y: np.ndarray, beta: np.ndarray, block: int = 256) -> np.ndarray:
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    beta = np.asarray(beta, dtype=np.float64)
    out = np.empty(beta.shape[0])
    for s0 in range(0, beta.shape[0], block):
        eta = X @ beta[s0:s0 + block].T                      # [N, b]
        out[s0:s0 + block] = (y[:, None] * eta - np.logaddexp(0.0, eta)).sum(axis=0)
    return out
