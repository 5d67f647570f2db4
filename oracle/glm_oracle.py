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


def glm_lp(family: str, X, y, beta, trials=None, sigma=None, block: int = 256) -> np.ndarray:
    """Likelihood block of the packaged GLM families in fp64 (Stan ``~`` semantics: terms without
    parameters are dropped).  Specs: ``provenance.py:567-606`` (binomial_logit), ``:609-640``
    (poisson_log), ``stan_models/sblri-blr.stan:12-17`` (normal, sigma per draw)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    beta = np.asarray(beta, dtype=np.float64)
    out = np.empty(beta.shape[0])
    for s0 in range(0, beta.shape[0], block):
        eta = X @ beta[s0:s0 + block].T                      # [N, b]
        if family == "bernoulli_logit":
            t = y[:, None] * eta - np.logaddexp(0.0, eta)
        elif family == "binomial_logit":
            t = y[:, None] * eta - np.asarray(trials, dtype=np.float64).reshape(-1, 1) * np.logaddexp(0.0, eta)
        elif family == "poisson_log":
            t = y[:, None] * eta - np.exp(eta)
        elif family == "normal":
            sg = np.asarray(sigma, dtype=np.float64)[s0:s0 + block]
            out[s0:s0 + block] = -X.shape[0] * np.log(sg) - ((y[:, None] - eta) ** 2).sum(axis=0) / (2.0 * sg * sg)
            continue
        else:
            raise ValueError(family)
        out[s0:s0 + block] = t.sum(axis=0)
    return out


def glm_term_scale(family: str, X, y, beta, trials=None, sigma=None, block: int = 256) -> np.ndarray:
    """sum_n of the ABSOLUTE values of the two terms each observation contributes: the magnitude the
    1e-6 tolerance is relative to (lp itself can be a small difference of large sums, e.g. Poisson
    counts near their means)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    beta = np.asarray(beta, dtype=np.float64)
    out = np.empty(beta.shape[0])
    for s0 in range(0, beta.shape[0], block):
        eta = X @ beta[s0:s0 + block].T
        if family == "bernoulli_logit":
            t = np.abs(y[:, None] * eta) + np.logaddexp(0.0, eta)
        elif family == "binomial_logit":
            t = np.abs(y[:, None] * eta) + np.asarray(trials, dtype=np.float64).reshape(-1, 1) * np.logaddexp(0.0, eta)
        elif family == "poisson_log":
            t = np.abs(y[:, None] * eta) + np.exp(eta)
        else:
            sg = np.asarray(sigma, dtype=np.float64)[s0:s0 + block]
            t = (y[:, None] - eta) ** 2 / (2.0 * sg * sg) + np.abs(np.log(sg))
        out[s0:s0 + block] = t.sum(axis=0)
    return out
