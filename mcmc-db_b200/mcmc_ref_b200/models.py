"""Registry of the packaged posteriordb-style models whose likelihood is a GLM contraction, and
their batched log-density over draws (EXTENSION, parity unpinned: mcmc-ref ships the Stan text and
the data of these models -- ``reference.stan_data`` / ``reference.model_code``, reference.py:20-27 --
but never evaluates them; the Stan text is the specification, cited per model below).

Every entry restates the model block's likelihood statement as

    response ~ family(design @ coef [, sigma | trials])

where ``response`` and the design columns are expressions over the model's ``stan_data`` (the
``transformed data`` block of the Stan file, written out), and ``coef`` / ``sigma`` name columns of
the packaged draws table.  ``packaged_lp`` builds X [N, D] and beta [S, D] and runs the tcgen05
kernel behind ``logdensity.glm_lp`` (csrc/glm.cuh); priors, Jacobians and the constants Stan's ``~``
drops are outside this function, as in ``logdensity.glm_lp``.
"""

from __future__ import annotations

from collections.abc import Mapping, Sequence
from dataclasses import dataclass

import numpy as np

from .store import PackagedStore

_STAN = "packages/mcmc-ref-data/src/mcmc_ref_data/data/stan_models"


@dataclass(frozen=True)
class GlmModel:
    family: str                    # a key of _native.GLM_FAMILIES
    response: str                  # expression over stan_data
    columns: tuple[str, ...]       # design-matrix columns: expressions over stan_data; "@M" = every column of matrix M
    coef: tuple[str, ...]          # draw columns, one per design column; "v[*]" / "v[*,2]" = index 1 .. k over an "@M" block
    sigma: str | None = None       # draw column (normal family); "=1" = the constant 1 (scale folded into the design)
    trials: str | None = None      # expression over stan_data (binomial_logit)
    source: str = ""               # Stan file (under stan_models/) and the lines restated


def _beta(k: int) -> tuple[str, ...]:
    return tuple(f"beta[{i}]" for i in range(1, k + 1))


def _lin(response: str, *columns: str, source: str, sigma: str = "sigma", coef: Sequence[str] | None = None) -> GlmModel:
    cols = ("1", *columns)
    return GlmModel("normal", response, cols, tuple(coef) if coef is not None else _beta(len(cols)), sigma=sigma, source=source)


_Z_HEIGHT = "(height - mean(height)) / sd(height)"
_LOGVOL = "log(diam1 * diam2 * canopy_height)"
_LOGAREA = "log(diam1 * diam2)"
_LOGSHAPE = "log(diam1 / diam2)"
_C_HS, _C_IQ = "(mom_hs - mean(mom_hs))", "(mom_iq - mean(mom_iq))"
_C2_HS, _C2_IQ = "(mom_hs - 0.5)", "(mom_iq - 100)"
_Z_HS, _Z_IQ = "((mom_hs - mean(mom_hs)) / (2 * sd(mom_hs)))", "((mom_iq - mean(mom_iq)) / (2 * sd(mom_iq)))"
_NES = ("real_ideo", "race_adj", "(age_discrete == 2)", "(age_discrete == 3)", "(age_discrete == 4)", "educ1",
        "gender", "income")

MODELS: dict[str, GlmModel] = {
    # ---- earnings (stan_models/earnings-*.stan, model blocks) ----
    "earnings-earn_height": _lin("earn", "height", source="earnings-earn_height.stan:10-12"),
    "earnings-log10earn_height": _lin("log10(earn)", "height", source="earnings-log10earn_height.stan:6-19"),
    "earnings-logearn_height": _lin("log(earn)", "height", source="earnings-logearn_height.stan:6-17"),
    "earnings-logearn_height_male": _lin("log(earn)", "height", "male", source="earnings-logearn_height_male.stan:7-18"),
    "earnings-logearn_interaction": _lin("log(earn)", "height", "male", "height * male",
                                         source="earnings-logearn_interaction.stan:7-20"),
    "earnings-logearn_interaction_z": _lin("log(earn)", _Z_HEIGHT, "male", f"({_Z_HEIGHT}) * male",
                                           source="earnings-logearn_interaction_z.stan:7-22"),
    "earnings-logearn_logheight_male": _lin("log(earn)", "log(height)", "male",
                                            source="earnings-logearn_logheight_male.stan:7-20"),
    # ---- kidiq ----
    "kidiq-kidscore_momhs": _lin("kid_score", "mom_hs", source="kidiq-kidscore_momhs.stan:10-13"),
    "kidiq-kidscore_momiq": _lin("kid_score", "mom_iq", source="kidiq-kidscore_momiq.stan:10-13"),
    "kidiq-kidscore_momhsiq": _lin("kid_score", "mom_hs", "mom_iq", source="kidiq-kidscore_momhsiq.stan:11-14"),
    "kidiq-kidscore_interaction": _lin("kid_score", "mom_hs", "mom_iq", "mom_hs * mom_iq",
                                       source="kidiq-kidscore_interaction.stan:7-20"),
    "kidiq_with_mom_work-kidscore_interaction_c": _lin("kid_score", _C_HS, _C_IQ, f"{_C_HS} * {_C_IQ}",
                                                       source="kidiq_with_mom_work-kidscore_interaction_c.stan:7-23"),
    "kidiq_with_mom_work-kidscore_interaction_c2": _lin("kid_score", _C2_HS, _C2_IQ, f"{_C2_HS} * {_C2_IQ}",
                                                        source="kidiq_with_mom_work-kidscore_interaction_c2.stan:7-23"),
    "kidiq_with_mom_work-kidscore_interaction_z": _lin("kid_score", _Z_HS, _Z_IQ, f"{_Z_HS} * {_Z_IQ}",
                                                       source="kidiq_with_mom_work-kidscore_interaction_z.stan:7-23"),
    "kidiq_with_mom_work-kidscore_mom_work": _lin("kid_score", "(mom_work == 2)", "(mom_work == 3)", "(mom_work == 4)",
                                                  source="kidiq_with_mom_work-kidscore_mom_work.stan:6-23"),
    # ---- mesquite ----
    "mesquite-mesquite": _lin("weight", "diam1", "diam2", "canopy_height", "total_height", "density", "group",
                              source="mesquite-mesquite.stan:15-19"),
    "mesquite-logmesquite": _lin("log(weight)", "log(diam1)", "log(diam2)", "log(canopy_height)", "log(total_height)",
                                 "log(density)", "group", source="mesquite-logmesquite.stan:11-35"),
    "mesquite-logmesquite_logva": _lin("log(weight)", _LOGVOL, _LOGAREA, "group",
                                       source="mesquite-logmesquite_logva.stan:9-24"),
    "mesquite-logmesquite_logvas": _lin("log(weight)", _LOGVOL, _LOGAREA, _LOGSHAPE, "log(total_height)", "log(density)",
                                        "group", source="mesquite-logmesquite_logvas.stan:11-35"),
    "mesquite-logmesquite_logvash": _lin("log(weight)", _LOGVOL, _LOGAREA, _LOGSHAPE, "log(total_height)", "group",
                                         source="mesquite-logmesquite_logvash.stan:10-31"),
    "mesquite-logmesquite_logvolume": _lin("log(weight)", _LOGVOL, source="mesquite-logmesquite_logvolume.stan:8-20"),
    "mesquite_logvolume_informed": _lin("log_weight_std", "log_canopy_volume_std", source="mesquite_logvolume_informed.stan:15-22"),
    # ---- national election study, one model per year ----
    **{f"nes{year}-nes": _lin("partyid7", *_NES, source=f"nes{year}-nes.stan:11-32")
       for year in (1972, 1976, 1980, 1984, 1988, 1992, 1996, 2000)},
    # ---- design matrix given as data ----
    "sblrc-blr": GlmModel("normal", "y", ("@X",), ("beta[*]",), sigma="sigma", source="sblrc-blr.stan:11-17"),
    "sblri-blr": GlmModel("normal", "y", ("@X",), ("beta[*]",), sigma="sigma", source="sblri-blr.stan:11-17"),
    "blr_informed": GlmModel("normal", "y_std", ("@X_std",), ("beta[*]",), sigma="sigma", source="blr_informed.stan:15-22"),
    # brms-generated: columns 2..K of X centred, population-level effects b, separate Intercept
    "diamonds-diamonds": GlmModel("normal", "Y", ("1", "@center(X[:, 1:])"), ("Intercept", "b[*]"), sigma="sigma",
                                  source="diamonds-diamonds.stan:13-41"),
    # ---- scalar-named coefficients ----
    "kidscore_momiq_informed": _lin("kid_score_std", "mom_iq_std", coef=("beta_0", "beta_1"),
                                    source="kidscore_momiq_informed.stan:15-23"),
    "logearn_height_informed": _lin("log_earn_std", "height_std", coef=("beta_0", "beta_1"),
                                    source="logearn_height_informed.stan:15-23"),
    "radon_pooled": _lin("log_radon", "floor_measure", coef=("beta_0", "beta_1"), source="radon_pooled.stan:14-21"),
    "radon_pooled_informed": _lin("log_radon_std", "floor_measure_std", coef=("alpha", "beta"), sigma="sigma_y",
                                  source="radon_pooled_informed.stan:15-23"),
    "kilpisjarvi_mod-kilpisjarvi": _lin("y", "x", coef=("alpha", "beta"), source="kilpisjarvi_mod-kilpisjarvi.stan:17-21"),
    # ---- hierarchical non-centred models: the packaged draws carry the transformed parameters (alpha[j],
    # theta[j], ...), so the likelihood block is a GLM over one-hot group designs ----
    "eight_schools-eight_schools_noncentered": GlmModel("normal", "y / sigma", ("@diag(1 / sigma)",), ("theta[*]",), sigma="=1",
                                                        source="eight_schools-eight_schools_noncentered.stan:11-21"),
    "eight_schools-noncentered": GlmModel("normal", "y / sigma", ("@diag(1 / sigma)",), ("theta[*]",), sigma="=1",
                                          source="eight_schools-noncentered.stan:12-20"),
    "hierarchical_lr-noncentered": GlmModel("normal", "y", ("@onehot(group, J)", "x"), ("alpha[*]", "beta"), sigma="sigma",
                                            source="hierarchical_lr-noncentered.stan:16-28"),
    "radon_mn-radon_hierarchical_intercept_noncentered": GlmModel(
        "normal", "log_radon", ("@onehot(county_idx, J)", "log_uppm", "floor_measure"), ("alpha[*]", "beta[1]", "beta[2]"),
        sigma="sigma_y", source="radon_mn-radon_hierarchical_intercept_noncentered.stan:16-36"),
    "varying_slopes-noncentered": GlmModel("normal", "y", ("@onehot(group, J)", "@onehot(group, J) * x[:, None]"),
                                           ("beta_group[*,1]", "beta_group[*,2]"), sigma="sigma",
                                           source="varying_slopes-noncentered.stan:15-30"),
    "bangladesh_contraceptive-noncentered": GlmModel("bernoulli_logit", "use",
                                                     ("@onehot(district, D)", "@onehot(district, D) * urban[:, None]"),
                                                     ("a[*]", "b[*]"), source="bangladesh_contraceptive-noncentered.stan:17-31"),
    # ---- autoregression: y[t] ~ normal(alpha + sum_k beta[k] * y[t - k], sigma), t = K + 1 .. T ----
    "arK-arK": GlmModel("normal", "y[int(K):]", ("1", "@lagged(y, K)"), ("alpha", "beta[*]"), sigma="sigma",
                        source="arK-arK.stan:11-26"),
    # ---- non-Gaussian GLMs ----
    "wells_data-wells_dist": GlmModel("bernoulli_logit", "switched", ("1", "dist"), _beta(2),
                                      source="wells_data-wells_dist.stan:9-11"),
    "GLM_Poisson_Data-GLM_Poisson_model": GlmModel("poisson_log", "C", ("1", "year", "year ** 2", "year ** 3"),
                                                   ("alpha", "beta1", "beta2", "beta3"),
                                                   source="GLM_Poisson_Data-GLM_Poisson_model.stan:6-30"),
    "GLM_Binomial_data-GLM_Binomial_model": GlmModel("binomial_logit", "C", ("1", "year", "year * year"),
                                                     ("alpha", "beta1", "beta2"), trials="N",
                                                     source="GLM_Binomial_data-GLM_Binomial_model.stan:7-32"),
}


def supported_models() -> list[str]:
    return sorted(MODELS)


def _env(data: Mapping[str, object]) -> dict[str, object]:
    env: dict[str, object] = {
        "log": np.log, "log10": np.log10, "exp": np.exp, "sqrt": np.sqrt, "square": np.square,
        "mean": lambda v: float(np.mean(v)),
        "sd": lambda v: float(np.std(v, ddof=1)),          # Stan's sd() is the sample standard deviation
        "center": lambda m: m - np.mean(m, axis=0),        # brms' Xc
        "diag": np.diag,
        "int": int,
        # column k (1-based) = y[t - k] for t = K + 1 .. T
        "lagged": lambda y, k: np.stack([y[int(k) - j:len(y) - j] for j in range(1, int(k) + 1)], axis=1),
        "onehot": lambda idx, k: (np.asarray(idx, dtype=np.int64)[:, None] == np.arange(1, int(k) + 1)[None, :]).astype(np.float64),
    }
    for key, value in data.items():
        env[key] = np.asarray(value, dtype=np.float64) if isinstance(value, (list, tuple)) else value
    return env


def _eval(expr: str, env: Mapping[str, object]):
    return eval(expr, {"__builtins__": {}}, dict(env))  # expressions come from MODELS above, never from input


def design(model: str, data: Mapping[str, object]) -> tuple[np.ndarray, np.ndarray, np.ndarray | None, list[int]]:
    """(X [N, D], y [N], trials [N] or None, width of every design block) of a registered model."""
    spec = MODELS[model]
    env = _env(data)
    y = np.asarray(_eval(spec.response, env), dtype=np.float64).reshape(-1)
    n = y.shape[0]
    blocks, widths = [], []
    for col in spec.columns:
        if col == "1":
            block = np.ones((n, 1))
        elif col.startswith("@"):
            block = np.asarray(_eval(col[1:], env), dtype=np.float64).reshape(n, -1)
        else:
            block = np.asarray(_eval(col, env), dtype=np.float64).reshape(n, 1)
        blocks.append(block)
        widths.append(block.shape[1])
    trials = None
    if spec.trials is not None:
        trials = np.asarray(_eval(spec.trials, env), dtype=np.float64).reshape(-1)
    return np.ascontiguousarray(np.concatenate(blocks, axis=1)), y, trials, widths


def _expand(name: str, width: int) -> list[str]:
    return [name.replace("*", str(i)) for i in range(1, width + 1)] if "*" in name else [name]


def coefficients(model: str, draws: Mapping[str, Sequence[float]], widths: Sequence[int]) -> tuple[np.ndarray, np.ndarray | None]:
    """(beta [S, D], sigma [S] or None) from a mapping draw column -> values (all chains pooled)."""
    spec = MODELS[model]
    cols = [c for name, width in zip(spec.coef, widths) for c in _expand(name, width)]
    const_sigma = spec.sigma is not None and spec.sigma.startswith("=")
    missing = [c for c in cols + ([spec.sigma] if spec.sigma and not const_sigma else []) if c not in draws]
    if missing:
        raise KeyError(f"draws for {model} lack column(s): {', '.join(missing)}")
    beta = np.stack([np.asarray(draws[c], dtype=np.float64) for c in cols], axis=1)
    sigma = None
    if const_sigma:
        sigma = np.full(beta.shape[0], float(spec.sigma[1:]))
    elif spec.sigma:
        sigma = np.asarray(draws[spec.sigma], dtype=np.float64)
    return np.ascontiguousarray(beta), sigma


def draw_columns(model: str, widths: Sequence[int]) -> list[str]:
    """The draw columns ``coefficients`` reads (column projection for ``store.open_draws``)."""
    spec = MODELS[model]
    cols = [c for name, width in zip(spec.coef, widths) for c in _expand(name, width)]
    if spec.sigma and not spec.sigma.startswith("="):
        cols.append(spec.sigma)
    return cols


def condition(family: str, X: np.ndarray, y: np.ndarray, beta: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Exact linear reparametrisation (fp64, host) that makes the contraction well conditioned for the
    split-bf16 tensor-core kernel, whose error is ~2^-17 of the MAGNITUDE of the terms of eta:

        eta = X beta = b0 + (X - m) beta = b0 + Q (R beta),     b0 = m . beta,   X - m = Q R  (thin QR)

    (m: column means; constant columns fold into b0).  Raw predictors such as ``mom_iq`` ~ 100 or
    ``dist`` ~ 50 would otherwise leave errors of 1e-5 * |eta| on residuals y - eta two orders of magnitude
    smaller, and collinear ones (the mesquite canopy volume / area / shape) make large terms cancel; the
    orthonormal columns of Q do neither.  The normal family sees only y - eta, so there the fit at the mean
    draw moves into the response as well: y - X beta_s = (y - X mean(beta)) - X (beta_s - mean(beta)), and the
    kernel contracts only the draws' deviations (a regression that explains 99.99 % of y, like sblri-blr,
    leaves residuals 250 times smaller than eta).  Returns (X' [N, 1 + D'], y', beta' [S, 1 + D']) with
    eta' = X' beta'^T (normal: y' - eta' = y - eta)."""
    n = X.shape[0]
    m = X.mean(axis=0)
    keep = X.std(axis=0) > 0
    q, r = np.linalg.qr(X[:, keep] - m[keep])
    root_n = np.sqrt(float(n))  # columns of unit mean square, like a standardised predictor
    Xc = np.concatenate([np.ones((n, 1)), q * root_n], axis=1)
    bc = np.concatenate([(beta @ m)[:, None], beta[:, keep] @ r.T / root_n], axis=1)
    yc = y
    if family == "normal":
        centre = bc.mean(axis=0)
        yc = y - Xc @ centre
        bc = bc - centre
    return np.ascontiguousarray(Xc), yc, np.ascontiguousarray(bc)


# Packaged models whose packaged ``stan_data`` is NOT the data their packaged draws were sampled on
# (residual sd of y - X E[beta] over E[sigma]: 296, 0.057, 0.20, 0.73 instead of ~1; measured with the
# reference checkout).  The registry restates their Stan text all the same.
PACKAGED_DATA_MISMATCH = ("blr_informed", "kidscore_momiq_informed", "logearn_height_informed", "mesquite_logvolume_informed")


def packaged_lp(model: str, draws: Mapping[str, Sequence[float]] | None = None, data: Mapping[str, object] | None = None,
                store=None):
    """lp[S] (float64 CUDA tensor) of the likelihood block of a packaged model over draws.

    ``draws`` defaults to the model's packaged reference draws (all chains, table row order) and
    ``data`` to its packaged ``stan_data`` (reference.py:20-22), both read through ``store``.
    """
    from . import logdensity

    if model not in MODELS:
        raise ValueError(f"no GLM restatement registered for model: {model}")
    store = store or PackagedStore()
    if data is None:
        data = store.read_stan_data(model)
    X, y, trials, widths = design(model, data)
    if draws is None:
        reader = store.open_draws(model, params=draw_columns(model, widths))
        table = reader.read_all() if hasattr(reader, "read_all") else reader
        draws = {c: table.column(c).to_numpy(zero_copy_only=False) for c in table.column_names if c not in {"chain", "draw"}}
    beta, sigma = coefficients(model, draws, widths)
    family = MODELS[model].family
    Xc, yc, bc = condition(family, X, y, beta)
    return logdensity.glm_lp(family, Xc, yc, bc, trials=trials, sigma=sigma)
