"""Fixtures for the packaged-model log-density registry (mcmc_ref_b200/models.py): stan_data and a
few packaged draws of a subset of the registered models, copied from the reference checkout, plus
the fp64 oracle's likelihood-block values.  Run in the build container (needs /root/reference):

    python tests/golden/make_models_golden.py

PARITY UNPINNED: the reference never evaluates a log-density; the oracle restates the Stan text.
For three models eta is ALSO computed here by a hand-written formula straight from the Stan file,
so the registry's expression evaluator is checked against an independent restatement.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "mcmc-db_b200"))
from mcmc_ref_b200 import models  # noqa: E402
from mcmc_ref_b200.store import PackagedStore  # noqa: E402
from oracle import glm_oracle  # noqa: E402

DATA = Path("/root/reference/packages/mcmc-ref-data/src/mcmc_ref_data/data")
SUBSET = ["kidiq-kidscore_interaction", "kidiq_with_mom_work-kidscore_mom_work", "kidiq_with_mom_work-kidscore_interaction_z",
          "earnings-logearn_interaction_z", "mesquite-logmesquite_logvas", "nes2000-nes", "sblri-blr",
          "kilpisjarvi_mod-kilpisjarvi", "wells_data-wells_dist", "GLM_Poisson_Data-GLM_Poisson_model",
          "GLM_Binomial_data-GLM_Binomial_model", "eight_schools-eight_schools_noncentered", "varying_slopes-noncentered",
          "bangladesh_contraceptive-noncentered"]
NDRAWS = 96


def hand_eta(model, d, dr):
    """eta [N, S] written directly from the Stan file (independent of models.design)."""
    g = lambda k: np.asarray(d[k], dtype=np.float64)[:, None]  # noqa: E731
    b = lambda k: np.asarray(dr[k], dtype=np.float64)[None, :]  # noqa: E731
    if model == "kidiq-kidscore_interaction":  # kidiq-kidscore_interaction.stan:7-20
        return b("beta[1]") + b("beta[2]") * g("mom_hs") + b("beta[3]") * g("mom_iq") + b("beta[4]") * (g("mom_hs") * g("mom_iq"))
    if model == "earnings-logearn_interaction_z":  # earnings-logearn_interaction_z.stan:7-22
        h = g("height")
        z = (h - h.mean()) / h.std(ddof=1)
        return b("beta[1]") + b("beta[2]") * z + b("beta[3]") * g("male") + b("beta[4]") * (z * g("male"))
    if model == "nes2000-nes":  # nes2000-nes.stan:11-32
        a = g("age_discrete")
        return (b("beta[1]") + b("beta[2]") * g("real_ideo") + b("beta[3]") * g("race_adj") + b("beta[4]") * (a == 2)
                + b("beta[5]") * (a == 3) + b("beta[6]") * (a == 4) + b("beta[7]") * g("educ1") + b("beta[8]") * g("gender")
                + b("beta[9]") * g("income"))
    return None


def main():
    store = PackagedStore(packaged_root=DATA, local_root=Path("/nonexistent"))
    rng = np.random.default_rng(4711)
    out = {}
    for m in SUBSET:
        spec = models.MODELS[m]
        data = store.read_stan_data(m)
        X, y, trials, widths = models.design(m, data)
        cols = models.draw_columns(m, widths)
        try:
            t = store.open_draws(m, params=cols).read_all()
            pick = np.linspace(0, t.num_rows - 1, NDRAWS).astype(int)
            draws = {c: t.column(c).to_numpy()[pick].tolist() for c in cols}
            origin = "packaged draws (evenly spaced rows)"
        except FileNotFoundError:
            # the checkout lacks this model's parquet (.MISSING_LARGE_BLOBS): draws around a crude fit
            if spec.family == "poisson_log":
                b0 = np.linalg.lstsq(X, np.log(y + 0.5), rcond=None)[0]
            elif spec.family == "bernoulli_logit":
                b0 = np.linalg.lstsq(X, 4.0 * (y - 0.5), rcond=None)[0]
            else:
                p = (y + 0.5) / (trials + 1.0)
                b0 = np.linalg.lstsq(X, np.log(p / (1 - p)), rcond=None)[0]
            bs = b0[None, :] + 0.01 * (1.0 + np.abs(b0[None, :])) * rng.standard_normal((NDRAWS, len(b0)))
            draws = {c: bs[:, i].tolist() for i, c in enumerate(cols)}
            origin = "synthetic draws around a least-squares fit (no packaged parquet in this checkout)"
        beta, sigma = models.coefficients(m, draws, widths)
        lp = glm_oracle.glm_lp(spec.family, X, y, beta, trials=trials, sigma=sigma)
        scale = glm_oracle.glm_term_scale(spec.family, X, y, beta, trials=trials, sigma=sigma)
        he = hand_eta(m, data, draws)
        if he is not None:
            assert np.allclose(he, X @ beta.T, rtol=1e-13, atol=1e-12), m
            sg = np.asarray(draws["sigma"])  # the three hand-checked models are normal-family
            lp_hand = -len(y) * np.log(sg) - ((y[:, None] - he) ** 2).sum(axis=0) / (2 * sg * sg)
            assert np.allclose(lp_hand, lp, rtol=1e-12), m
        keys = sorted({k for k in data if isinstance(data[k], (list, int, float))})
        out[m] = {"family": spec.family, "origin": origin, "data": {k: data[k] for k in keys}, "draws": draws,
                  "lp_oracle": lp.tolist(), "term_scale": scale.tolist(), "hand_checked": he is not None}
        print(m, spec.family, X.shape, origin, "hand-checked" if he is not None else "")
    path = Path(__file__).with_name("models_subset.json")
    path.write_text(json.dumps(out, separators=(",", ":")))
    print(path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
