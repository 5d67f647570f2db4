"""Golden values for RAGGED chains (different lengths), produced by the UNMODIFIED reference.

    python tests/golden/make_ragged_golden.py      (build container: /root/reference present)

The reference rank-normalises over all draws, splits each chain at its own midpoint and truncates every
(split) chain to the shortest one (diagnostics.py:76-85, :101-133, :140, :158).  -> ragged_cases.json
"""

from __future__ import annotations

import json
import math
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference/src")
from mcmc_ref import diagnostics as ref  # noqa: E402

HERE = Path(__file__).resolve().parent


def enc(v):
    if isinstance(v, float) and math.isnan(v):
        return "nan"
    if isinstance(v, float) and math.isinf(v):
        return "inf" if v > 0 else "-inf"
    return v


def ar1(rng, n, phi):
    x = np.empty(n)
    x[0] = rng.standard_normal()
    for i in range(1, n):
        x[i] = phi * x[i - 1] + math.sqrt(1 - phi * phi) * rng.standard_normal()
    return x


def main() -> None:
    rng = np.random.default_rng(20260119)
    cases = {
        "normal_100_80_120_95": [rng.standard_normal(n).tolist() for n in (100, 80, 120, 95)],
        "tiny_7_5_9_6": [rng.standard_normal(n).tolist() for n in (7, 5, 9, 6)],
        "halves_of_one_3_2_2_3": [rng.standard_normal(n).tolist() for n in (3, 2, 2, 3)],
        "tied_60_45_70_50": [np.round(rng.standard_normal(n), 1).tolist() for n in (60, 45, 70, 50)],
        "one_empty_10_0_10_10": [rng.standard_normal(n).tolist() for n in (10, 0, 10, 10)],
        "one_single_40_1_40_40": [rng.standard_normal(n).tolist() for n in (40, 1, 40, 40)],
        "sticky_400_350_500_420": [ar1(rng, n, 0.9).tolist() for n in (400, 350, 500, 420)],
        "shifted_300_280_310_290_305": [(rng.standard_normal(n) + 0.4 * (i % 2)).tolist()
                                        for i, n in enumerate((300, 280, 310, 290, 305))],
        "five_chains_odd_33_31_35_37_39": [rng.standard_normal(n).tolist() for n in (33, 31, 35, 37, 39)],
    }
    out = []
    for name, chains in cases.items():
        out.append({"name": name, "chains": chains, "rhat": enc(ref.split_rhat(chains)),
                    "ess_bulk": enc(ref.ess_bulk(chains)), "ess_tail": enc(ref.ess_tail(chains))})
        print(name, out[-1]["rhat"], out[-1]["ess_bulk"], out[-1]["ess_tail"])
    (HERE / "ragged_cases.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
