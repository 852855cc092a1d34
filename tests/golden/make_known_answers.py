"""Regenerates tests/golden/known_answers.json.

* philox4x32_10: the published Random123 known-answer vectors (kat_vectors), copied verbatim.
* update_epsilon_single_eps / update_epsilon_multi_eps / build_cdf: restatement-derived known answers of SURVEY.md App. F --
  the formulas of src/SimulatedAnnealingABC.jl:92-117 and src/cdf_estimators.jl:23-44 evaluated with scipy (brentq, rtol 1e-15)
  and exact rational arithmetic.  They are NOT outputs of the Julia reference (julia is not installed in this image).

    python tests/golden/make_known_answers.py [--check]
"""
import json
import math
import os
import sys
from fractions import Fraction

from scipy.optimize import brentq

HERE = os.path.dirname(os.path.abspath(__file__))


def eps_single(ubar, v):                      # root of eps^2 + v eps^(3/2) - ubar^2 on (0, ubar)   (:92-95)
    return brentq(lambda e: e * e + v * e ** 1.5 - ubar * ubar, 0.0, ubar, xtol=1e-300, rtol=1e-15, maxiter=500)


def eps_multi(ubar, v):                       # :100-117
    n = len(ubar)
    cn = math.factorial(2 * n + 2) / (math.factorial(n + 1) * math.factorial(n + 2))
    out = []
    for ui in ubar:
        q = [u / ui for u in ubar]
        num = 1 + sum(x ** (n / 2) for x in q)
        den = cn * (n + 1) * ui ** (1 + n / 2) * math.prod(q)
        g = lambda b: (1 - math.exp(-b) * (1 + b)) / (b * (1 - math.exp(-b))) - ui
        beta = brentq(g, 1e-3, 1e6, xtol=1e-300, rtol=1e-15, maxiter=500)
        out.append(1 / (beta + v * num / den))
    return out


def build_cdf_at(data, at):                   # cdf_estimators.jl:23-44 + Interpolations semantics, exact rationals
    pos = sorted(Fraction(x) for x in data if x > 0)
    K = [Fraction(0)] + pos + [pos[-1] * Fraction(3, 2)]
    L = len(K)
    res = []
    for x in at:
        xf = K[-1] if x == "Inf" else min(max(Fraction(x), K[0]), K[-1])
        k = next(i for i, kv in enumerate(K) if kv >= xf)          # searchsortedfirst (0-based)
        if k > 0:
            k -= 1
        m = (Fraction(k + 1, L - 1) - Fraction(k, L - 1)) / (K[k + 1] - K[k])
        res.append(float(Fraction(k, L - 1) + m * (xf - K[k])))
    return res


def make():
    return {
        "_provenance": "philox4x32_10: published Random123 kat_vectors. The rest: restatement-derived known answers of SURVEY.md "
                       "App. F (scipy brentq / exact rational arithmetic; NOT produced by the Julia reference, which cannot run here). "
                       "Regenerate with tests/golden/make_known_answers.py.",
        "philox4x32_10": [
            [["0", "0", "0", "0"], ["0", "0"], ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]],
            [["ffffffff"] * 4, ["ffffffff"] * 2, ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]],
            [["243f6a88", "85a308d3", "13198a2e", "03707344"], ["a4093822", "299f31d0"], ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]],
        ],
        "update_epsilon_single_eps": [[u, v, eps_single(u, v)] for u, v in ((0.5, 1.0), (0.3, 1.0), (0.1, 1.0), (0.01, 1.0), (0.3, 0.5), (0.0001, 2.0))],
        "update_epsilon_multi_eps": [[u, v, eps_multi(u, v)] for u, v in (([0.3], 1.0), ([0.3, 0.2], 1.0), ([0.1, 0.1, 0.1], 1.0), ([0.45, 0.05], 10.0))],
        "build_cdf": [
            {"data": [1, 2, 2, 3, 3, 3], "at": [2.0, 2.5, 3.0, "Inf"], "want": build_cdf_at([1, 2, 2, 3, 3, 3], [2.0, 2.5, 3.0, "Inf"])},
            {"data": [1, 0, 2, 0, 3], "at": [2.0, 2.5, 3.0], "want": build_cdf_at([1, 0, 2, 0, 3], [2.0, 2.5, 3.0])},
        ],
    }


if __name__ == "__main__":
    new = make()
    path = os.path.join(HERE, "known_answers.json")
    if "--check" in sys.argv:
        old = json.load(open(path))
        for key in ("update_epsilon_single_eps", "update_epsilon_multi_eps"):
            for a, b in zip(old[key], new[key]):
                xa, xb = (a[2], b[2]) if isinstance(a[2], list) else ([a[2]], [b[2]])
                assert all(abs(p - q) <= 1e-12 * abs(q) for p, q in zip(xa, xb)), (key, a, b)
        assert old["build_cdf"] == new["build_cdf"] and old["philox4x32_10"] == new["philox4x32_10"]
        print("known_answers.json is reproducible")
    else:
        json.dump(new, open(path, "w"), indent=1)
        print("wrote", path)
