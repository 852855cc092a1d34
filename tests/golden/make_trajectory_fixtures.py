"""Generates tests/golden/trajectories.json from the ORACLE: for a few small configurations the counters, the full eps history
and SHA-256 digests of population / u / rho after a fixed number of population updates.  The GPU tests compare the CUDA engine
with these committed values; a CPU test checks that the live oracle still reproduces them (i.e. that nobody changed the written
numerical specification of DESIGN.md section 3 without regenerating the fixtures).

    python tests/golden/make_trajectory_fixtures.py [--check]
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_binding as ob  # noqa: E402
import sabc_b200 as sb  # noqa: E402
from helpers import model_cases  # noqa: E402

CASES = [  # name, N, updates, algorithm, proposal
    ("gauss_mean", 1000, 99, "single_eps", "de"),             # BASELINE configs[0]
    ("gauss_sample_d2s2", 1500, 30, "multi_eps", "de"),
    ("gauss_sample_d2s2", 1500, 30, "single_eps", "stretch"),
    ("gauss_sample_d2s1", 800, 25, "single_eps", "rw"),
    ("logistic", 1200, 15, "single_eps", "de"),
    ("sir_tauleap", 1024, 12, "single_eps", "de"),
    ("sir_gillespie_s3", 2000, 20, "multi_eps", "de"),
    ("gauss_sample_d2s2_gambeta", 1000, 20, "multi_eps", "de"),   # Gamma x Beta prior
    ("gauss_sample_d2s2_laplinvg", 1000, 15, "single_eps", "de"), # Laplace x InverseGamma prior
    ("gauss_sample_d2s2_cauweib", 1000, 15, "multi_eps", "stretch"),  # Cauchy x Weibull prior
]


def proposal_of(kind, d):
    return {"de": sb.DifferentialEvolution(n_para=d), "stretch": sb.StretchMove(), "rw": sb.RandomWalk(n_para=d)}[kind]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes(order="C")).hexdigest()


def run_case(engine_factory, name, N, n_upd, alg, prop):
    model, prior = model_cases()[name]
    e = engine_factory(model, prior, n_particles=N, algorithm=alg, proposal=proposal_of(prop, model.n_para), resample=N, v=1.0, delta=0.1, seed=0x5ABC)
    e.init(); e.update(n_upd * N)
    th, u, rho = e.get_population(); eps, cnt = e.get_state(); eh, uh, rh = e.get_history()
    return {"case": [name, N, n_upd, alg, prop], "counters": [int(c) for c in cnt], "eps": [float(x).hex() for x in eps],
            "eps_history": [[float(x).hex() for x in row] for row in eh], "theta_sha256": digest(np.asfortranarray(th).T),
            "u_sha256": digest(np.asfortranarray(u).T), "rho_sha256": digest(np.asfortranarray(rho).T),
            "u_history_last": [float(x).hex() for x in uh[-1]], "rho_history_last": [float(x).hex() for x in rh[-1]]}


if __name__ == "__main__":
    out = [run_case(ob.OracleEngine, *c) for c in CASES]
    path = os.path.join(HERE, "trajectories.json")
    if "--check" in sys.argv:
        assert json.load(open(path)) == out
        print("trajectories.json is reproducible")
    else:
        json.dump(out, open(path, "w"), indent=1)
        print("wrote", path, [o["counters"] for o in out])
