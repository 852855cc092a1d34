"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel family of the path runs once or more."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sabc_b200 as sb
from helpers import model_cases
for name, prop, alg in (("gauss_mean", sb.DifferentialEvolution(n_para=1), "single_eps"), ("gauss_sample_d2s2", sb.RandomWalk(n_para=2), "multi_eps"),
                        ("sir_tauleap", sb.StretchMove(), "single_eps"), ("logistic", sb.DifferentialEvolution(n_para=3), "single_eps")):
    model, prior = model_cases()[name]
    N = 1500
    eng = sb.Engine(model, prior, n_particles=N, algorithm=alg, proposal=prop, resample=N // 2, v=1.0, delta=0.1)
    eng.init(); eng.update(4 * N)
    th, u, rho = eng.get_population(); eps, cnt = eng.get_state()
    eng.update_host(th, u, rho, eps, cnt, 2 * N)
    print(name, "ok", cnt.tolist(), eps.tolist())
    eng.close()
# pipelined host call (>= 32768 rows per half)
model, prior = model_cases()["gauss_sample_d2s2"]
N = 70_000
eng = sb.Engine(model, prior, n_particles=N, algorithm="multi_eps", proposal=sb.DifferentialEvolution(n_para=2), resample=N // 2, v=1.0, delta=0.1)
eng.init(); th, u, rho = eng.get_population(); eps, cnt = eng.get_state()
eng.update_host(th, u, rho, eps, cnt, 2 * N)
print("pipelined ok", cnt.tolist())
