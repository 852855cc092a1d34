import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sabc_b200 as sb, oracle_binding as ob, numpy as np
from bench import workload
m, pr, alg, n, desc = workload("c2")
N = 100000
for flags in (0, sb.SABC_FLAG_NO_GRAPH):
    kw = dict(n_particles=N, algorithm=alg, proposal=sb.DifferentialEvolution(n_para=2), resample=2 * N, v=1.0, delta=0.1)
    e = sb.Engine(m, pr, flags=flags, **kw); o = ob.OracleEngine(m, pr, **kw)
    e.init(); o.init()
    for k in range(12):
        try:
            e.update(20 * N)
        except Exception as ex:
            print("flags", flags, "chunk", k, "ERR", ex); eh, uh, rh = e.get_history(); print(eh[-3:], uh[-3:]); break
        o.update(20 * N)
        same = all(np.array_equal(a, b) for a, b in zip(e.get_population(), o.get_population()))
        print("flags", flags, "chunk", k, "same", same, e.get_state()[0], o.get_state()[0], e.get_state()[1], o.get_state()[1], flush=True)
        if not same:
            break
