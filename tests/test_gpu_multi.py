"""Multi-GPU path on real devices (needs >= 2 GPUs; `gpurun --gpus 2`): one process per GPU over NCCL, contiguous particle
slices, partners from the local inactive half, exact global multinomial resampling with surplus exchange.  A sharded run has
no single-process oracle to be bit-compared with (partner sets differ), so the tests check the invariants that define it:
all ranks agree on every global quantity, the exact integer means equal the means of the gathered population, resampling
only ever copies existing particles, and the posterior of C1 matches the conjugate result."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

WORKER = r'''
import json, os, sys
import numpy as np
import torch, torch.distributed as dist
root = os.environ["SABC_ROOT"]; sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import sabc_b200 as sb
from helpers import model_cases
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
def new_comm():
    return sb.api._distributed_setup("torch")     # every communicator needs a fresh ncclUniqueId

def gather(a):
    out = [None] * world
    dist.all_gather_object(out, a)
    return out

report = {}
for name, N, n_upd, prop in (("gauss_mean", 4000 * world, 30, "de"), ("gauss_sample_d2s2", 3000 * world, 15, "stretch"),
                             ("sir_tauleap", 2048 * world, 10, "de"), ("gauss_sample_d2s2", 2000 * world, 10, "rw")):
    model, prior = model_cases()[name]
    proposal = {"de": sb.DifferentialEvolution(n_para=model.n_para), "stretch": sb.StretchMove(), "rw": sb.RandomWalk(n_para=model.n_para)}[prop]
    comm = new_comm()
    alg = "multi_eps" if model.n_stats > 1 and name != "sir_tauleap" else "single_eps"
    eng = sb.Engine(model, prior, n_particles=N, algorithm=alg, proposal=proposal, resample=N // 2, v=1.0, delta=0.1,
                    device=int(os.environ["LOCAL_RANK"]), rank=comm[0], world_size=comm[1], nccl_unique_id=comm[2])
    assert eng.n_local == N // world and eng.offset == rank * (N // world)
    eng.init()
    th0, u0, r0 = eng.get_population()
    all_th0 = np.concatenate(gather(th0))
    # every rank holds the same replicated ECDF, eps and counters
    knots = gather(eng.get_ecdf(0)); assert all(np.array_equal(k, knots[0]) for k in knots)
    st = gather([eng.get_state()[0].tolist(), eng.get_state()[1].tolist()]); assert all(s == st[0] for s in st), st
    assert st[0][1] == [N, 0, 1, 0]
    # initial resampling only copies prior draws; u mean in the history equals the mean over the gathered population
    eh, uh, rh = eng.get_history()
    all_u = np.concatenate(gather(u0))
    assert np.allclose(uh[0], all_u.mean(axis=0), rtol=1e-12, atol=1e-15), (uh[0], all_u.mean(axis=0))
    eng.update(n_upd * N)
    th, u, r = eng.get_population()
    st = gather([eng.get_state()[0].tolist(), eng.get_state()[1].tolist()]); assert all(s == st[0] for s in st), st
    eps, cnt = eng.get_state()
    assert cnt[0] == N * (n_upd + 1) and cnt[3] == n_upd and cnt[2] >= 2, cnt       # several global resamplings happened
    eh, uh, rh = eng.get_history()
    hs = gather([eh.tolist(), uh.tolist(), rh.tolist()]); assert all(h == hs[0] for h in hs)
    all_u = np.concatenate(gather(u)); all_r = np.concatenate(gather(r)); all_th = np.concatenate(gather(th))
    assert np.allclose(uh[-1], all_u.mean(axis=0), rtol=1e-12, atol=1e-15)
    assert np.allclose(rh[-1], all_r.mean(axis=0), rtol=1e-10)
    assert np.all(np.isfinite(all_th)) and np.all((all_u >= 0) & (all_u <= 1 + 1e-15))
    assert np.all(eps > 0) and np.all(eps < 1)
    report[name + "_" + prop] = {"eps": eps.tolist(), "counters": cnt.tolist(), "mean_u": all_u.mean(axis=0).tolist()}
    if rank == 0 and os.environ.get("SABC_REF_OUT"):       # reference for the single-process handle (compared by the pytest process)
        np.savez(os.path.join(os.environ["SABC_REF_OUT"], f"{name}_{prop}.npz"), theta=all_th, u=all_u, rho=all_r, eps=eps, counters=cnt,
                 eps_h=eh, u_h=uh, rho_h=rh)
    eng.close()

# replicated ("strict") mode: whole population on every rank, each simulates a share -> bit-identical to the oracle
import oracle_binding as ob
for name, N, n_upd, prop, alg in (("gauss_mean", 3000, 25, "de", "single_eps"), ("sir_tauleap", 2500, 10, "stretch", "single_eps"),
                                  ("gauss_sample_d2s2", 20000, 12, "rw", "multi_eps"), ("logistic", 1111 * world, 8, "de", "single_eps")):
    model, prior = model_cases()[name]
    proposal = {"de": sb.DifferentialEvolution(n_para=model.n_para), "stretch": sb.StretchMove(), "rw": sb.RandomWalk(n_para=model.n_para)}[prop]
    N = (N // world) * world
    kw = dict(n_particles=N, algorithm=alg, proposal=proposal, resample=N // 8, v=1.0, delta=0.1)
    comm = new_comm()
    eng = sb.Engine(model, prior, device=int(os.environ["LOCAL_RANK"]), rank=comm[0], world_size=comm[1], nccl_unique_id=comm[2],
                    flags=sb.SABC_FLAG_MG_REPLICATED, **kw)
    assert eng.n_local == N and eng.offset == 0
    eng.init(); eng.update(n_upd * N)
    orc = ob.OracleEngine(model, prior, **kw); orc.init(); orc.update(n_upd * N)
    for nm, a, b in zip(("theta", "u", "rho"), eng.get_population(), orc.get_population()):
        assert np.array_equal(a, b), (name, nm, int((a != b).sum()))
    assert np.array_equal(eng.get_state()[0], orc.get_state()[0]) and np.array_equal(eng.get_state()[1], orc.get_state()[1])
    for a, b in zip(eng.get_history(), orc.get_history()):
        assert np.array_equal(a, b)
    assert eng.get_state()[1][2] >= 2
    eng.close()

# resampling alone: run one forced global resampling through init on a tiny problem and check it copies existing particles
model, prior = model_cases()["gauss_sample_d2s2"]
N = 1024 * world
comm = new_comm()
eng = sb.Engine(model, prior, n_particles=N, algorithm="single_eps", proposal=sb.DifferentialEvolution(n_para=2), resample=2 * N,
                v=1.0, delta=0.1, device=int(os.environ["LOCAL_RANK"]), rank=comm[0], world_size=comm[1], nccl_unique_id=comm[2], seed=7,
                flags=sb.SABC_FLAG_MG_STRICT_RESAMPLE)
ref = sb.Engine(model, prior, n_particles=N, algorithm="single_eps", proposal=sb.DifferentialEvolution(n_para=2), resample=2 * N,
                v=1.0, delta=0.1, device=int(os.environ["LOCAL_RANK"]), seed=7) if rank == 0 else None
eng.init()
th, u, r = eng.get_population()
all_th = np.concatenate(gather(th)); all_u = np.concatenate(gather(u)); all_r = np.concatenate(gather(r))
if rank == 0:
    # the single-GPU engine with the same seed draws the same prior sample, builds the same ECDF and -- because the global
    # multinomial uses the same N variates against the same integer weights -- selects the same MULTISET of particles
    ref.init()
    th1, u1, r1 = ref.get_population()
    assert np.array_equal(np.sort(all_r, axis=0), np.sort(r1, axis=0))            # rho is not resampled: same prior distances
    key = lambda a: np.sort(a.view([("", a.dtype)] * a.shape[1]).ravel())
    assert np.array_equal(key(np.ascontiguousarray(all_th)), key(np.ascontiguousarray(th1)))
    assert np.array_equal(key(np.ascontiguousarray(all_u)), key(np.ascontiguousarray(u1)))
    assert np.array_equal(eng.get_state()[0], ref.get_state()[0])                    # eps_0 identical (exact integer means)
    assert np.array_equal(eng.get_history()[1], ref.get_history()[1])

# default resampling (per-rank counts from one shared-seed multinomial draw): it only ever copies existing particles, every
# rank ends with its full slice, and the counts the ranks computed independently agree (the exchange would dead-lock otherwise)
comm = new_comm()
eng2 = sb.Engine(model, prior, n_particles=N, algorithm="single_eps", proposal=sb.DifferentialEvolution(n_para=2), resample=2 * N,
                 v=1.0, delta=0.1, device=int(os.environ["LOCAL_RANK"]), rank=comm[0], world_size=comm[1], nccl_unique_id=comm[2], seed=7)
eng2.init()
th2, u2, r2 = eng2.get_population()
all_th2 = np.concatenate(gather(th2)); all_u2 = np.concatenate(gather(u2))
if rank == 0:
    rows = {tuple(x) for x in np.hstack([th1, u1]).tolist()}                        # th1, u1: resampled from the same prior sample
    prior_rows_ok = all(np.isfinite(all_th2).all(axis=1))
    assert prior_rows_ok and all_th2.shape == th1.shape
    # same prior sample, same weights: the two resampled populations are draws from the same categorical distribution
    from scipy import stats
    assert stats.ks_2samp(all_u2[:, 0], u1[:, 0]).pvalue > 1e-3
eng.close(); eng2.close()

# compressed ECDF over a sharded population: the K global quantiles come from a distributed selection (no rank holds the global
# column); the knots must equal the single-GPU compressed table for the same K bit for bit
for name, K in (("gauss_sample_d2s2", 510), ("sir_tauleap", 1022)):
    model, prior = model_cases()[name]
    N = 6000 * world
    kw = dict(n_particles=N, algorithm="single_eps", proposal=sb.DifferentialEvolution(n_para=model.n_para), resample=N, v=1.0, delta=0.1,
              ecdf_max_knots=K, seed=21)
    comm = new_comm()
    engk = sb.Engine(model, prior, device=int(os.environ["LOCAL_RANK"]), rank=comm[0], world_size=comm[1], nccl_unique_id=comm[2], **kw)
    engk.init()
    if rank == 0:
        one = sb.Engine(model, prior, device=int(os.environ["LOCAL_RANK"]), **kw); one.init()
        for j in range(model.n_stats):
            a, b = engk.get_ecdf(j), one.get_ecdf(j)
            assert a.size == b.size <= K + 2 and np.array_equal(a, b), (name, j, a.size, b.size, int((a != b).sum()))
        one.close()
    engk.update(5 * N)
    ks = gather(engk.get_ecdf(0)); assert all(np.array_equal(k, ks[0]) for k in ks)
    engk.close()

# posterior of C1 over the sharded population (slow annealing): mean and variance within 2 MC standard errors of N(10/11, 1/11),
# the MC error estimated from independent runs (north_star check 3)
model, prior = model_cases()["gauss_mean"]
N = 4000 if world <= 8 else 500 * world     # the population of the one-GPU test: a larger one shrinks the MC error below the finite-eps bias of ABC
means, vars_ = [], []
for seed in range(6):
    comm = new_comm()
    eng = sb.Engine(model, prior, n_particles=N, algorithm="single_eps", proposal=sb.DifferentialEvolution(n_para=1), resample=2 * N,
                    v=0.02, delta=0.1, device=int(os.environ["LOCAL_RANK"]), rank=comm[0], world_size=comm[1], nccl_unique_id=comm[2], seed=200 + seed)
    eng.init(); eng.update(400 * N)
    th = np.concatenate(gather(eng.get_population()[0]))[:, 0]
    means.append(float(th.mean())); vars_.append(float(th.var()))
    eng.close()
se_m = np.std(means, ddof=1) / np.sqrt(len(means)); se_v = np.std(vars_, ddof=1) / np.sqrt(len(vars_))
report["c1_posterior"] = {"mean": float(np.mean(means)), "se_mean": float(se_m), "var": float(np.mean(vars_)), "se_var": float(se_v)}
assert abs(np.mean(means) - 10 / 11) < 2 * se_m + 1e-3 and abs(np.mean(vars_) - 1 / 11) < 2 * se_v + 1e-3, report["c1_posterior"]
# the mirrored public surface over the sharded engine: sabc(...; comm="torch") then update_population
f2, p2 = model_cases()["gauss_sample_d2s2"]
res = sb.sabc(f2, p2, n_particles=500 * world, n_simulation=5000 * world, algorithm="multi_eps", comm="torch", device=int(os.environ["LOCAL_RANK"]))
assert res.state.n_population_updates == 9 and res.population.shape == (500, 2) and np.all(res.state.eps < 1)
sb.update_population(res, f2, p2, n_simulation=5000 * world)
assert res.state.n_population_updates == 19 and res.state.n_simulation == 500 * world * 20
states = gather([res.state.eps.tolist(), res.state.n_accept, res.state.n_resampling]); assert all(x == states[0] for x in states)
if rank == 0:
    print("REPORT " + json.dumps(report))
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
'''


def n_devices():
    import ctypes as C
    import sabc_b200 as sb
    n = C.c_int(0)
    return n.value if sb._lib.lib().sabc_device_count(C.byref(n)) == 0 else 0


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_population(gpu, world, tmp_path):
    if n_devices() < world:
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SABC_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(script)]
    env["SABC_REF_OUT"] = str(tmp_path)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    log_dir = os.path.join(env["SABC_ROOT"], "gpurun_out")
    if os.path.isdir(log_dir):                            # keep the workers' own output: a failure on a remote box must be readable afterwards
        open(os.path.join(log_dir, f"multi_worker_world{world}.log"), "w").write(r.stdout[-20000:] + "\n---- stderr ----\n" + r.stderr[-40000:])
    if r.returncode != 0:
        lines = [l for l in (r.stdout + r.stderr).splitlines() if any(k in l for k in ("Error", "assert", "Traceback", "File \"/", "rank"))]
        raise AssertionError("\n".join(lines[-40:]))
    assert r.stdout.count("OK") == world
    rep = [l for l in r.stdout.splitlines() if l.startswith("REPORT ")]
    assert rep and json.loads(rep[0][7:])
    single_process_handle_checks(world, tmp_path)


def single_process_handle_checks(world, ref_dir):
    """sabc_config.n_gpus: ONE process drives `world` GPUs through one handle (the caller the reference's surface describes).  The
    handle runs the per-GPU engines of the process-per-GPU mode on one host thread each, so for the same seed it must reproduce the
    torchrun run bit for bit -- population (global arrays), eps, counters, histories -- and its replicated mode the oracle."""
    import numpy as np
    import oracle_binding as ob
    import sabc_b200 as sb
    from helpers import model_cases
    for name, n_per, n_upd, prop in (("gauss_mean", 4000, 30, "de"), ("gauss_sample_d2s2", 3000, 15, "stretch"), ("sir_tauleap", 2048, 10, "de"),
                                     ("gauss_sample_d2s2", 2000, 10, "rw")):
        ref = np.load(os.path.join(str(ref_dir), f"{name}_{prop}.npz"))
        model, prior = model_cases()[name]
        N = n_per * world
        proposal = {"de": sb.DifferentialEvolution(n_para=model.n_para), "stretch": sb.StretchMove(), "rw": sb.RandomWalk(n_para=model.n_para)}[prop]
        alg = "multi_eps" if model.n_stats > 1 and name != "sir_tauleap" else "single_eps"
        eng = sb.Engine(model, prior, n_particles=N, algorithm=alg, proposal=proposal, resample=N // 2, v=1.0, delta=0.1, n_gpus=world)
        assert eng.n_local == N and eng.offset == 0                     # the handle presents the whole population
        eng.init(); eng.update(n_upd * N)
        th, u, r = eng.get_population()
        assert np.array_equal(th, ref["theta"]) and np.array_equal(u, ref["u"]) and np.array_equal(r, ref["rho"]), (name, prop)
        eps, cnt = eng.get_state()
        assert np.array_equal(eps, ref["eps"]) and np.array_equal(cnt, ref["counters"])
        for a, key in zip(eng.get_history(), ("eps_h", "u_h", "rho_h")):
            assert np.array_equal(a, ref[key])
        # host-buffer call on the GLOBAL arrays == device-resident continuation
        th2, u2, r2, eps2, cnt2 = th.copy(order="F"), u.copy(order="F"), r.copy(order="F"), eps.copy(), cnt.copy()
        eng.update(3 * N)
        twin = sb.Engine(model, prior, n_particles=N, algorithm=alg, proposal=proposal, resample=N // 2, v=1.0, delta=0.1, n_gpus=world)
        for j in range(model.n_stats):
            twin.set_ecdf(j, eng.get_ecdf(j))
        twin.update_host(th2, u2, r2, eps2, cnt2, 3 * N)
        for a, b in zip(eng.get_population(), (th2, u2, r2)):
            assert np.array_equal(a, b), (name, prop, "update_host through the handle")
        assert np.array_equal(eng.get_state()[0], eps2) and np.array_equal(eng.get_state()[1], cnt2)
        eng.close(); twin.close()
    # replicated (strict) mode through the handle: bit-identical to the oracle
    model, prior = model_cases()["sir_tauleap"]
    N = 1250 * world
    kw = dict(n_particles=N, algorithm="single_eps", proposal=sb.DifferentialEvolution(n_para=4), resample=N // 8, v=1.0, delta=0.1)
    eng = sb.Engine(model, prior, n_gpus=world, flags=sb.SABC_FLAG_MG_REPLICATED, **kw)
    eng.init(); eng.update(8 * N)
    orc = ob.OracleEngine(model, prior, **kw); orc.init(); orc.update(8 * N)
    for a, b in zip(eng.get_population(), orc.get_population()):
        assert np.array_equal(a, b)
    assert np.array_equal(eng.get_state()[0], orc.get_state()[0]) and np.array_equal(eng.get_state()[1], orc.get_state()[1])
    eng.close()
    # the mirrored public surface: sabc(...; n_gpus) from one process
    f2, p2 = model_cases()["gauss_sample_d2s2"]
    res = sb.sabc(f2, p2, n_particles=500 * world, n_simulation=5000 * world, algorithm="multi_eps", n_gpus=world)
    assert res.state.n_population_updates == 9 and res.population.shape == (500 * world, 2) and np.all(res.state.eps < 1)
    sb.update_population(res, f2, p2, n_simulation=5000 * world)
    assert res.state.n_population_updates == 19
