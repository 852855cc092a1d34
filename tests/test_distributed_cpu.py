"""world_size-2 (and 4) gloo tests on CPU for the host side of the multi-GPU path: the ncclUniqueId rendezvous and the surplus
exchange of the global resampling, executed for real with gloo send/recv according to sabc_mg_exchange_plan (the same plan
multi_gpu.inl feeds to ncclSend/ncclRecv)."""
import os
import socket
import subprocess
import sys

import pytest

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["SABC_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SABC_ROOT"], "tests"))
import sabc_b200 as sb
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# 1. rendezvous: every rank ends up with rank 0's 128-byte id
r, w, uid = sb.api._distributed_setup("torch")
assert (r, w) == (rank, world) and len(uid) == 128
ids = [None] * world
dist.all_gather_object(ids, uid)
assert all(i == ids[0] for i in ids) and any(b != 0 for b in uid)
# 2. exchange: global draws select counts[g] particles on rank g; slot p of the concatenated selection goes to rank p // n
n_local = 1000 + 37 * world
for trial in range(5):
    rng = np.random.default_rng(100 + trial)                   # same stream on every rank
    counts = rng.multinomial(n_local * world, rng.dirichlet(np.ones(world) * 3)).astype(np.int64)
    C0 = int(counts[:rank].sum())
    mine = np.arange(C0, C0 + counts[rank], dtype=np.float64) * 10 + rank / 10     # payload tagged with its global slot
    arrs = [np.zeros(world, dtype=np.int64) for _ in range(4)]
    assert sb._lib.lib().sabc_mg_exchange_plan(sb._lib.ptr(counts), world, n_local, rank, *[sb._lib.ptr(a) for a in arrs]) == 0
    s_off, s_cnt, r_off, r_cnt = arrs
    out = np.full(n_local, np.nan)
    out[r_off[rank]:r_off[rank] + r_cnt[rank]] = mine[s_off[rank]:s_off[rank] + s_cnt[rank]]
    reqs = []
    bufs = {}
    for g in range(world):
        if g != rank and r_cnt[g] > 0:
            bufs[g] = torch.empty(int(r_cnt[g]), dtype=torch.float64)
            reqs.append(dist.irecv(bufs[g], src=g))
    for d in range(world):
        if d != rank and s_cnt[d] > 0:
            reqs.append(dist.isend(torch.from_numpy(mine[s_off[d]:s_off[d] + s_cnt[d]].copy()), dst=d))
    for q in reqs:
        q.wait()
    for g, b in bufs.items():
        out[r_off[g]:r_off[g] + r_cnt[g]] = b.numpy()
    slots = np.floor(out / 10 + 1e-9)
    assert np.array_equal(slots, np.arange(rank * n_local, (rank + 1) * n_local)), (rank, trial)
    dist.barrier()
# 3. integer all-reduce of the u limbs is exact and order-free: emulate with int64 tensors
u = np.random.default_rng(7).random(n_local * world)
q = (u * 2.0 ** 62).astype(np.uint64)
hi, lo = (q >> np.uint64(31)).astype(np.int64), (q & np.uint64(0x7fffffff)).astype(np.int64)
sl = slice(rank * n_local, (rank + 1) * n_local)
t = torch.tensor([hi[sl].sum(), lo[sl].sum()], dtype=torch.int64)
dist.all_reduce(t)
assert int(t[0]) == int(hi.sum()) and int(t[1]) == int(lo.sum())
# 4. the sharded resampling's host half, as multi_gpu.inl runs it: every rank all-gathers the per-rank weight totals, computes the
#    multinomial split of the N draws on its own (same seed, same resampling count -> same counts, or the exchange would
#    dead-lock), packs c_me "selected particles" and runs the surplus exchange of the plan; every rank must end with a full slice
rng = np.random.default_rng(5 + rank)
for rc in range(4):
    w_me = int(rng.integers(2**35, 2**40)) if not (rc == 3 and rank == 0) else 0      # last round: one rank weighs nothing
    ws = [None] * world
    dist.all_gather_object(ws, w_me)
    w = np.array(ws, dtype=np.uint64)
    counts = np.zeros(world, dtype=np.int64)
    assert sb._lib.lib().sabc_multinomial_split(n_local * world, sb._lib.ptr(w), world, 0x5ABC, rc, sb._lib.ptr(counts)) == 0
    allc = [None] * world
    dist.all_gather_object(allc, counts.tolist())
    assert all(c == allc[0] for c in allc) and sum(allc[0]) == n_local * world
    if w_me == 0:
        assert counts[rank] == 0
    C0 = int(counts[:rank].sum())
    mine = np.arange(C0, C0 + counts[rank], dtype=np.float64)                         # payload = global slot of the selection
    arrs = [np.zeros(world, dtype=np.int64) for _ in range(4)]
    assert sb._lib.lib().sabc_mg_exchange_plan(sb._lib.ptr(counts), world, n_local, rank, *[sb._lib.ptr(a) for a in arrs]) == 0
    s_off, s_cnt, r_off, r_cnt = arrs
    out = np.full(n_local, np.nan)
    out[r_off[rank]:r_off[rank] + r_cnt[rank]] = mine[s_off[rank]:s_off[rank] + s_cnt[rank]]
    reqs, bufs = [], {}
    for g in range(world):
        if g != rank and r_cnt[g] > 0:
            bufs[g] = torch.empty(int(r_cnt[g]), dtype=torch.float64); reqs.append(dist.irecv(bufs[g], src=g))
    for d in range(world):
        if d != rank and s_cnt[d] > 0:
            reqs.append(dist.isend(torch.from_numpy(mine[s_off[d]:s_off[d] + s_cnt[d]].copy()), dst=d))
    for q in reqs:
        q.wait()
    for g, b in bufs.items():
        out[r_off[g]:r_off[g] + r_cnt[g]] = b.numpy()
    assert np.array_equal(out, np.arange(rank * n_local, (rank + 1) * n_local, dtype=np.float64)), (rank, rc)
    dist.barrier()
dist.destroy_process_group()
print("OK", rank)
'''


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_rendezvous_and_exchange(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SABC_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == world
