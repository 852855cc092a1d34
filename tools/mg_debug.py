"""per-step timing components of the host-buffer call under torchrun (debug aid)"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import sabc_b200 as sb
from bench import workload
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
model, prior, alg, npg, desc = workload("c4")
N = npg * world
comm = sb.api._distributed_setup("torch") if world > 1 else (0, 1, None)
eng = sb.Engine(model, prior, n_particles=N, algorithm=alg, proposal=sb.DifferentialEvolution(n_para=4), resample=2 * N, v=1.0, delta=0.1,
                device=lr, rank=comm[0], world_size=comm[1], nccl_unique_id=comm[2])
eng.init(); eng.update(5 * N)
nl = eng.n_local
bufs = []
for sz in (nl * 4, nl * 3, nl * 3):
    p = C.c_void_p(); sb._lib.check(sb._lib.lib().sabc_host_alloc(C.byref(p), sz * 8))
    bufs.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(sz,)))
th, uu, rr = eng.get_population()
bufs[0][:] = th.ravel(order="F"); bufs[1][:] = uu.ravel(order="F"); bufs[2][:] = rr.ravel(order="F")
eps, cnt = eng.get_state(); eps = eps.copy(); cnt = cnt.copy()
for i in range(30):
    t0 = time.perf_counter()
    eng.update_host(bufs[0], bufs[1], bufs[2], eps, cnt, N)
    t1 = time.perf_counter()
    t = eng.timing()
    if rank == 0:
        print(f"step {i}: wall {1e3 * (t1 - t0):.2f} ms  h2d {t['h2d_ms']:.2f} update {t['update_ms']:.2f} d2h {t['d2h_ms']:.2f}  n_res {cnt[2]} resample_ms {t['resample_ms']:.2f}", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
