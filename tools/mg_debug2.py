"""device-resident multi-GPU loop: total vs resampling time (debug aid)"""
import os, sys
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
for flags in (0, sb.SABC_FLAG_NO_GRAPH):
    comm = sb.api._distributed_setup("torch") if world > 1 else (0, 1, None)
    eng = sb.Engine(model, prior, n_particles=N, algorithm=alg, proposal=sb.DifferentialEvolution(n_para=4), resample=2 * N, v=1.0, delta=0.1,
                    device=lr, rank=comm[0], world_size=comm[1], nccl_unique_id=comm[2], flags=flags)
    eng.init(); eng.update(3 * N)
    for chunk in (20, 60, 120):
        eng.update(chunk * N)
        t = eng.timing()
        if rank == 0:
            print(f"world {world} flags {flags} steps {chunk}: {t['update_ms'] / chunk:.3f} ms/step, resample {t['resample_ms']:.2f} ms in {t['resample_events']} events, "
                  f"n_res {eng.get_state()[1][2]}", flush=True)
    eng.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
