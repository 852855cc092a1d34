"""C5: particle-count scaling sweep on the 1-D Gaussian (BASELINE.json configs[4]) -- device-resident throughput of the update
loop per population size, plus the other configs at their nominal sizes.  Run under torchrun for N > 1 GPUs.
    python tools/sweep.py [--sizes 10000,100000,...] [--steps 20] [--cpu]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sabc_b200 as sb
from bench import workload, cpu_oracle_throughput

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="10000,100000,1000000,10000000,100000000")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle (rank 0)")
ap.add_argument("--configs", default="c5,c1,c2,c3,c4")
args = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(lr); dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rows = []
for cfg in args.configs.split(","):
    model, prior, alg, n_default, desc = workload(cfg)
    sizes = [int(x) for x in args.sizes.split(",")] if cfg == "c5" else [n_default * world]
    for N in sizes:
        if N % world or N // world < 4:
            continue
        comm = sb.api._distributed_setup("torch") if world > 1 else (0, 1, None)
        eng = sb.Engine(model, prior, n_particles=N, algorithm=alg, proposal=sb.DifferentialEvolution(n_para=model.n_para), resample=2 * N,
                        v=1.0, delta=0.1, device=lr, rank=comm[0], world_size=comm[1], nccl_unique_id=comm[2])
        eng.init(); eng.update(3 * N)
        if dist: dist.barrier()
        eng.update(args.steps * N)
        ms = eng.timing()["update_ms"]
        if dist:
            import torch
            t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        row = {"config": cfg, "n_particles": N, "n_gpus": world, "ms_per_update": ms / args.steps, "updates_per_s": args.steps * N / (ms * 1e-3)}
        if args.cpu and rank == 0:
            val, cores, sample, _ = cpu_oracle_throughput(model, prior, alg, target_seconds=6.0, steps=3, warmup=1, max_particles=min(N, 2_000_000), fixed_steps=False)
            row.update(cpu_updates_per_s=val, cpu_cores=cores, cpu_sample=sample)
        eng.close()
        if rank == 0:
            rows.append(row); print(json.dumps(row), flush=True)
if dist:
    dist.barrier(); dist.destroy_process_group()
