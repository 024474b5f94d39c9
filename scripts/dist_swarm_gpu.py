"""Multi-GPU check of the sharded PSO swarm (run under torchrun, one process per GPU, NCCL; also works with one process): the
shards' best positions after G generations must equal, bit for bit, pgc_pso_evolve_device on the whole swarm on one GPU.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/dist_swarm_gpu.py [atoms] [swarm] [lbest|gbest]"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi  # noqa: E402
from pagmo2_b200.swarm import DeviceShard, GbestShard, GbestSwarm, ShardedSwarm  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ATOMS = int(sys.argv[1]) if len(sys.argv) > 1 else 38
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
TOPO = sys.argv[3] if len(sys.argv) > 3 else "lbest"
GENS, RADIUS, SEED = 6, 2, 9
ctx = capi.Context(local)
prob = capi.Problem(ctx, "lennard_jones", dim=ATOMS)
lb, ub = prob.bounds()
x = np.random.default_rng(4).uniform(lb, ub, (N, prob.nx))  # the same swarm on every rank
f = prob.eval_host(x)[:, 0]
n_loc = N // world
sl = slice(rank * n_loc, (rank + 1) * n_loc)
if TOPO == "gbest":
    swarm = GbestSwarm(GbestShard(ctx, prob, x[sl], f[sl], rank * n_loc), seed=SEED, first_generation=1)
else:
    swarm = ShardedSwarm(DeviceShard(ctx, prob, x[sl], f[sl], rank * n_loc, RADIUS), seed=SEED, first_generation=1)
swarm.evolve(1)
ctx.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
swarm.evolve(GENS - 1)
ctx.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
bx, bf = swarm.shard.best()
if world > 1:
    tx = torch.from_numpy(bx).cuda()
    parts = [torch.empty_like(tx) for _ in range(world)]
    dist.all_gather(parts, tx)
    bx_all = np.vstack([p.cpu().numpy() for p in parts])
else:
    bx_all = bx
if rank == 0:
    lbx, lbf, _, _ = prob.pso_evolve(x, f, gens=GENS, seed=SEED, first_generation=1, neighb_type=1 if TOPO == "gbest" else 2)  # whole swarm, one GPU
    out = {"world": world, "topology": TOPO, "atoms": ATOMS, "swarm": N, "generations": GENS, "identical_to_single_gpu": bool(np.array_equal(bx_all, lbx)),
           "seconds_for_generations_2_to_6": dt, "generations_per_s": (GENS - 1) / dt, "best_f": float(lbf.min())}
    print(json.dumps(out))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"dist_swarm_{TOPO}_{world}gpu.json").write_text(json.dumps(out, indent=1))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
