"""Multi-GPU check of the migration path (run under torchrun, one process per GPU, NCCL): 8 device islands (sade on cec2013 f12, D=50,
pop 1024, ring) sharded over the ranks must end with exactly the populations of the same archipelago run inside one process.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/dist_archi_gpu.py"""
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
from pagmo2_b200.archipelago import Archipelago, DeviceIsland  # noqa: E402
from pagmo2_b200 import synth  # noqa: E402  (synthetic data tables)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mr, os_ = synth.cec2013_tables(50)
ISLANDS, ROUNDS = 8, 4


def make(device):
    def f(g):
        return DeviceIsland(device, "cec2013", capi.algo_desc("sade", gens=20, seed=11 + g, ftol=0.0, xtol=0.0), 1024, seed=200 + g, prob_id=12,
                            dim=50, rotation=mr, shift=os_)
    return f


a = Archipelago(ISLANDS, make(local), topology="ring", seed=1)
a.evolve(1)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
a.evolve(ROUNDS - 1)
torch.cuda.synchronize()
dist.barrier()
dt = time.perf_counter() - t0
mine = np.stack([isl.population().x for isl in a.islands])
gathered = [torch.empty_like(torch.from_numpy(mine).cuda()) for _ in range(world)]
dist.all_gather(gathered, torch.from_numpy(mine).cuda())
if rank == 0:
    multi = np.concatenate([g.cpu().numpy() for g in gathered])
    b = Archipelago(ISLANDS, make(0), topology="ring", seed=1, distributed=False)
    b.evolve(ROUNDS)
    single = np.stack([isl.population().x for isl in b.islands])
    out = {"world": world, "islands": ISLANDS, "rounds": ROUNDS, "identical_to_single_process": bool(np.array_equal(multi, single)),
           "seconds_for_rounds_2_to_4": dt, "evals_per_s": (ROUNDS - 1) * 20 * 1024 * ISLANDS / dt, "migrants_logged_rank0": len(a.log),
           "backend": dist.get_backend()}
    print(json.dumps(out))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"dist_archi_{world}gpu.json").write_text(json.dumps(out, indent=1))
dist.barrier()
dist.destroy_process_group()
