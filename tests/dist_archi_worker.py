"""Worker of tests/test_archipelago.py: one process of a torch.distributed (gloo) archipelago of oracle-backed islands.
usage: torchrun ... dist_archi_worker.py OUT.npz N_ISLANDS TOPOLOGY MIGRATION_TYPE HANDLING ROUNDS"""
import sys
from pathlib import Path

import numpy as np
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle.pyoracle import oracle  # noqa: E402
from oracle_island import OracleIsland  # noqa: E402
from pagmo2_b200.archipelago import Archipelago  # noqa: E402


def build(n_islands, topology, mtype, handling):
    orc = oracle()
    return Archipelago(n_islands, lambda g: OracleIsland(orc, "rastrigin", 6, 16, seed=100 + g, algo="sade", gens=2, algo_seed=7 + g, s_rate=2,
                                                         r_rate=2, ftol=0.0, xtol=0.0),
                       topology=topology, weight=0.75, migration_type=mtype, migrant_handling=handling, seed=5)


def main():
    out, n_islands, topology, mtype, handling, rounds = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5], int(sys.argv[6])
    dist.init_process_group("gloo")
    a = build(n_islands, topology, mtype, handling)
    a.evolve(rounds)
    pops = [isl.population() for isl in a.islands]
    np.savez(f"{out}.rank{a.rank}.npz", first=a.first, x=np.stack([p.x for p in pops]), f=np.stack([p.f for p in pops]),
             ids=np.stack([p.ids for p in pops]), log=np.array([(e.round, e.id % (1 << 62), e.src, e.dst) for e in a.log], dtype=np.int64).reshape(-1, 4))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
