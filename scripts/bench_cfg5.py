"""BASELINE cfg5 (run on the GPU box): cec2013 D=50, 8 islands x pop 1024 (our choice of island size), ring(1.0), sade
defaults, device archipelago with migration every `GENS` generations; plus the raw cec2013 evaluator on a 1 Mi batch.
CPU beside it: the unmodified reference sade::evolve on ONE island of the same size (the reference runs its islands on one
thread each, so 8 islands on >= 8 cores take the time of one)."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi  # noqa: E402
from pagmo2_b200.archipelago import Archipelago, DeviceIsland  # noqa: E402
from pagmo2_b200 import synth  # noqa: E402  (synthetic data tables)

D, ISLANDS, POP, GENS, ROUNDS = 50, 8, 1024, 50, 4
mr, os_ = synth.cec2013_tables(D)
out = {"config": {"dim": D, "islands": ISLANDS, "pop_per_island": POP, "gens_per_round": GENS, "rounds": ROUNDS, "topology": "ring(1.0)",
                  "algo": "sade defaults, ftol = xtol = 0"}}
for func in (12, 28):
    def dev(g):
        return DeviceIsland(0, "cec2013", capi.algo_desc("sade", gens=GENS, seed=11 + g, ftol=0.0, xtol=0.0), POP, seed=200 + g, prob_id=func, dim=D,
                            rotation=mr, shift=os_)
    a = Archipelago(ISLANDS, dev, topology="ring", seed=1)
    a.evolve(1)
    for isl in a.islands:
        isl.ctx.synchronize()
    l0 = sum(isl.ctx.launches for isl in a.islands)
    t0 = time.perf_counter()
    a.evolve(ROUNDS)
    for isl in a.islands:
        isl.ctx.synchronize()
    dt = time.perf_counter() - t0
    l1 = sum(isl.ctx.launches for isl in a.islands)
    gens = ROUNDS * GENS
    out[f"f{func}"] = {"seconds": dt, "generations_per_s_per_island": gens / dt, "evals_per_s_all_islands": gens * POP * ISLANDS / dt,
                       "launches_per_generation_per_island": (l1 - l0) / gens / ISLANDS, "migrants_logged": len(a.log),
                       "champions_f": a.champions_f().tolist()}
    try:
        from oracle.pyoracle import reference
        rp = reference().problem("cec2013", func, D)
        secs, _, fr, fev = rp.evolve("sade", POP, 20, 200, 11)
        out[f"f{func}"]["cpu_reference_one_island"] = {"generations_per_s": 20 / secs, "evals_per_s": 20 * POP / secs, "cores": 1,
                                                       "what": "unmodified sade::evolve, 20 generations, pop 1024"}
    except Exception as e:  # noqa: BLE001
        out[f"f{func}"]["cpu_reference_one_island"] = {"unavailable": str(e)[:200]}
    del a

# raw evaluator throughput, 1 Mi individuals resident in HBM
ctx = capi.Context(0)
N = 1 << 20
x = torch.rand((N, D), dtype=torch.float64, device="cuda:0") * 200 - 100
f = torch.empty(N, dtype=torch.float64, device="cuda:0")
stream = torch.cuda.ExternalStream(ctx.stream)
ev = {}
for func in range(1, 29):
    prob = capi.Problem(ctx, "cec2013", prob_id=func, dim=D, rotation=mr, shift=os_)
    for _ in range(2):
        prob.eval_device(x.data_ptr(), N, f.data_ptr(), ctx.stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        prob.eval_device(x.data_ptr(), N, f.data_ptr(), ctx.stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ev[func] = N / (e0.elapsed_time(e1) / 3 * 1e-3)
    prob.close()
out["eval_1Mi_D50_evals_per_s"] = ev
out["eval_1Mi_D50_geomean"] = float(np.exp(np.mean(np.log(list(ev.values())))))
try:
    from oracle.pyoracle import reference
    import os
    R = reference()
    cores = os.cpu_count() or 1
    xs = x[: 512 * cores].cpu().numpy()
    cpu = {}
    for func in (1, 12, 28):
        rp = R.problem("cec2013", func, D)
        t0 = time.perf_counter()
        rp.thread_bfe(xs, nthreads=cores)
        cpu[func] = xs.shape[0] / (time.perf_counter() - t0)
    out["cpu_reference_thread_bfe_evals_per_s"] = {"cores": cores, **cpu}
except Exception as e:  # noqa: BLE001
    out["cpu_reference_thread_bfe_evals_per_s"] = {"unavailable": str(e)[:200]}
print(json.dumps(out, indent=1))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bench_cfg5.json").write_text(json.dumps(out, indent=1))
