"""Measurements for the SURVEY 8(a) rows that had none (run on the GPU box): WFG evaluator throughput (a8), sga generations/s (a21),
population construction = batch_random_decision_vector + batch evaluation (a24), and the meta-problems (8f row 1).
Device-resident, wall clock around synchronised calls after one warm-up."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi  # noqa: E402

ctx = capi.Context(0)
L = capi.lib()
out = {}


def timed(fn, reps=5):
    fn()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.synchronize()
    return (time.perf_counter() - t0) / reps


n = 1 << 20
rng = np.random.default_rng(3)
# a8: WFG1..9, 24 decision variables, 3 objectives, k = 4 (the shapes of tests/wfg.cpp scaled up)
for pid in (1, 4, 9):
    p = capi.Problem(ctx, "wfg", prob_id=pid, dim=24, nobj=3, param=4)
    lb, ub = p.bounds()
    dx, df = ctx.to_device(rng.uniform(lb, ub, (n, 24))), ctx.malloc(8 * n * 3)
    dt = timed(lambda: p.eval_device(dx, n, df))
    out[f"wfg{pid}_nx24_m3"] = {"n": n, "ms": dt * 1e3, "evals_per_s": n / dt, "hbm_gbs": n * 8 * 27 / dt / 1e9}
    ctx.free(dx); ctx.free(df); p.close()

# a24: population(prob, bfe, n, seed) on the device: random decision vectors + one batch evaluation + ids
p = capi.Problem(ctx, "rastrigin", dim=30)
dx, df, di = ctx.malloc(8 * n * 30), ctx.malloc(8 * n), ctx.malloc(8 * n)
dt = timed(lambda: capi.check(L.pgc_population_init_device(p._h, n, 42, dx, df, di, None)))
out["population_init_rastrigin_D30"] = {"n": n, "ms": dt * 1e3, "individuals_per_s": n / dt}

# a21: sga (reference defaults: exponential crossover, polynomial mutation, tournament selection) at two population sizes
for NP in (1024, 1 << 16):
    lb, ub = p.bounds()
    x = rng.uniform(lb, ub, (NP, 30))
    f = p.eval_host(x)
    ddx, ddf = ctx.to_device(x), ctx.to_device(f)
    gens = 50
    algo = capi.algo_desc("sga", gens=gens, seed=7)
    done = C.c_uint()

    def run(first):
        capi.check(L.pgc_algo_evolve_device(p._h, C.byref(algo), ddx, ddf, NP, first, C.byref(done), None))
    dt = timed(lambda: run(1), reps=3)
    out[f"sga_rastrigin_D30_pop{NP}"] = {"generations": gens, "ms_per_generation": dt / gens * 1e3, "generations_per_s": gens / dt,
                                        "evals_per_s": gens * NP / dt}
    ctx.free(ddx); ctx.free(ddf)

# 8f row 1: translate{rastrigin D=30} and decompose{dtlz2 nx=12, m=3} against their inner problems
t = rng.uniform(-1, 1, 30)
pt = p.translate(t)
xs = ctx.to_device(rng.uniform(-4, 4, (n, 30)))
a = timed(lambda: p.eval_device(xs, n, df))
b = timed(lambda: pt.eval_device(xs, n, df))
out["translate_rastrigin_D30"] = {"n": n, "inner_ms": a * 1e3, "translated_ms": b * 1e3}
pd3 = capi.Problem(ctx, "dtlz", prob_id=2, dim=12, nobj=3, param=100)
pdd = pd3.decompose([0.2, 0.3, 0.5], [0.0, 0.0, 0.0], "tchebycheff")
xd, fd = ctx.to_device(rng.uniform(0, 1, (n, 12))), ctx.malloc(8 * n * 3)
a = timed(lambda: pd3.eval_device(xd, n, fd))
b = timed(lambda: pdd.eval_device(xd, n, df))
out["decompose_dtlz2_nx12_m3"] = {"n": n, "inner_ms": a * 1e3, "decomposed_ms": b * 1e3}
print(json.dumps(out, indent=1))
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bench_misc.json").write_text(json.dumps(out, indent=1))
