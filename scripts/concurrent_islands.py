"""How well do K evolve() loops overlap on ONE GPU?  K threads, each with its own context (stream), problem and population
(cec2013 f12 D=50, pop 1024, sade, 50 generations per call), no migration.  Prints island-generations/s for K = 1, 2, 4, 8."""
import ctypes as C
import json
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi, synth  # noqa: E402

L = capi.lib()
mr, os_ = synth.cec2013_tables(50)
NP, GENS, REPS = 1024, 50, 8
out = {}
for K in (1, 2, 4, 8):
    isl = []
    for g in range(K):
        ctx = capi.Context(0)
        p = capi.Problem(ctx, "cec2013", prob_id=12, dim=50, rotation=mr, shift=os_)
        d_x, d_f = ctx.malloc(8 * NP * 50), ctx.malloc(8 * NP)
        capi.check(L.pgc_population_init_device(p._h, NP, 23 + g, d_x, d_f, None, None))
        isl.append((ctx, p, d_x, d_f, capi.algo_desc("sade", gens=GENS, seed=41 + g, ftol=0.0, xtol=0.0)))

    def run(i, first, reps):
        ctx, p, d_x, d_f, a = isl[i]
        for k in range(reps):
            capi.check(L.pgc_algo_evolve_device(p._h, C.byref(a), d_x, d_f, NP, first + GENS * k, None, None))
        ctx.synchronize()

    with ThreadPoolExecutor(K) as pool:
        list(pool.map(lambda i: run(i, 1, 2), range(K)))
        t0 = time.perf_counter()
        list(pool.map(lambda i: run(i, 1 + 2 * GENS, REPS), range(K)))
        dt = time.perf_counter() - t0
    out[K] = {"island_generations_per_s": K * GENS * REPS / dt, "us_per_generation_per_island": dt / (GENS * REPS) * 1e6}
    for ctx, p, d_x, d_f, a in isl:
        ctx.free(d_x)
        ctx.free(d_f)
        p.close()
print(json.dumps(out, indent=1))
