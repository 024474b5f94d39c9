"""Which part of an island-generation stops overlapping when K islands share ONE GPU?  K threads, one context (stream) each:
  (a) eval only: cec2013 f12 D=50 on 1024 rows, back to back on each stream (the island-sized cec13 launch);
  (b) the generation loop with a cheap evaluator: sade on rastrigin D=50, pop 1024, launch-per-phase path (PGC_DE_RESIDENT=0): trial + finish;
  (c) the full cfg5 island loop (sade on cec2013 f12 D=50).
Prints microseconds per (island, generation or evaluation) for K = 1, 2, 4, 8."""
import ctypes as C
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

os.environ.setdefault("PGC_DE_RESIDENT", "0")
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi, synth  # noqa: E402

L = capi.lib()
mr, os_ = synth.cec2013_tables(50)
NP, GENS, REPS = 1024, 50, 8
out = {}
for mode in os.environ.get("PARTS_MODES", "eval_cec2013,loop_rastrigin,loop_cec2013").split(","):
    out[mode] = {}
    for K in [int(k) for k in os.environ.get("PARTS_KS", "1,2,4,8").split(",")]:
        isl = []
        for g in range(K):
            ctx = capi.Context(0)
            L.pgc_ctx_set_sharers.argtypes = [C.c_void_p, C.c_int]
            capi.check(L.pgc_ctx_set_sharers(ctx._h, K))
            if mode == "loop_rastrigin":  # (modes island_*: cec2013 f12 through capi.Island)
                p = capi.Problem(ctx, "rastrigin", dim=50)
            else:
                p = capi.Problem(ctx, "cec2013", prob_id=12, dim=50, rotation=mr, shift=os_)
            d_x, d_f = ctx.malloc(8 * NP * 50), ctx.malloc(8 * NP)
            capi.check(L.pgc_population_init_device(p._h, NP, 23 + g, d_x, d_f, None, None))
            isl.append((ctx, p, d_x, d_f, capi.algo_desc("sade", gens=GENS, seed=41 + g, ftol=0.0, xtol=0.0)))
            if mode.startswith("island"):  # the resident island object: (select | replace + select) around every evolve call, no migrants
                I = capi.Island(p, NP, 1, 1)
                I.init(200 + g)
                isl[-1] = isl[-1] + (I,)

        def run(i, first, reps):
            ctx, p, d_x, d_f, a = isl[i][:5]
            for k in range(reps):
                if mode.startswith("island"):
                    I = isl[i][5]
                    if mode == "island_replace_select":
                        I.replace_enqueue(1, [0], log=False)
                    I.evolve(a)
                    if mode != "island_evolve":
                        I.select(1)
                elif mode == "eval_cec2013":
                    for _ in range(GENS):
                        p.eval_device(d_x, NP, d_f, ctx.stream)
                else:
                    capi.check(L.pgc_algo_evolve_device(p._h, C.byref(a), d_x, d_f, NP, first + GENS * k, None, None))
            ctx.synchronize()

        with ThreadPoolExecutor(K) as pool:
            list(pool.map(lambda i: run(i, 1, 2), range(K)))
            t0 = time.perf_counter()
            list(pool.map(lambda i: run(i, 1 + 2 * GENS, REPS), range(K)))
            dt = time.perf_counter() - t0
        out[mode][K] = {"us_per_island_step": dt / (GENS * REPS) * 1e6, "steps_per_s_all_islands": K * GENS * REPS / dt}
        for ctx, p, d_x, d_f, a, *rest in isl:
            for I in rest:
                I.close()
            ctx.free(d_x)
            ctx.free(d_f)
            p.close()
print(json.dumps(out, indent=1))
