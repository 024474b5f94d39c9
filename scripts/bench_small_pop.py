"""Launch-bound generation loops on the GPU box: cfg1 (rastrigin D=10, pop 1024, de1220, 100 generations per evolve()) and ONE
cfg5 island (cec2013 f12 D=50, pop 1024, sade, 50 generations per evolve()), with the cached generation graph and with plain
launches (PGC_GRAPHS=0)."""
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi, synth  # noqa: E402

L = capi.lib()
ctx = capi.Context(0)
out = {}
for name, family, kw, algo, gens in (("cfg1_rastrigin_d10_de1220", "rastrigin", dict(dim=10), "de1220", 100),
                                     ("cfg5_island_cec2013_f12_d50_sade", "cec2013", dict(prob_id=12, dim=50), "sade", 50)):
    for graphs in ("1", "0"):
        os.environ["PGC_GRAPHS"] = graphs
        if family == "cec2013":
            mr, os_ = synth.cec2013_tables(kw["dim"])
            p = capi.Problem(ctx, family, rotation=mr, shift=os_, **kw)
        else:
            p = capi.Problem(ctx, family, **kw)
        NP = 1024
        d_x, d_f = ctx.malloc(8 * NP * p.nx), ctx.malloc(8 * NP)
        a = capi.algo_desc(algo, gens=gens, seed=41, ftol=0.0, xtol=0.0)
        capi.check(L.pgc_population_init_device(p._h, NP, 23, d_x, d_f, None, None))
        for k in range(2):
            capi.check(L.pgc_algo_evolve_device(p._h, C.byref(a), d_x, d_f, NP, 1 + gens * k, None, None))
        ctx.synchronize()
        reps = 10
        l0 = ctx.launches
        t0 = time.perf_counter()
        for k in range(reps):
            capi.check(L.pgc_algo_evolve_device(p._h, C.byref(a), d_x, d_f, NP, 1 + gens * (k + 2), None, None))
        ctx.synchronize()
        dt = time.perf_counter() - t0
        out[f"{name}_graphs{graphs}"] = {"generations_per_s": gens * reps / dt, "us_per_generation": dt / (gens * reps) * 1e6,
                                         "evals_per_s": gens * reps * NP / dt, "launches_per_generation": (ctx.launches - l0) / (gens * reps)}
        ctx.free(d_x)
        ctx.free(d_f)
        p.close()
print(json.dumps(out, indent=1))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bench_small_pop.json").write_text(json.dumps(out, indent=1))
