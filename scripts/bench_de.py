"""BASELINE cfg1 (run on the GPU box): rastrigin D=10, population 1024, de1220, 100 generations.  Whole evolve() on the device
(trial construction + batch fitness + selection per generation) against the unmodified reference de1220::evolve on one host
core (the reference's DE is sequential).  Also a large-population point where the device is not launch-bound."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi  # noqa: E402

ctx = capi.Context(0)
out = {}
for NP, dim, gens in ((1024, 10, 100), (1 << 20, 10, 100), (1 << 18, 100, 50)):
    prob = capi.Problem(ctx, "rastrigin", dim=dim)
    lb, ub = prob.bounds()
    rng = np.random.default_rng(5)
    x = rng.uniform(lb, ub, (NP, dim))
    f = prob.eval_host(x)[:, 0]
    dx, df = ctx.to_device(x), ctx.to_device(f)
    al = np.array([2, 3, 7, 10, 13, 14, 15, 16], dtype=np.uint32)
    import ctypes as C

    def run(g, first):
        done = C.c_uint()
        capi.check(capi.lib().pgc_de_evolve_device(prob._h, dx, df, NP, g, 2, 2, 1, 0.8, 0.9, al.ctypes.data_as(C.c_void_p), al.size, 0.0, 0.0,
                                                   None, None, None, 3, first, C.byref(done), None))
        ctx.synchronize()
        return done.value

    run(3, 1)
    t0 = time.perf_counter()
    done = run(gens, 4)
    dt = time.perf_counter() - t0
    fb = ctx.from_device(df, f.shape)
    key = f"rastrigin_D{dim}_pop{NP}"
    out[key] = {"algo": "de1220", "generations": done, "seconds": dt, "generations_per_s": done / dt, "evals_per_s": done * NP / dt,
                "best_f_before": float(f.min()), "best_f_after": float(fb.min())}
    ctx.free(dx)
    ctx.free(df)
    prob.close()
    if NP == 1024:
        try:
            from oracle.pyoracle import reference
            rp = reference().problem("rastrigin", dim)
            secs, _, fr, fev = rp.evolve("de1220", NP, gens, 5, 3)
            out[key]["cpu_reference"] = {"seconds": secs, "evals_per_s": NP * gens / secs, "cores": 1, "best_f_after": float(np.min(fr)),
                                         "what": "unmodified de1220::evolve, sequential problem::fitness"}
        except Exception as e:  # noqa: BLE001
            out[key]["cpu_reference"] = {"unavailable": str(e)[:200]}
print(json.dumps(out, indent=1))
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bench_de.json").write_text(json.dumps(out, indent=1))
