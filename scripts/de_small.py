import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi
ctx = capi.Context(0)
prob = capi.Problem(ctx, "rastrigin", dim=10)
lb, ub = prob.bounds()
x = np.random.default_rng(5).uniform(lb, ub, (1024, 10))
f = prob.eval_host(x)
g = int(sys.argv[1]) if len(sys.argv) > 1 else 20
prob.evolve(capi.algo_desc("de1220", gens=g, seed=3, ftol=0.0, xtol=0.0), x, f)
t0 = time.perf_counter()
prob.evolve(capi.algo_desc("de1220", gens=g, seed=3, ftol=0.0, xtol=0.0), x, f)
print("per gen us", (time.perf_counter() - t0) / g * 1e6)
