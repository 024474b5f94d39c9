import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi
ctx = capi.Context(0)
prob = capi.Problem(ctx, "rastrigin", dim=10)
lb, ub = prob.bounds()
n = 1 << 20
x = np.random.default_rng(5).uniform(lb, ub, (n, 10))
f = prob.eval_host(x)
prob.evolve(capi.algo_desc("de1220", gens=6, seed=3, ftol=0.0, xtol=0.0), x, f)
