import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from pagmo2_b200 import capi
ctx = capi.Context(0)
rng = np.random.default_rng(0)
p = capi.Problem(ctx, "dtlz", prob_id=2, dim=12, nobj=4, param=100)
lb, ub = p.bounds()
for n in (512, 2048):
    x = rng.uniform(lb, ub, (n, p.nx)); f = p.eval_host(x)
    t = time.time(); xg, fg, st, done = p.maco_evolve(x, f, gens=4, seed=1); ctx.synchronize()
    print("maco 4 objectives", n, done, round(time.time() - t, 3), "s", bool(np.allclose(p.eval_host(xg), fg, rtol=1e-12)), flush=True)
p.close()
p = capi.Problem(ctx, "rastrigin", dim=30)
lb, ub = p.bounds()
n = 262144
x = rng.uniform(lb, ub, (n, 30)); f = p.eval_host(x)
t = time.time(); xg, fg, st, done = p.gaco_evolve(x, f, gens=10, seed=1); ctx.synchronize()
print("gaco", n, done, round(time.time() - t, 3), "s", fg.min() < f.min(), flush=True)
t = time.time(); xg, fg, st, done = p.gaco_evolve(x, f, gens=5, ker=4096, seed=1); ctx.synchronize()
print("gaco ker 4096", done, round(time.time() - t, 3), "s", flush=True)
pts = rng.uniform(0.05, 1, (4000, 4)); pts /= np.linalg.norm(pts, axis=1, keepdims=True)
t = time.time(); hv = ctx.hv_compute(pts, np.full(4, 1.2)); print("wfg 4 x 4000 compute", round(time.time() - t, 3), "s", hv, flush=True)
t = time.time(); c = ctx.hv_contributions(pts[:2500], np.full(4, 1.2)); print("wfg 4 x 2500 contributions", round(time.time() - t, 3), "s", c.min() >= 0, flush=True)
