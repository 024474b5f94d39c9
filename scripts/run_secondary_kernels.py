"""Launch each secondary kernel a few times on representative sizes (for `ncu -k regex:<name>` captures, scripts/gpu_ncu_secondary.sh).
argv[1] selects one: cec13 | lj | fnds | gram | sample | de | hv"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi  # noqa: E402
from pagmo2_b200 import synth  # noqa: E402  (synthetic data tables)

which = sys.argv[1]
ctx = capi.Context(0)
lib = capi.lib()
g = torch.Generator(device="cuda:0").manual_seed(1)
if which == "cec13":
    mr, os_ = synth.cec2013_tables(50)
    prob = capi.Problem(ctx, "cec2013", prob_id=12, dim=50, rotation=mr, shift=os_)
    n = 1 << 20
    x = torch.rand((n, 50), dtype=torch.float64, device="cuda:0", generator=g) * 200 - 100
    f = torch.empty(n, dtype=torch.float64, device="cuda:0")
    for _ in range(3):
        prob.eval_device(x.data_ptr(), n, f.data_ptr(), ctx.stream)
elif which == "lj":
    prob = capi.Problem(ctx, "lennard_jones", dim=150)
    n = 1 << 16
    lb, ub = prob.bounds()
    x = torch.rand((n, prob.nx), dtype=torch.float64, device="cuda:0", generator=g) * torch.tensor(ub - lb, device="cuda:0") + torch.tensor(lb, device="cuda:0")
    f = torch.empty(n, dtype=torch.float64, device="cuda:0")
    for _ in range(3):
        prob.eval_device(x.data_ptr(), n, f.data_ptr(), ctx.stream)
elif which == "fnds":
    n = 1 << 16
    f = np.random.default_rng(31).uniform(0, 1, (n, 2))
    for _ in range(2):
        ctx.fnds(f)
elif which in ("gram", "sample"):
    D, lam = 100, 65536
    rng = np.random.default_rng(3)
    x = rng.normal(size=(lam, D))
    w = np.full(lam // 2, 2.0 / lam)
    for _ in range(2):
        if which == "gram":
            ctx.weighted_gram(x, w, idx=np.arange(lam // 2), center=np.zeros(D), scale_div=1.0)
        else:
            ctx.cmaes_sample(np.zeros(D), np.eye(D), 1.0, lam, 1, 1)
elif which == "de":
    prob = capi.Problem(ctx, "rastrigin", dim=10)
    n = 1 << 20
    lb, ub = prob.bounds()
    x = np.random.default_rng(5).uniform(lb, ub, (n, 10))
    f = prob.eval_host(x)
    prob.evolve(capi.algo_desc("de1220", gens=3, seed=3, ftol=0.0, xtol=0.0), x, f)
elif which == "hv":
    rng = np.random.default_rng(1)
    f = rng.uniform(0, 1, (8192, 3))
    f = f / np.linalg.norm(f, axis=1, keepdims=True)
    ctx.hv_contributions(f, np.full(3, 1.25))
ctx.synchronize()
print("done", which)
