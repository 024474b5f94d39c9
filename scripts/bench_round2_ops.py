"""The components added late in round 2, device against the unmodified reference on one host core (both on the same box):
gaco / maco / xnes generations per second, the fully informed swarm, the constrained population sort, and the Bringmann-Friedrich
approximations.  Reference runs are kept to a few seconds each (smaller generation counts, the same populations)."""
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
try:
    from oracle.pyoracle import reference
    R = reference()
except Exception as e:  # noqa: BLE001
    R = None
    print("reference unavailable:", e)
out = {}
rng = np.random.default_rng(5)


def timed(fn, reps=1):
    fn()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.synchronize()
    return (time.perf_counter() - t0) / reps


# ---- gaco: rastrigin D = 30, reference defaults (ker 63) ------------------------------------------------------------------------
prob = capi.Problem(ctx, "rastrigin", dim=30)
lb, ub = prob.bounds()
for n in (1024, 65536):
    x = rng.uniform(lb, ub, (n, 30))
    f = prob.eval_host(x)
    gens = 50
    dt = timed(lambda: prob.gaco_evolve(x, f, gens=gens, seed=1))
    row = {"device_gens_per_s": gens / dt, "device_evals_per_s": gens * n / dt}
    if R is not None and n <= 65536:
        rp = R.problem("rastrigin", 30)
        g = 20 if n == 1024 else 2
        t0 = time.perf_counter()
        R.evolve_from(rp, "gaco", [63, 1.0, 0.0, 0.01, 1, 7, 100000, 100000, 0.0], x, g, 1)
        row["reference_gens_per_s"] = g / (time.perf_counter() - t0)
    out[f"gaco_rastrigin30_pop{n}"] = row
    print(n, row, flush=True)
prob.close()

# ---- maco: ZDT1 nx = 30 and DTLZ2 m = 3 ----------------------------------------------------------------------------------------
for name, kw, rargs in (("zdt1_nx30", dict(prob_id=1, dim=30), ("zdt", 1, 30)), ("dtlz2_nx12_m3", dict(prob_id=2, dim=12, nobj=3, param=100), ("dtlz", 2, 12, 3, 100))):
    prob = capi.Problem(ctx, "zdt" if name.startswith("zdt") else "dtlz", **kw)
    lb, ub = prob.bounds()
    for n in (1024, 16384):
        x = rng.uniform(lb, ub, (n, prob.nx))
        f = prob.eval_host(x)
        gens = 10
        dt = timed(lambda: prob.maco_evolve(x, f, gens=gens, seed=1))
        row = {"device_gens_per_s": gens / dt}
        if R is not None:
            rp = R.problem(*rargs)
            g = 3 if n == 1024 else 1
            t0 = time.perf_counter()
            if n <= 1024 or name.startswith("zdt"):
                R.evolve_from(rp, "maco", [63, 1.0, 1, 7, 100000, 0.0], x, g + 1, 1)  # generation 1 only sorts the population
                row["reference_gens_per_s"] = (g + 1) / (time.perf_counter() - t0)
        out[f"maco_{name}_pop{n}"] = row
        print(name, n, row, flush=True)
    prob.close()

# ---- xnes: rosenbrock D = 30 ----------------------------------------------------------------------------------------------------
prob = capi.Problem(ctx, "rosenbrock", dim=30)
lb, ub = prob.bounds()
for n in (64, 4096):
    x = rng.uniform(lb, ub, (n, 30))
    f = prob.eval_host(x)[:, 0]
    gens = 50
    dt = timed(lambda: prob.xnes_evolve(x, f, gens=gens, ftol=0.0, xtol=0.0, seed=1))
    out[f"xnes_rosenbrock30_pop{n}"] = {"device_gens_per_s": gens / dt}
    print("xnes", n, out[f"xnes_rosenbrock30_pop{n}"], flush=True)
# ---- the fully informed swarm (pso_gen variant 6) --------------------------------------------------------------------------------
for n, ntype in ((4096, 2), (4096, 3), (1024, 4)):
    x = rng.uniform(lb, ub, (n, 30))
    f = prob.eval_host(x)[:, 0]
    gens = 50
    dt = timed(lambda: prob.pso_evolve(x, f, gens=gens, variant=6, neighb_type=ntype, seed=1))
    row = {"device_gens_per_s": gens / dt}
    if R is not None:
        rp = R.problem("rosenbrock", 30)
        t0 = time.perf_counter()
        R.evolve_from(rp, "pso_gen", [0.7298, 2.05, 2.05, 0.5, 6, ntype, 4], x, 10, 1)
        row["reference_gens_per_s"] = 10 / (time.perf_counter() - t0)
    out[f"fips_rosenbrock30_pop{n}_topology{ntype}"] = row
    print("fips", n, ntype, row, flush=True)
prob.close()

# ---- sort_population_con at 1 Mi individuals (2 equality + 3 inequality constraints) ---------------------------------------------
L = capi.lib()
vp, sz = C.c_void_p, C.c_size_t
L.pgc_sort_population_con_device.argtypes = [vp, vp, sz, sz, sz, vp, vp, vp]
n = 1 << 20
fc = rng.normal(size=(n, 6))
tol = np.full(5, 1e-2)
d_f, d_o = ctx.to_device(fc), ctx.malloc(4 * n)
dt = timed(lambda: capi.check(L.pgc_sort_population_con_device(ctx._h, d_f, n, 2, 3, tol.ctypes.data, d_o, None)), reps=5)
out["sort_population_con_1Mi"] = {"device_ms": dt * 1e3}
print("sort_population_con", out["sort_population_con_1Mi"], flush=True)

# ---- bf_fpras / bf_approx ---------------------------------------------------------------------------------------------------------
for m, n in ((3, 200), (5, 100), (8, 100)):
    f = rng.uniform(0.05, 1, (n, m))
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    r = np.full(m, 1.2)
    row = {}
    row["fpras_device_ms"] = timed(lambda: ctx.hv_fpras(f, r, 0.01, 0.01, 1)) * 1e3
    row["fpras_device"] = ctx.hv_fpras(f, r, 0.01, 0.01, 1)
    if m <= 5:
        row["exact"] = ctx.hv_compute(f, r)
    if R is not None:
        t0 = time.perf_counter()
        row["fpras_reference"] = R.hv_fpras(f, r, 0.01, 0.01, 1)
        row["fpras_reference_ms"] = (time.perf_counter() - t0) * 1e3
    if m <= 5:
        row["approx_least_device_ms"] = timed(lambda: ctx.hv_approx_extreme(f, r, False, True, seed=1)) * 1e3
        if R is not None and m <= 3:
            t0 = time.perf_counter()
            R.hv_approx_extreme(f, r, False, True, seed=1)
            row["approx_least_reference_ms"] = (time.perf_counter() - t0) * 1e3
    out[f"hv_approx_m{m}_n{n}"] = row
    print(m, n, row, flush=True)

print(json.dumps(out, indent=1))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bench_round2_ops.json").write_text(json.dumps(out, indent=1))
