"""Evaluate chosen CEC2014 functions on a resident batch a few times (target for ncu captures / quick timings).

    python scripts/run_cec14.py 8 10 [--n 1048576] [--reps 3]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("funcs", type=int, nargs="+")
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--dim", type=int, default=100)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
ctx = capi.Context(0)
rng = np.random.default_rng(1)
d_x = ctx.to_device(rng.uniform(-100, 100, (a.n, a.dim)))
d_f = ctx.malloc(8 * a.n)
for f in a.funcs:
    mr, os_c, s = synth.cec2014_tables(f, a.dim)
    p = capi.Problem(ctx, "cec2014", prob_id=f, dim=a.dim, rotation=mr, shift=os_c, shuffle=s)
    p.eval_device(d_x, a.n, d_f)
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.reps):
        p.eval_device(d_x, a.n, d_f)
    ctx.synchronize()
    print(f"f{f}: {(time.perf_counter() - t0) / a.reps * 1e3:.3f} ms per pass", flush=True)
    p.close()
