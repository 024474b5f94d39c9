"""Small-population generation loops under ncu (launch list): cfg1 (rastrigin D=10, pop 1024, de1220) and one cfg5 island
(cec2013 f12 D=50, pop 1024, sade), a few generations each with plain launches.

    PGC_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/small_pop_launches.csv \
        python scripts/small_pop_launches.py
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi, synth  # noqa: E402

ctx = capi.Context(0)
al = np.array([2, 3, 7, 10, 13, 14, 15, 16], dtype=np.uint32)
for family, dim, code, kw in (("rastrigin", 10, 2, {}), ("cec2013", 50, 1, {})):
    if family == "cec2013":
        mr, os_ = synth.cec2013_tables(dim)
        prob = capi.Problem(ctx, "cec2013", prob_id=12, dim=dim, rotation=mr, shift=os_)
    else:
        prob = capi.Problem(ctx, family, dim=dim)
    lb, ub = prob.bounds()
    x = np.random.default_rng(5).uniform(lb, ub, (1024, dim))
    f = prob.eval_host(x)[:, 0]
    dx, df = ctx.to_device(x), ctx.to_device(f)
    done = C.c_uint()
    capi.check(capi.lib().pgc_de_evolve_device(prob._h, dx, df, 1024, 6, code, 2, 1, 0.8, 0.9, al.ctypes.data_as(C.c_void_p), al.size, 0.0, 0.0,
                                               None, None, None, 3, 1, C.byref(done), None))
    ctx.synchronize()
