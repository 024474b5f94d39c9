"""nspso on the device at the NSGA-II headline size (swarm 65 536, ZDT1 nx=30 and DTLZ2 m=3): generations/s for the three diversity
mechanisms (max min is O(N^2 M) per generation in the reference too and is run at 8192), with the restated loop timed at a small
swarm beside it for scale."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi  # noqa: E402

ctx = capi.Context(0)
L = capi.lib()
import ctypes as C  # noqa: E402

L.pgc_nspso_evolve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_double,
                                      C.c_double, C.c_uint, C.c_uint, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
out = {}
for name, kw in (("zdt1_nx30", dict(family="zdt", prob_id=1, dim=30)), ("dtlz2_nx12_m3", dict(family="dtlz", prob_id=2, dim=12, nobj=3, param=100))):
    for div, NP in ((0, 65536), (1, 65536), (2, 8192)):
        p = capi.Problem(ctx, **kw)
        d_x, d_f = ctx.malloc(8 * NP * p.nx), ctx.malloc(8 * NP * p.nf)
        capi.check(L.pgc_population_init_device(p._h, NP, 31, d_x, d_f, None, None))
        gens = 5
        args = (0.6, 2.0, 2.0, 1.0, 0.5, 60, div, 7)
        capi.check(L.pgc_nspso_evolve_device(p._h, d_x, d_f, NP, 2, *args, 1, None, None, None, None))
        ctx.synchronize()
        t0 = time.perf_counter()
        capi.check(L.pgc_nspso_evolve_device(p._h, d_x, d_f, NP, gens, *args, 3, None, None, None, None))
        ctx.synchronize()
        dt = time.perf_counter() - t0
        out[f"{name}_div{div}_pop{NP}"] = {"generations_per_s": gens / dt, "ms_per_generation": dt / gens * 1e3}
        for d in (d_x, d_f):
            ctx.free(d)
        p.close()
print(json.dumps(out, indent=1))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bench_nspso.json").write_text(json.dumps(out, indent=1))
