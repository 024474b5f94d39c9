"""NSGA-II ranking path timings at BASELINE cfg3 scale (run on the GPU box): ZDT1 (nx=30) / DTLZ2 (M=3, nx=12) evaluation,
fast_non_dominated_sorting at N and 2N, crowding of all fronts, select_best_N_mo(2N -> N).  Device-resident, CUDA events."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ctx = capi.Context(0)
L = capi.lib()
stream = torch.cuda.ExternalStream(ctx.stream)
dev = "cuda:0"
out = {}


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append((e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    return min(t[0] for t in ts), min(t[1] for t in ts)


for name, fam, kw, nx, m in (("zdt1", "zdt", dict(prob_id=1, dim=30), 30, 2), ("dtlz2", "dtlz", dict(prob_id=2, dim=12, nobj=3), 12, 3)):
    prob = capi.Problem(ctx, fam, **kw)
    g = torch.Generator(device=dev).manual_seed(31)
    x = torch.rand((2 * N, nx), dtype=torch.float64, device=dev, generator=g)
    f = torch.empty((2 * N, m), dtype=torch.float64, device=dev)
    ev = timed(lambda: prob.eval_device(x.data_ptr(), 2 * N, f.data_ptr(), ctx.stream))
    torch.cuda.synchronize()
    rank = torch.empty(2 * N, dtype=torch.int32, device=dev)
    dc = torch.empty(2 * N, dtype=torch.int32, device=dev)
    order = torch.empty(2 * N, dtype=torch.int32, device=dev)
    foff = torch.empty(2 * N + 1, dtype=torch.int32, device=dev)
    cd = torch.empty(2 * N, dtype=torch.float64, device=dev)
    sel = torch.empty(2 * N, dtype=torch.int32, device=dev)
    nf = C.c_uint32()
    res = {"eval_2N_ms": ev[0], "eval_evals_per_s": 2 * N / (ev[0] * 1e-3)}
    for tag, n in (("N", N), ("2N", 2 * N)):
        t = timed(lambda: capi.check(L.pgc_fnds_device(ctx._h, f.data_ptr(), n, m, rank.data_ptr(), dc.data_ptr(), order.data_ptr(),
                                                       foff.data_ptr(), C.byref(nf), None)))
        res[f"fnds_{tag}_ms"] = t[1]
        res[f"fnds_{tag}_fronts"] = nf.value
        t = timed(lambda: capi.check(L.pgc_crowding_fronts_device(ctx._h, f.data_ptr(), n, m, order.data_ptr(), foff.data_ptr(), nf.value, 1,
                                                                  cd.data_ptr(), None)))
        res[f"crowding_{tag}_ms"] = t[1]
    no = C.c_uint32()
    t = timed(lambda: capi.check(L.pgc_select_best_N_mo_device(ctx._h, f.data_ptr(), 2 * N, m, N, sel.data_ptr(), C.byref(no), None)))
    res["select_best_2N_to_N_ms"] = t[1]
    res["ranking_ms_per_generation"] = res["fnds_N_ms"] + res["crowding_N_ms"] + res["select_best_2N_to_N_ms"]
    res["generations_per_s_ranking_only"] = 1e3 / res["ranking_ms_per_generation"]
    out[name] = res
    print(name, json.dumps(res), flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"bench_mo_{N}.json").write_text(json.dumps(out, indent=1))
