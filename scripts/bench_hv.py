"""Hypervolume micro-benchmark (run on the GPU box): exclusive contributions and the indicator for random non-dominated fronts,
device (points resident in HBM) against the unmodified reference (hv2d / HyCon3D) on one host core."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi  # noqa: E402

ctx = capi.Context(0)
lib = capi.lib()
stream = torch.cuda.ExternalStream(ctx.stream)
out = {}
rng = np.random.default_rng(1)
only_wfg = "--wfg" in sys.argv
for m in (4, 5) if only_wfg else (2, 3, 4, 5):
    for n in ((128, 512, 1024) if m >= 4 else (1024, 8192, 65536)):  # m >= 4: the device WFG against hvwfg
        f = rng.uniform(0, 1, (n, m))
        f = f / np.linalg.norm(f, axis=1, keepdims=True)
        r = np.full(m, 1.25)
        d = torch.from_numpy(f).cuda()
        o = torch.empty(n, dtype=torch.float64, device="cuda:0")
        rp = r.ctypes.data_as(capi.C.POINTER(capi.C.c_double))
        res = {}
        for name, mode in (("contributions", 0), ("compute", 1)):
            for _ in range(2):
                capi.check(lib.pgc_hv_device(ctx._h, d.data_ptr(), n, m, rp, mode, o.data_ptr(), None))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3):
                capi.check(lib.pgc_hv_device(ctx._h, d.data_ptr(), n, m, rp, mode, o.data_ptr(), None))
            e1.record(stream)
            torch.cuda.synchronize()
            res[name + "_ms"] = e0.elapsed_time(e1) / 3
            print(m, n, name, res[name + "_ms"], flush=True)
        try:
            from oracle.pyoracle import reference
            R = reference()
            t0 = time.perf_counter(); R.hv_contributions(f, r); res["cpu_reference_contributions_ms"] = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter(); R.hv_compute(f, r); res["cpu_reference_compute_ms"] = (time.perf_counter() - t0) * 1e3
        except Exception as e:  # noqa: BLE001
            res["cpu_reference"] = str(e)[:100]
        out[f"m{m}_n{n}"] = res
print(json.dumps(out, indent=1))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / ("bench_hv_wfg.json" if only_wfg else "bench_hv.json")).write_text(json.dumps(out, indent=1))
