"""FP64 pipe probes (run on the GPU box): DMMA vs DFMA, alone and mixed, by warps per CTA (1 CTA per SM)."""
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi  # noqa: E402

ctx = capi.Context(0)
L = capi.lib()
L.pgc_debug_fp64_mix_probe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
res = []
for total, dm in [(4, 4), (8, 8), (16, 16), (4, 0), (8, 0), (16, 0), (8, 4), (16, 8), (16, 4), (12, 4), (12, 8)]:
    out = (C.c_double * 2)()
    capi.check(L.pgc_debug_fp64_mix_probe(ctx._h, 4000, total, dm, out))
    res.append({"warps_per_sm": total, "dmma_warps": dm, "dmma_tflops": out[0], "dfma_tflops": out[1], "sum": out[0] + out[1]})
    print(res[-1], flush=True)
print("dfma peak probe", ctx.fp64_peak_tflops(4096), "dmma peak probe", ctx.fp64_mma_peak_tflops(4096))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "probe_fp64.json").write_text(json.dumps(res, indent=1))
