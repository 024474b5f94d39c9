#!/usr/bin/env python
"""cycles per tcgen05.mma kind::i8 instruction (M = 128, K = 32) on the B200: N x issue pattern (pgc_debug_mma_i8_probe)."""
import ctypes as C, json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi
ctx = capi.Context(0)
out = np.zeros(32)
capi.lib().pgc_debug_mma_i8_probe.argtypes = [C.c_void_p, C.c_void_p]
capi.check(capi.lib().pgc_debug_mma_i8_probe(ctx._h, out.ctypes.data))
out = out.reshape(4, 4, 2)
pat = ["one accumulator", "7 accumulators in turn", "7 accumulators + collector::a reuse", "fresh A and B tiles"]
res = {}
for ni, N in enumerate((32, 64, 128, 256)):
    for p in range(4):
        issue, done = out[ni, p]
        macs = 128 * N * 32 / done
        res[f"N={N} / {pat[p]}"] = {"cycles_per_mma_issue": issue, "cycles_per_mma_complete": done, "mac_per_cycle": macs}
        print(f"N={N:3d} {pat[p]:38s} issue {issue:7.1f}  complete {done:7.1f} cyc/MMA  -> {macs:7.0f} MAC/cycle/SM (peak 8192)")
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "mma_i8_probe.json").write_text(json.dumps(res, indent=1))
