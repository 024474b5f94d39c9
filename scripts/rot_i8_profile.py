#!/usr/bin/env python
"""One 1 Mi x 100 launch pair of the tcgen05 i8 rotation prototype (for ncu)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi, synth  # noqa: E402

D = 100
ctx = capi.Context(0)
L = capi.lib()
L.pgc_debug_rot_i8.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_size_t,
                               C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_void_p]
mr, os_c, shuf = synth.cec2014_tables(1, D)
M = np.ascontiguousarray(mr[:D * D].reshape(D, D))
os_ = np.ascontiguousarray(os_c[:D])
coef = 10.0 ** (6.0 * np.arange(D) / (D - 1))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
X = torch.rand((n, D), dtype=torch.float64, device="cuda:0") * 200 - 100
F = torch.empty(n, dtype=torch.float64, device="cuda:0")
ms = C.c_float()
capi.check(L.pgc_debug_rot_i8(ctx._h, M.ctypes.data, D, os_.ctypes.data, coef.ctypes.data, 1.0, 100.0, X.data_ptr(), n, F.data_ptr(), 2, C.byref(ms), None))
print("ms per launch", ms.value)
