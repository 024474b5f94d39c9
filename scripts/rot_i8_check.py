#!/usr/bin/env python
"""GPU check of the tensor-core (tcgen05.mma kind::i8, Ozaki digit planes) rotation prototype, pgc_debug_rot_i8:
 1. z = M (x - os) entry by entry against an extended-precision product (numpy longdouble), small and ragged batches;
 2. f = sum_j c_j z_j^2 + bias (the f1 / ellipsoid epilogue) against the oracle-free formula on the exact z;
 3. time per launch at the headline size (1 Mi x 100) next to the DMMA stage kernel (cec2014 f1)."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

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
rng = np.random.default_rng(0)
out = {}


def rot(x, coef=None, rate=1.0, bias=0.0):
    n = x.shape[0]
    dx = ctx.to_device(np.ascontiguousarray(x))
    do = ctx.malloc(8 * n * (D if coef is None else 1))
    capi.check(L.pgc_debug_rot_i8(ctx._h, M.ctypes.data, D, os_.ctypes.data, coef.ctypes.data if coef is not None else None, rate, bias, dx, n,
                                  do, 1, None, None))
    ctx.synchronize()
    r = ctx.from_device(do, (n, D) if coef is None else (n,))
    ctx.free(dx)
    ctx.free(do)
    return r


for name, x in (("n=1", rng.uniform(-100, 100, (1, D))), ("n=33", rng.uniform(-100, 100, (33, D))), ("n=4099", rng.uniform(-100, 100, (4099, D))),
                ("near optimum", os_ + rng.normal(0, 1, (257, D))), ("optimum", np.tile(os_, (40, 1))), ("tiny", os_ + rng.normal(0, 1e-9, (64, D)))):
    z = rot(x)
    y = (x - os_).astype(np.longdouble)
    exact = np.array(y @ M.T.astype(np.longdouble), dtype=np.float64)
    scale = np.abs(exact).max(axis=1, keepdims=True)
    scale[scale == 0] = 1.0
    err = float((np.abs(z - exact) / scale).max())
    fp64 = float((np.abs((x - os_) @ M.T - exact) / scale).max())
    out[name] = {"max_err_over_rowmax": err, "plain_fp64_matmul_err": fp64}
    print(name, "i8 err %.3e  (plain FP64 matmul %.3e)" % (err, fp64), flush=True)
    assert err < 5e-15, (name, err)
# f1-type epilogue
coef = 10.0 ** (6.0 * np.arange(D) / (D - 1))
x = rng.uniform(-100, 100, (5000, D))
f = rot(x, coef=coef, bias=100.0)
y = (x - os_).astype(np.longdouble)
zx = y @ M.T.astype(np.longdouble)
fx = np.array((coef * zx * zx).sum(axis=1) + 100.0, dtype=np.float64)
rel = float(np.abs(f / fx - 1).max())
print("ellipsoid epilogue rel err %.3e" % rel, flush=True)
out["ellipsoid_rel_err"] = rel
assert rel < 1e-13
# against the product's own f1 (DMMA stage kernel)
p1 = capi.Problem(ctx, "cec2014", prob_id=1, dim=D, rotation=mr, shift=os_c, shuffle=shuf)
f_dmma = p1.eval_host(x)[:, 0]
print("vs DMMA stage kernel f1: %.3e" % float(np.abs(f / f_dmma - 1).max()), flush=True)
# timing at the headline size
import torch  # noqa: E402
n = 1 << 20
X = torch.rand((n, D), dtype=torch.float64, device="cuda:0") * 200 - 100
F = torch.empty(n, dtype=torch.float64, device="cuda:0")
stream = torch.cuda.ExternalStream(ctx.stream)


def timed(fn, reps=5):
    fn()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    ctx.synchronize()
    return e0.elapsed_time(e1) / reps


ms = C.c_float()
capi.check(L.pgc_debug_rot_i8(ctx._h, M.ctypes.data, D, os_.ctypes.data, coef.ctypes.data, 1.0, 100.0, X.data_ptr(), n, F.data_ptr(), 10, C.byref(ms),
                              None))
t_i8 = ms.value
t_dmma = timed(lambda: p1.eval_device(X.data_ptr(), n, F.data_ptr(), ctx.stream))
out["ms_per_launch_1Mi"] = {"tcgen05_i8_kernel": t_i8, "dmma_stage_kernel_f1": t_dmma}
print("1 Mi x 100: tcgen05 i8 kernel %.3f ms, DMMA f1 %.3f ms" % (t_i8, t_dmma), flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "rot_i8_check.json").write_text(json.dumps(out, indent=1))
