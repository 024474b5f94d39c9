"""BASELINE cfg4 (run on the GPU box): Lennard-Jones 150 atoms (D=444), PSO swarm of 262144: evals/s of the pair-energy kernel
and generations/s of the whole pso_gen generation on the device; reference lennard_jones::fitness on the host cores beside it."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
ATOMS = 150
ctx = capi.Context(0)
prob = capi.Problem(ctx, "lennard_jones", dim=ATOMS)
stream = torch.cuda.ExternalStream(ctx.stream)
lb, ub = prob.bounds()
g = torch.Generator(device="cuda:0").manual_seed(4)
x = torch.rand((N, prob.nx), dtype=torch.float64, device="cuda:0", generator=g) * torch.tensor(ub - lb, device="cuda:0") + torch.tensor(lb, device="cuda:0")
f = torch.empty(N, dtype=torch.float64, device="cuda:0")
for _ in range(3):
    prob.eval_device(x.data_ptr(), N, f.data_ptr(), ctx.stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
REPS = 5
for _ in range(REPS):
    prob.eval_device(x.data_ptr(), N, f.data_ptr(), ctx.stream)
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / REPS
flops, _, byts = prob.work()
peak = ctx.fp64_peak_tflops(4096)
out = {"atoms": ATOMS, "swarm": N, "eval_ms": ms, "evals_per_s": N / (ms * 1e-3), "fp64_tflops": flops * N / (ms * 1e-3) / 1e12,
       "fp64_peak_tflops": peak, "frac_of_fp64_peak": flops * N / (ms * 1e-3) / 1e12 / peak, "hbm_gbs": byts * N / (ms * 1e-3) / 1e9}
GENS = 5
lib = capi.lib()
capi.check(lib.pgc_pso_evolve_device(prob._h, x.data_ptr(), f.data_ptr(), None, None, N, 1, 0.7298, 2.05, 2.05, 0.5, 5, 2, 4, 9, 1, None))
torch.cuda.synchronize()
t0 = time.perf_counter()
capi.check(lib.pgc_pso_evolve_device(prob._h, x.data_ptr(), f.data_ptr(), None, None, N, GENS, 0.7298, 2.05, 2.05, 0.5, 5, 2, 4, 9, 2, None))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out.update({"pso_generations": GENS, "pso_seconds": dt, "pso_generations_per_s": GENS / dt})
try:
    from oracle.pyoracle import reference
    R = reference()
    rp = R.problem("lennard_jones", ATOMS)
    cores = os.cpu_count() or 1
    xs = x[: 256 * cores].cpu().numpy()
    t0 = time.perf_counter()
    rp.thread_bfe(xs, nthreads=cores)
    dt = time.perf_counter() - t0
    out["cpu_reference"] = {"evals_per_s": xs.shape[0] / dt, "cores": cores, "sample": f"{xs.shape[0]} individuals, thread_bfe"}
except Exception as e:  # noqa: BLE001
    out["cpu_reference"] = {"unavailable": str(e)[:200]}
print(json.dumps(out))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bench_pso_lj.json").write_text(json.dumps(out, indent=1))
