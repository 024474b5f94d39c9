#!/usr/bin/env python
"""bench.py - headline benchmark: CEC2014 f1..f30, D=100, 1 Mi decision vectors per GPU through the C ABI.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--n INDIVIDUALS]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One STEP = one pass of the batch through all 30 CEC2014 functions (54 stage/combine kernel launches).
`value` = fitness evaluations / second, whole job (all ranks), inputs resident in HBM, timed with CUDA events
on the launching stream, max over ranks.  `e2e` = the same metric through pgc_eval_host (the pagmo::bfe contract:
host vectors in, host vectors out), pinned host buffers, H2D and D2H inside the timed region.
Multi-GPU: the batch shards by individual with no data-path collective (weak scaling, n per GPU fixed).

--impl reference times the reference's own CPU path (pagmo::thread_bfe over the unmodified cec2014 UDP, compiled
from the reference sources into oracle/_ref/libpagmo_ref.so) on all host cores, on a bounded sample of the same
workload.  The oracle is only ever used here as the CPU baseline / checker, never on the measured GPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DIM = 100
FUNCS = list(range(1, 31))
N_DEFAULT = 1 << 20
STAGE_DRAM_BYTES = 8.467e8  # measured once with ncu (see roofline.traffic_note); per launch at the default batch
ROTATIONS = {**{f: 1 for f in range(1, 23)}, 8: 0, 10: 0, 23: 4, 24: 2, 25: 3, 26: 5, 27: 5, 28: 5, 29: 3, 30: 3}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock / throttle-reason sampler for one GPU during the timed region (recipe: B200_PROFILING.md, clocks line).
    NVML in-process (nvidia_ml_py), one sample every 10 ms - a freshly spawned `nvidia-smi -lms` needs longer to print its first
    line than a 0.3 s timed region lasts, more so on an 8-GPU box; nvidia-smi stays as the fallback when NVML is not importable."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc, self._stop_evt = index, [], None, threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is not None:
            N = self.nvml
            reasons_fn = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._stop_evt.is_set():
                try:
                    sm = float(N.nvmlDeviceGetClockInfo(self.handle, N.NVML_CLOCK_SM))
                    mask = int(reasons_fn(self.handle))
                    self.rows.append([sm, self.max_sm, mask])
                except Exception:
                    pass
                self._stop_evt.wait(0.01)
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                c = [x.strip() for x in line.split(",")]
                mask = sum(bit for (nm, bit), v in zip(self.REASONS.items(), c[3:7]) if v.lower().startswith("active"))
                self.rows.append([float(c[0]), float(c[1]), mask])
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [r[0] for r in self.rows]
        mx = max([r[1] for r in self.rows], default=0)
        reasons = sorted(nm for nm, bit in self.REASONS.items() if any(r[2] & bit for r in self.rows))
        # the sampler also sees idle moments at the edges; "under load" = the upper half of the samples
        sm_load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(sm_load) if sm_load else None, "sm_max_mhz": mx or None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(sample_per_func: int, nthreads: int, steps: int = 1, warmup: int = 0):
    """evals/s of pagmo::bfe{thread_bfe{}} over the unmodified reference cec2014 UDPs (oracle/_ref), all 30 functions
    on `sample_per_func` individuals each.  Returns (evals_per_s, seconds_per_step)."""
    from oracle.pyoracle import reference
    R = reference()
    rng = np.random.default_rng(20141)
    xs = rng.uniform(-100.0, 100.0, (sample_per_func, DIM))
    probs = [R.problem("cec2014", f, DIM) for f in FUNCS]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for p in probs:
            p.thread_bfe(xs, nthreads=nthreads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    return len(FUNCS) * sample_per_func / per_step, per_step


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    sample = 512 * cores
    try:
        rate, per_step = cpu_reference_rate(sample, cores, steps=args.steps, warmup=args.warmup)
    except Exception as e:  # the prebuilt checker is missing: say so, never fake a number
        OUT.emit(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref not loadable: {e}"[:200]}))
        return 0
    desc = f"{sample} individuals x 30 functions per step (of the 1Mi-vector workload), thread_bfe on {cores} threads"
    line = {
        "impl": "reference", "metric": "fitness evals/sec (CEC2014 D=100)", "value": rate, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.n, args.gpus),
        "cpu_baseline": {"value": rate, "unit": "evals/s", "cores": cores, "kind": "reference", "sample": desc},
        "e2e": {"value": rate, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    OUT.emit(json.dumps(line))
    return 0


def workload_config(n, gpus):
    return {"workload": "cec2014 f1-f30 shifted/rotated/hybrid/composition, D=100, one pass of the batch per function",
            "individuals_per_gpu": n, "global_batch": n * gpus, "dim": DIM, "functions": 30,
            "tables": "synthetic seeded: orthogonal Mr, Os~U[-80,80), random shuffles (pagmo2_b200/synth.py on the native arm, oracle/cec_synth.c on the reference arm)",
            "inputs": "x~U[-100,100), seed 20141+rank; 839 MB per GPU (> 126 MB L2, no flush needed)",
            "parallelism": f"shard-by-individual x{gpus}, no collective"}


# ----------------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    from pagmo2_b200 import capi

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if world > 1:
        # one process per GPU: run (and first-touch the pinned e2e buffers) on the cores next to this GPU, so that eight host->device
        # streams do not cross the socket interconnect
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        except Exception:
            pass
    ctx = capi.Context(local)  # raises if there is no device: no CPU fallback
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    n = args.n

    from pagmo2_b200 import synth  # seeded synthetic data tables (numpy); nothing under oracle/ is touched on this arm's GPU path
    probs = []
    for f in FUNCS:
        mr, os_c, s = synth.cec2014_tables(f, DIM)
        probs.append(capi.Problem(ctx, "cec2014", prob_id=f, dim=DIM, rotation=mr, shift=os_c, shuffle=s))
    work = [p.work() for p in probs]  # (flops, transcendentals, bytes) per eval

    # inputs: pinned host copy (for e2e) and a resident device copy (for value)
    rng = np.random.default_rng(20141 + rank)
    h_x = ctx.pinned_array((n, DIM))
    CH = 1 << 16
    for i in range(0, n, CH):
        h_x[i:i + CH] = rng.uniform(-100.0, 100.0, (min(CH, n - i), DIM))
    h_f = ctx.pinned_array((n, 1))
    d_x = torch.empty((n, DIM), dtype=torch.float64, device=f"cuda:{local}")
    d_f = torch.empty((len(FUNCS), n), dtype=torch.float64, device=f"cuda:{local}")
    d_x.copy_(torch.from_numpy(h_x))
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(events=None):
        for i, p in enumerate(probs):
            if events is not None:
                events[i].record(stream)
            p.eval_device(d_x.data_ptr(), n, d_f[i].data_ptr(), ctx.stream)
        if events is not None:
            events[len(probs)].record(stream)

    # ---- FP64 ceiling (live, same run) ----
    fp64_peak = ctx.fp64_peak_tflops(2048)
    fp64_peak = max(fp64_peak, ctx.fp64_peak_tflops(8192))
    fp64_mma_peak = ctx.fp64_mma_peak_tflops(4096)

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launches
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(probs) + 1)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(args.steps):
        one_step(ev[k])
    e1.record(stream)
    barrier()
    launches = ctx.launches - launches0
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    ms_per_step = ms_total_max / args.steps
    evals_per_step = len(FUNCS) * n * world
    value = evals_per_step / (ms_per_step * 1e-3)

    per_func_ms = [statistics.mean(ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(args.steps)) for i in range(len(probs))]

    # ---- e2e: host vectors in / host vectors out through pgc_eval_host, pinned memory ----
    e2e_steps = max(0, min(args.steps, args.e2e_steps))
    for p in probs[:2]:
        p.eval_host_into(h_x, h_f)  # warm the staging ring
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        for p in probs:
            p.eval_host_into(h_x, h_f)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = evals_per_step * e2e_steps / float(t.item()) if e2e_steps else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline bookkeeping ----
    flops_step = sum(w[0] + w[1] for w in work) * n          # algorithmic FP64 ops, each libm call counted as 1
    rot_flops_step = sum(ROTATIONS[f] for f in FUNCS) * 2.0 * DIM * DIM * n
    step_s_rank = (ms_total / args.steps) * 1e-3
    achieved = flops_step / step_s_rank / 1e12
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    per_function = []
    for f, ms, w in zip(FUNCS, per_func_ms, work):
        per_function.append({
            "f": f, "ms": round(ms, 4), "evals_per_s": n / (ms * 1e-3),
            "fp64_tflops": (w[0] + w[1]) * n / (ms * 1e-3) / 1e12,
            "rotation_frac_of_fp64_peak": ROTATIONS[f] * 2.0 * DIM * DIM * n / (ms * 1e-3) / 1e12 / fp64_peak,
            "hbm_gbs": w[2] * n / (ms * 1e-3) / 1e9,
        })
    geomean = float(np.exp(np.mean([np.log(p["evals_per_s"]) for p in per_function])))

    # ---- BASELINE.json's second headline metric, measured in the same run on rank 0's GPU: NSGA-II generations/s at pop 65 536
    # (whole generations on the device: shuffles, FNDS + crowding, variation, batch evaluation, select_best_N_mo; nsga2.cpp:91-307)
    secondary = None
    if not args.no_secondary:
        try:
            import ctypes as C
            secondary = {"metric": "NSGA-II generations/sec (pop 65536)", "unit": "generations/s", "generations_timed": 5}
            for name, kw in (("zdt1_nx30", dict(family="zdt", prob_id=1, dim=30)), ("dtlz2_nx12_m3", dict(family="dtlz", prob_id=2, dim=12, nobj=3, param=100))):
                p2 = capi.Problem(ctx, **kw)
                NP = 65536
                d_x2, d_f2 = ctx.malloc(8 * NP * p2.nx), ctx.malloc(8 * NP * p2.nf)
                capi.check(capi.lib().pgc_population_init_device(p2._h, NP, 31, d_x2, d_f2, None, None))
                capi.check(capi.lib().pgc_nsga2_evolve_device(p2._h, d_x2, d_f2, NP, 2, 0.95, 10.0, 0.01, 50.0, 7, 0, None))
                ctx.synchronize()
                t0 = time.perf_counter()
                capi.check(capi.lib().pgc_nsga2_evolve_device(p2._h, d_x2, d_f2, NP, 5, 0.95, 10.0, 0.01, 50.0, 7, 2, None))
                ctx.synchronize()
                secondary[name] = 5.0 / (time.perf_counter() - t0)
                ctx.free(d_x2)
                ctx.free(d_f2)
                p2.close()
            secondary["value"] = secondary["zdt1_nx30"]
        except Exception as e:  # noqa: BLE001
            secondary = {"metric": "NSGA-II generations/sec (pop 65536)", "unavailable": str(e)[:200]}

    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = 512 * cores
        try:
            rate, secs = cpu_reference_rate(sample, cores, steps=2, warmup=0)
            cpu = {"value": rate, "unit": "evals/s", "cores": cores, "kind": "reference",
                   "sample": f"{sample} individuals x 30 functions x 2 passes, pagmo::thread_bfe (unmodified reference "
                             f"sources, oracle/_ref) on {cores} threads, {secs:.2f} s per pass"}
        except Exception as e:
            cpu = {"value": None, "unit": "evals/s", "cores": cores, "kind": "reference", "sample": f"unavailable: {e}"[:200]}

    line = {
        "metric": "fitness evals/sec (CEC2014 D=100)", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": len(FUNCS) * n * DIM * 8,
                "d2h_bytes_per_step": len(FUNCS) * n * 8, "steps": e2e_steps,
                "path": "pgc_eval_host: pinned host -> chunked H2D -> kernels -> D2H -> pinned host"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak, "traffic": STAGE_DRAM_BYTES,
                     "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one rotated stage launch (1 Mi x 100), ncu --set full, "
                                     "profiles/r1z_stage_kernel_ncu_full.csv (839.0 MB read + 6-8 MB written); algorithmic bytes of that launch = 8*(100+1)*2^20 = 8.47e8",
                     "peak_source": "pgc_measure_fp64_peak (DFMA loop, this run); MEASURED_PEAKS.json has no FP64 figure",
                     "dmma_probe_tflops": fp64_mma_peak,
                     "rotation_only_frac": rot_flops_step / step_s_rank / 1e12 / fp64_peak,
                     "hbm_gbs_achieved": sum(w[2] for w in work) * n / step_s_rank / 1e9, "hbm_peak_gbs": hbm_peak,
                     "note": "aggregate over the 62 launches of one step (50 rotated stage + 4 separable + 8 combine); per-function split in per_function"},
        # the same step against the driver-measured HBM copy bandwidth (MEASURED_PEAKS.json), in the contract's own vocabulary:
        # far below 1 because the step is bound by the FP64 pipe, not by memory (roofline above)
        "roofline_hbm": {"bound": "hbm", "achieved": sum(w[2] for w in work) * n / step_s_rank / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": sum(w[2] for w in work) * n / step_s_rank / 1e9 / hbm_peak, "traffic": STAGE_DRAM_BYTES,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks.get("hbm_gbs") else "of fallback 6650 GB/s"},
        "cpu_baseline": cpu,
        "secondary": secondary,
        "geomean_evals_per_s_per_gpu": geomean,
        "per_function": per_function,
    }
    OUT.emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


class QuietStdout:
    """Keep stdout to the ONE JSON line: libraries (NCCL's version banner, torchrun helpers) that print to fd 1 while the benchmark
    runs are sent to stderr; `emit` writes to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self._real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text: str):
        sys.stdout.flush()
        os.write(self._real, (text + "\n").encode())


OUT = None


def main():
    global OUT
    OUT = QuietStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--no-secondary", action="store_true", help="skip the NSGA-II generations/s measurement")
    ap.add_argument("--n", type=int, default=N_DEFAULT, help="individuals per GPU")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
