#!/usr/bin/env python
"""bench.py - headline benchmark: CEC2014 f1..f30, D=100, 1 Mi decision vectors per GPU through the C ABI.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--n INDIVIDUALS]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One STEP = one pass of the batch through all 30 CEC2014 functions (54 stage/combine kernel launches).
`value` = fitness evaluations / second, whole job (all ranks), inputs resident in HBM, timed with CUDA events
on the launching stream, max over ranks.  `e2e` = the same metric through pgc_eval_host (the pagmo::bfe contract:
host vectors in, host vectors out), pinned host buffers, H2D and D2H inside the timed region.
Multi-GPU: the batch shards by individual with no data-path collective (weak scaling, n per GPU fixed).

--workload cfg5 measures BASELINE config 5 instead (8 GPU islands running sade on CEC2013 D=50, ring migration through
pgc_migrate: ncclSend / ncclRecv between GPUs): island-generations/s and migrations/s, strong scaling (the 8 islands are spread
over the ranks).  The default workload (cfg2, the configuration BASELINE.json's metric is quoted on) also reports, on rank 0 and
after its timed region, BASELINE's second headline (NSGA-II generations/s at pop 65 536) and one-GPU figures for cfg1 / cfg4 / cfg5
under "secondary" - the LAST key of the line, mirrored in config.secondary.

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


def device_of_rank(local, world):
    """Which GPU local rank `local` drives.  On this pool's 8-GPU boxes GPUs 0-3 and 4-7 sit behind two host bridges of ~116 GB/s
    each (scripts/h2d_probe.py, profiles/r2j_h2d_probe.json: {0,1,2,3} copies 116 GB/s from pinned host memory, {0,1,4,5} 218), so
    when fewer ranks than visible GPUs run, the ranks alternate between the two halves - the host-buffer (e2e) path is PCIe-bound.
    PGC_BENCH_SPREAD=0 keeps rank r on GPU r."""
    try:
        import torch
        visible = torch.cuda.device_count()
    except Exception:
        return local
    if os.environ.get("PGC_BENCH_SPREAD", "1") == "0" or world <= 1 or visible < 2 * world or visible % 2:
        return local
    half = visible // 2
    return (local % 2) * half + local // 2


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
def _timed(fn, sync):
    sync()
    t0 = time.perf_counter()
    fn()
    sync()
    return time.perf_counter() - t0


def measure_secondary(ctx, capi, device):
    """One-GPU figures for the other BASELINE configs, each guarded on its own: a failure is reported, never hidden."""
    import ctypes as C
    L = capi.lib()
    out = {"metric": "NSGA-II generations/sec (pop 65536)", "unit": "generations/s"}

    def guarded(key, fn):
        try:
            out[key] = fn()
        except Exception as e:  # noqa: BLE001
            out[key] = {"unavailable": str(e)[:200]}

    def nsga2():
        r = {"generations_timed": 5, "what": "nsga2::evolve on the device: shuffles, FNDS + crowding of N, tournament + SBX + mutation, batch "
                                             "evaluation, select_best_N_mo over 2N (nsga2.cpp:91-307)"}
        ref = {}
        try:
            ref = json.loads((ROOT / "profiles" / "r2_ref_nsga2_timing.json").read_text())["cases"]
        except Exception:
            pass
        for name, kw in (("zdt1_nx30", dict(family="zdt", prob_id=1, dim=30)), ("dtlz2_nx12_m3", dict(family="dtlz", prob_id=2, dim=12, nobj=3, param=100))):
            p2 = capi.Problem(ctx, **kw)
            NP = 65536
            d_x2, d_f2 = ctx.malloc(8 * NP * p2.nx), ctx.malloc(8 * NP * p2.nf)
            capi.check(L.pgc_population_init_device(p2._h, NP, 31, d_x2, d_f2, None, None))
            capi.check(L.pgc_nsga2_evolve_device(p2._h, d_x2, d_f2, NP, 2, 0.95, 10.0, 0.01, 50.0, 7, 0, None))
            dt = _timed(lambda: capi.check(L.pgc_nsga2_evolve_device(p2._h, d_x2, d_f2, NP, 5, 0.95, 10.0, 0.01, 50.0, 7, 2, None)), ctx.synchronize)
            fronts = len(ctx.fnds(ctx.from_device(d_f2, (NP, p2.nf)))["fronts"])
            r[name] = {"generations_per_s": 5.0 / dt, "ms_per_generation": dt / 5 * 1e3, "fronts_in_population": fronts,
                       # latency model: the select_best_N_mo sort of 2N points peels about this many levels per generation
                       "us_per_level": dt / 5 * 1e6 / max(fronts, 1)}
            try:  # roofline of the second headline metric: the dominance relation over the 2N points of a generation
                import torch
                pr = torch.cuda.get_device_properties(device)
                clock_hz = 1e3 * getattr(pr, "clock_rate", 1965000)
                peak = pr.multi_processor_count * 64 * clock_hz  # FP64 compares (DSETP) per second: 64 lanes per SM per clock
                tests = float(2 * NP) ** 2 * p2.nf               # objective compares of an all-pairs dominance pass over 2N points
                r[name]["roofline"] = {"bound": "fp64 compare issue, all-pairs model", "achieved": tests / (dt / 5), "peak": peak,
                                       "unit": "objective compares/s", "frac": tests / (dt / 5) / peak,
                                       "note": "(2N)^2 * M compares per generation / generation time; the sort-based count (M = 2) and the "
                                               "position ranges of the resident level loop do fewer compares than all-pairs, and the level "
                                               "loop is latency-bound (~17 us per level), so this is an equivalent rate, not pipe utilisation"}
            except Exception:
                pass
            if name in ref:
                r[name]["cpu_reference_generations_per_s"] = ref[name]["extrapolated_generations_per_s_at_65536"]
                r[name]["cpu_reference_note"] = ("unmodified nsga2::evolve on one core, measured at pop 2048-16384 and extrapolated with the "
                                                 f"fitted exponent {ref[name]['fitted_exponent']:.2f} (profiles/r2_ref_nsga2_timing.json)")
            ctx.free(d_x2)
            ctx.free(d_f2)
            p2.close()
        return r

    def cfg1():  # Rastrigin D=10, pop 1024, de1220, 100 generations
        p = capi.Problem(ctx, "rastrigin", dim=10)
        NP = 1024
        d_x, d_f = ctx.malloc(8 * NP * 10), ctx.malloc(8 * NP)
        a = capi.algo_desc("de1220", gens=100, seed=41, ftol=0.0, xtol=0.0)
        capi.check(L.pgc_population_init_device(p._h, NP, 23, d_x, d_f, None, None))
        capi.check(L.pgc_algo_evolve_device(p._h, C.byref(a), d_x, d_f, NP, 1, None, None))
        reps = 5
        dt = _timed(lambda: [capi.check(L.pgc_algo_evolve_device(p._h, C.byref(a), d_x, d_f, NP, 101 + 100 * k, None, None)) for k in range(reps)],
                    ctx.synchronize)
        ctx.free(d_x)
        ctx.free(d_f)
        p.close()
        return {"what": "cfg1: rastrigin D=10, pop 1024, de1220, 100 generations per evolve()", "generations_per_s": 100 * reps / dt,
                "evals_per_s": 100 * reps * NP / dt, "us_per_generation": dt / (100 * reps) * 1e6}

    def cfg4():  # Lennard-Jones 150 atoms (D=444), the swarm of BASELINE cfg4 (262 144 particles), pairwise-energy kernel
        import torch
        p = capi.Problem(ctx, "lennard_jones", dim=150)
        NP = 262144
        x = torch.rand((NP, p.nx), dtype=torch.float64, device=f"cuda:{device}") * 6 - 3
        f = torch.empty(NP, dtype=torch.float64, device=f"cuda:{device}")
        p.eval_device(x.data_ptr(), NP, f.data_ptr(), ctx.stream)
        dt = _timed(lambda: [p.eval_device(x.data_ptr(), NP, f.data_ptr(), ctx.stream) for _ in range(5)], ctx.synchronize)
        p.close()
        pairs = 150 * 149 // 2
        rate = 5 * NP / dt
        # 13 FP64-pipe instructions per pair (7 of them FMAs: 20 flop): the pipe's issue ceiling at the DFMA probe's rate
        ceiling = ctx.fp64_peak_tflops(4096) * 1e12 / 2.0 / (pairs * 13)
        return {"what": "cfg4 kernel: lennard_jones 150 atoms (D=444), 262144 individuals resident", "evals_per_s": rate,
                "fp64_tflops_20_per_pair": rate * pairs * 20 / 1e12,
                "roofline": {"bound": "fp64 issue", "achieved": rate, "peak": ceiling, "unit": "evals/s", "frac": rate / ceiling,
                             "note": "13 FP64-pipe instructions per pair; peak = DFMA probe TF/s / 2 flop per thread-level DFMA / (11175 pairs x 13 thread-level instructions)"}}

    def cfg5():
        return measure_cfg5(capi, [device], rank=0, world=1, comm=None, rounds=4)

    def next_rows():  # the bfe-caller UDAs and the many-objective hypervolume, one figure each
        import numpy as np
        rng = np.random.default_rng(7)
        r = {}
        p = capi.Problem(ctx, "rastrigin", dim=30)
        lb, ub = p.bounds()
        x = rng.uniform(lb, ub, (65536, 30))
        f = p.eval_host(x)
        p.gaco_evolve(x, f, gens=2, seed=1)
        dt = _timed(lambda: p.gaco_evolve(x, f, gens=20, seed=1), ctx.synchronize)
        r["gaco_rastrigin30_pop65536_gens_per_s"] = 20 / dt
        p.close()
        p = capi.Problem(ctx, "zdt", prob_id=1, dim=30)
        lb, ub = p.bounds()
        x = rng.uniform(lb, ub, (16384, 30))
        f = p.eval_host(x)
        p.maco_evolve(x, f, gens=2, seed=1)
        dt = _timed(lambda: p.maco_evolve(x, f, gens=10, seed=1), ctx.synchronize)
        r["maco_zdt1_pop16384_gens_per_s"] = 10 / dt
        p.close()
        pts = rng.uniform(0.05, 1, (1024, 4))
        pts /= np.linalg.norm(pts, axis=1, keepdims=True)
        ctx.hv_compute(pts, np.full(4, 1.25))
        r["hv_wfg_4obj_1024pts_compute_ms"] = _timed(lambda: ctx.hv_compute(pts, np.full(4, 1.25)), ctx.synchronize) * 1e3
        # 8(f) row 1: unconstrain{luksan_vlcek1 D=30} (death penalty) against its inner problem alone, 1 Mi rows resident
        import torch
        inner = capi.Problem(ctx, "luksan_vlcek1", dim=30)
        inner.set_c_tol(1.0)
        un = inner.unconstrain("death penalty")
        n_u = 1 << 20
        xs = torch.rand((n_u, 30), dtype=torch.float64, device=f"cuda:{device}") * 3 - 1.5
        fi = torch.empty((n_u, inner.nf), dtype=torch.float64, device=f"cuda:{device}")
        fu = torch.empty(n_u, dtype=torch.float64, device=f"cuda:{device}")
        for q, o in ((inner, fi), (un, fu)):
            q.eval_device(xs.data_ptr(), n_u, o.data_ptr(), ctx.stream)
        ctx.synchronize()
        r["luksan_vlcek1_d30_1Mi_ms"] = _timed(lambda: [inner.eval_device(xs.data_ptr(), n_u, fi.data_ptr(), ctx.stream) for _ in range(5)],
                                                ctx.synchronize) / 5 * 1e3
        r["unconstrain_luksan_vlcek1_d30_1Mi_ms"] = _timed(lambda: [un.eval_device(xs.data_ptr(), n_u, fu.data_ptr(), ctx.stream) for _ in range(5)],
                                                            ctx.synchronize) / 5 * 1e3
        un.close()
        inner.close()
        r["what"] = ("gaco (rastrigin D=30, 65536 ants) and maco (ZDT1 nx=30, 16384 ants) generations/s incl. the host round trip of the "
                     "population per call; hypervolume of 1024 points in 4 objectives (device WFG) incl. upload; one batch of 1 Mi decision vectors "
                     "through luksan_vlcek1 D=30 (1 objective + 28 equality constraints) and through unconstrain{luksan_vlcek1} (death penalty)")
        return r

    guarded("nsga2_pop65536", nsga2)
    guarded("next_rows", next_rows)
    guarded("cfg1_de1220", cfg1)
    guarded("cfg4_lennard_jones", cfg4)
    guarded("cfg5_islands", cfg5)
    v = out.get("nsga2_pop65536", {}).get("zdt1_nx30", {})
    out["value"] = v.get("generations_per_s") if isinstance(v, dict) else None
    return out


def secondary_summary(sec):
    """the headline numbers of `secondary`, short enough to live inside `config`"""
    def g(*path):
        d = sec
        for k in path:
            d = d.get(k) if isinstance(d, dict) else None
        return round(d, 3) if isinstance(d, (int, float)) else d
    return {"nsga2_pop65536_zdt1_gens_per_s": g("nsga2_pop65536", "zdt1_nx30", "generations_per_s"),
            "nsga2_pop65536_dtlz2_gens_per_s": g("nsga2_pop65536", "dtlz2_nx12_m3", "generations_per_s"),
            "nsga2_cpu_reference_zdt1_gens_per_s": g("nsga2_pop65536", "zdt1_nx30", "cpu_reference_generations_per_s"),
            "cfg1_de1220_gens_per_s": g("cfg1_de1220", "generations_per_s"),
            "cfg4_lj150_evals_per_s": g("cfg4_lennard_jones", "evals_per_s"),
            "cfg5_island_gens_per_s": g("cfg5_islands", "island_generations_per_s"),
            "cfg5_evals_per_s": g("cfg5_islands", "evals_per_s"), "cfg5_migrations_per_s": g("cfg5_islands", "migrations_per_s"),
            "gaco_pop65536_gens_per_s": g("next_rows", "gaco_rastrigin30_pop65536_gens_per_s"),
            "maco_pop16384_gens_per_s": g("next_rows", "maco_zdt1_pop16384_gens_per_s"),
            "hv_wfg_4obj_1024pts_ms": g("next_rows", "hv_wfg_4obj_1024pts_compute_ms")}


CFG5 = {"islands": 8, "pop": 1024, "dim": 50, "func": 12, "gens_per_round": 50}


def measure_cfg5(capi, devices, rank, world, comm, rounds, warm_rounds=1, log=True):
    """BASELINE cfg5: 8 islands (one per GPU when 8 ranks run) x sade on CEC2013 D=50, ring topology, one migration per 50
    generations through pgc_migrate.  This process materialises the islands of its rank(s)."""
    from pagmo2_b200 import synth
    from pagmo2_b200.archipelago import ResidentArchipelago
    c = CFG5
    mr, os_ = synth.cec2013_tables(c["dim"])
    per = c["islands"] // world
    spec = []
    for g in range(c["islands"]):
        owner = g // per
        spec.append(dict(device=devices[(g % per) % len(devices)] if world > 1 else devices[g % len(devices)], family="cec2013",
                         problem_kw=dict(prob_id=c["func"], dim=c["dim"], rotation=mr, shift=os_),
                         algo=capi.algo_desc("sade", gens=c["gens_per_round"], seed=11 + g, ftol=0.0, xtol=0.0), pop_size=c["pop"], seed=200 + g,
                         r_rate=1, s_rate=1, owner=owner if world > 1 else 0))
    a = ResidentArchipelago(spec, topology="ring", seed=1, comm=comm, my_ranks=(rank,), log=log)
    a.evolve(warm_rounds)
    a.synchronize()
    l0 = sum(a.ctx[g].launches for g in a.local)
    m0 = len(a.log)
    a.phase_seconds.clear()
    t0 = time.perf_counter()
    a.evolve(rounds)
    a.synchronize()
    dt = time.perf_counter() - t0
    gens = rounds * c["gens_per_round"]
    return {"what": f"cfg5: {c['islands']} islands x sade (defaults, ftol = xtol = 0) on cec2013 f{c['func']} D={c['dim']}, pop {c['pop']} per island, "
                    f"ring(1.0), 1 migrant per edge every {c['gens_per_round']} generations; device-resident islands (pgc_island) and "
                    "pgc_migrate (NCCL send/recv between GPUs, device copy inside one)",
            "seconds": dt, "rounds": rounds, "local_islands": len(a.local),
            "island_generations_per_s": gens * c["islands"] / dt, "evals_per_s": gens * c["pop"] * c["islands"] / dt,
            "migrations_per_s": (len(a.log) - m0) * (c["islands"] / max(len(a.local), 1)) / dt, "migrations_logged_locally": len(a.log) - m0,
            "launches_per_generation_per_island": (sum(a.ctx[g].launches for g in a.local) - l0) / gens / max(len(a.local), 1),
            "host_seconds_per_phase": dict(a.phase_seconds), "champions_f_local": a.champions_f().tolist()}


def measure_adapter_e2e(n):
    """The same batch through the COMPILED plugin call - pagmo::bfe{cuda_bfe{}}(problem{cuda_cec2014}, std::vector<double>) on a
    pageable vector, returning a fresh vector (tests/cpp/test_adapters --bench) - next to e2e's ctypes + pinned-buffer path."""
    exe = ROOT / "tests" / "cpp" / "_bin" / "test_adapters"
    if not exe.exists():
        return None
    try:
        r = subprocess.run([str(exe), "--bench", str(n)], capture_output=True, text=True, timeout=300)
        for ln in r.stdout.splitlines():
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": (r.stdout + r.stderr)[-200:]}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:200]}


def run_cfg5(args):
    """--workload cfg5 (run by hand / gpurun --gpus N; the driver's default is cfg2)."""
    import torch
    from pagmo2_b200 import capi
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    local = device_of_rank(local, world)
    torch.cuda.set_device(local)
    comm, dist = None, None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [capi.Comm.unique_id() if rank == 0 else None]  # the 128-byte NCCL id travels through the application's own channel
        dist.broadcast_object_list(box, src=0)
        comm = capi.Comm.from_unique_id(local, world, rank, box[0])
    sampler = ClockSampler(local)
    sampler.start()
    # NCCL opens its send/recv channels lazily, peer by peer: the first ~8 exchanges of a ring over 8 ranks take ~100 ms each
    # (measured: 200 ms/round with 3 warm-up rounds, 2.2 ms/round after 10), so multi-rank runs warm up for at least 10 rounds
    res = measure_cfg5(capi, [local], rank, world, comm, rounds=max(args.steps, 1), warm_rounds=max(args.warmup, 10 if world > 1 else 1))
    clocks = sampler.stop()
    if dist is not None:
        t = torch.tensor([res["seconds"], float(res["migrations_logged_locally"])], dtype=torch.float64, device=f"cuda:{local}")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        secs, migr = float(tmax[0].item()), float(t[1].item())
    else:
        secs, migr = res["seconds"], float(res["migrations_logged_locally"])
    if rank == 0:
        c = CFG5
        gens = res["rounds"] * c["gens_per_round"]
        line = {"metric": "island generations/sec (cfg5: 8 GPU islands, sade, CEC2013 D=50, ring migration)", "value": gens * c["islands"] / secs,
                "unit": "island-generations/s", "n_gpus": world, "steps": res["rounds"], "warmup": max(args.warmup, 1),
                "ms_per_step": secs / res["rounds"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": res["what"], **c, "parallelism": f"{c['islands']} islands over {world} GPU(s), "
                                                "migrants over pgc_migrate (ncclSend/ncclRecv)"},
                "evals_per_s": gens * c["pop"] * c["islands"] / secs, "migrations_per_s": migr / secs, "migrations": migr, "clocks": clocks,
                "launches_per_generation_per_island": res["launches_per_generation_per_island"], "rank0": res}
        OUT.emit(json.dumps(line))
    if comm is not None:
        torch.cuda.synchronize()
        comm.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0

# ----------------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    from pagmo2_b200 import capi

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    local = device_of_rank(local, world)
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
    fp64_dfma_peak = fp64_peak
    # the rotation issues DMMA and the epilogues DFMA; both go through ONE pipe (profiles/r1_probe_fp64_dmma_dfma_share_pipe.json),
    # so the roofline denominator is the higher of the two probes of that pipe
    fp64_peak = max(fp64_dfma_peak, fp64_mma_peak)

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
            "f": f, "ms": round(ms, 4), "evals_per_s": round(n / (ms * 1e-3)),
            "fp64_tflops": round((w[0] + w[1]) * n / (ms * 1e-3) / 1e12, 3),
            "rotation_frac_of_fp64_peak": round(ROTATIONS[f] * 2.0 * DIM * DIM * n / (ms * 1e-3) / 1e12 / fp64_peak, 4),
            "hbm_gbs": round(w[2] * n / (ms * 1e-3) / 1e9, 1),
        })
    geomean = float(np.exp(np.mean([np.log(p["evals_per_s"]) for p in per_function])))

    # ---- BASELINE.json's second headline metric (NSGA-II generations/s at pop 65 536) and one-GPU figures of the other configs,
    # measured in the same run on rank 0's GPU after the timed region
    secondary = None
    if not args.no_secondary:
        secondary = measure_secondary(ctx, capi, local)

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
        "config": {**workload_config(n, world), "device_of_rank0": local,
                   "placement": "ranks alternate between GPUs 0-3 and 4-7 when fewer ranks than visible GPUs run (device_of_rank)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": len(FUNCS) * n * DIM * 8,
                "d2h_bytes_per_step": len(FUNCS) * n * 8, "steps": e2e_steps,
                "path": "pgc_eval_host: pinned host -> chunked H2D -> kernels -> D2H -> pinned host",
                "h2d_gbs": e2e_value * DIM * 8 / 1e9 if e2e_value else None,
                "ceiling": "host -> device copies of this box, measured by scripts/h2d_probe.py (profiles/r2j_h2d_probe.json): 55.5 GB/s for "
                           "one GPU, 111 for two, 116 for GPUs {0,1,2,3} (one host bridge), 218 for {0,1,4,5}, 188 for all eight"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak, "traffic": STAGE_DRAM_BYTES,
                     "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one rotated stage launch (1 Mi x 100), ncu --set full, "
                                     "profiles/r1z_stage_kernel_ncu_full.csv (839.0 MB read + 6-8 MB written); algorithmic bytes of that launch = 8*(100+1)*2^20 = 8.47e8",
                     "peak_source": "max(DFMA probe, DMMA probe) of the FP64 pipe, both measured in this run (pgc_measure_fp64_peak / "
                                    "pgc_measure_fp64_mma_peak); MEASURED_PEAKS.json has no FP64 figure",
                     "dfma_probe_tflops": fp64_dfma_peak, "dmma_probe_tflops": fp64_mma_peak,
                     "rotation_only_frac": rot_flops_step / step_s_rank / 1e12 / fp64_peak,
                     "hbm_gbs_achieved": sum(w[2] for w in work) * n / step_s_rank / 1e9, "hbm_peak_gbs": hbm_peak,
                     "note": "aggregate over the 62 launches of one step (50 rotated stage + 4 separable + 8 combine); per-function split in per_function"},
        # the same step against the driver-measured HBM copy bandwidth (MEASURED_PEAKS.json), in the contract's own vocabulary:
        # far below 1 because the step is bound by the FP64 pipe, not by memory (roofline above)
        "roofline_hbm": {"bound": "hbm", "achieved": sum(w[2] for w in work) * n / step_s_rank / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": sum(w[2] for w in work) * n / step_s_rank / 1e9 / hbm_peak, "traffic": STAGE_DRAM_BYTES,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks.get("hbm_gbs") else "of fallback 6650 GB/s"},
        "cpu_baseline": cpu,
        "geomean_evals_per_s_per_gpu": geomean,
        "per_function": per_function,
        "secondary": secondary,  # LAST key: survives in the tail of a truncated record
    }
    if secondary:
        line["config"]["secondary"] = secondary_summary(secondary)
    adapter = measure_adapter_e2e(n)
    if adapter:
        line["e2e"]["adapter"] = adapter
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
    ap.add_argument("--workload", choices=["cfg2", "cfg5"], default="cfg2")
    ap.add_argument("--cfg5-pop", type=int, default=CFG5["pop"], help="population per island of --workload cfg5 (the configuration leaves it open)")
    args = ap.parse_args()
    CFG5["pop"] = args.cfg5_pop
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "cfg5":
        return run_cfg5(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
