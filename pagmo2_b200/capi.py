"""ctypes binding of libpgc.so - the harness-side view of the C ABI in include/pagmo_cuda/pgc.h.

This module is plumbing for tests/ and bench.py: the product is the shared library and the header-only C++
adapters in include/pagmo_cuda/.  Nothing here computes: every call goes through the C ABI, and importing it
without a built library (or calling it without a CUDA device) fails loudly - there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
# PGC_LIBRARY_PATH: another build of the same library (kernel-variant experiments, scripts/build_variants.py)
LIB_PATH = Path(os.environ.get("PGC_LIBRARY_PATH") or HERE / "libpgc.so")

PGC_OK = 0
PGC_ERR_INVALID_ARGUMENT = -1
PGC_ERR_UNSUPPORTED = -2
PGC_ERR_CUDA = -3
PGC_ERR_OUT_OF_MEMORY = -4

FAMILY = {
    "rastrigin": 1, "ackley": 2, "griewank": 3, "schwefel": 4, "rosenbrock": 5, "cec2014": 6, "cec2013": 7, "zdt": 8,
    "dtlz": 9, "wfg": 10, "lennard_jones": 11, "hock_schittkowski_71": 14, "luksan_vlcek1": 15,
}
UNCONSTRAIN_METHODS = {"death penalty": 0, "kuri": 1, "weighted": 2, "ignore_c": 3, "ignore_o": 4}


class PgcError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"pgc status {status}: {msg}")
        self.status = status


class ProblemDesc(C.Structure):
    _fields_ = [
        ("family", C.c_int32), ("prob_id", C.c_uint32), ("dim", C.c_uint32), ("nobj", C.c_uint32), ("param", C.c_uint32),
        ("rotation", C.POINTER(C.c_double)), ("rotation_len", C.c_size_t),
        ("shift", C.POINTER(C.c_double)), ("shift_len", C.c_size_t),
        ("shuffle", C.POINTER(C.c_int32)), ("shuffle_len", C.c_size_t),
    ]


class AlgoDesc(C.Structure):
    """pgc_algo_desc: constructor arguments of the reference UDAs."""
    _fields_ = [
        ("algo", C.c_int32), ("gens", C.c_uint32), ("variant", C.c_uint32), ("variant_adptv", C.c_uint32), ("neighb_type", C.c_uint32),
        ("neighb_param", C.c_uint32), ("n_allowed", C.c_uint32), ("allowed_variants", C.c_uint32 * 18),
        ("F", C.c_double), ("CR", C.c_double), ("ftol", C.c_double), ("xtol", C.c_double),
        ("omega", C.c_double), ("eta1", C.c_double), ("eta2", C.c_double), ("max_vel", C.c_double),
        ("cr", C.c_double), ("eta_c", C.c_double), ("m", C.c_double), ("eta_m", C.c_double), ("seed", C.c_uint64),
        ("param_m", C.c_double), ("param_s", C.c_uint32), ("crossover", C.c_uint32), ("mutation", C.c_uint32), ("selection", C.c_uint32),
        ("cma_cc", C.c_double), ("cma_cs", C.c_double), ("cma_c1", C.c_double), ("cma_cmu", C.c_double), ("sigma0", C.c_double),
        ("force_bounds", C.c_uint32), ("memory", C.c_uint32),
        ("nspso_c1", C.c_double), ("nspso_c2", C.c_double), ("nspso_chi", C.c_double), ("nspso_v_coeff", C.c_double),
        ("leader_selection_range", C.c_uint32), ("diversity", C.c_uint32),
    ]


class GacoState(C.Structure):
    """pgc_gaco_state: the scalar members pagmo::gaco keeps between evolve() calls."""
    _fields_ = [("oracle", C.c_double), ("q", C.c_double), ("n_evalstop", C.c_uint32), ("n_impstop", C.c_uint32), ("gen_mark", C.c_uint32),
                ("initialized", C.c_uint32), ("fevals", C.c_uint64), ("memory", C.c_uint32), ("counter", C.c_uint32), ("h_archive", C.c_void_p),
                ("h_archive_len", C.c_size_t), ("has_champion", C.c_uint32), ("reserved_", C.c_uint32), ("champion_f", C.c_double)]


class MacoState(C.Structure):
    """pgc_maco_state: the members pagmo::maco keeps between evolve() calls."""
    _fields_ = [("q", C.c_double), ("n_evalstop", C.c_uint32), ("gen_mark", C.c_uint32), ("initialized", C.c_uint32), ("reserved_", C.c_uint32)]


class AlgoMemory(C.Structure):
    """pgc_algo_memory: the device arrays a UDA with memory = true keeps between evolve() calls."""
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p), ("u", C.c_void_p), ("initialized", C.c_int32), ("reserved_", C.c_int32),
                ("h_state", C.c_void_p), ("h_state_len", C.c_size_t)]


ALGO = {"de": 1, "sade": 2, "de1220": 3, "pso_gen": 4, "nsga2": 5, "sga": 6, "cmaes": 7, "nspso": 8, "xnes": 9}
NSPSO_DIVERSITY = {"crowding distance": 0, "niche count": 1, "max min": 2}
SGA_CROSSOVER = {"exponential": 0, "binomial": 1, "single": 2, "sbx": 3}
SGA_MUTATION = {"gaussian": 0, "uniform": 1, "polynomial": 2}
SGA_SELECTION = {"tournament": 0, "truncated": 1}
TOPOLOGY = {"unconnected": 0, "ring": 1, "fully_connected": 2}


class log_capture:
    """`with log_capture(ctx, verbosity, max_rows, row_len) as cap: prob.gaco_evolve(...)` then `cap.rows`: the reference's log lines of the
    gaco / maco / moead_gen call made inside (pgc_log_capture_begin / _end; same thread)."""

    def __init__(self, ctx, verbosity: int, max_rows: int, row_len: int):
        self.ctx, self.v, self.max_rows, self.row_len, self.rows = ctx, verbosity, max_rows, row_len, None

    def __enter__(self):
        L = lib()
        L.pgc_log_capture_begin.argtypes = [C.c_void_p, C.c_uint, C.c_size_t, C.c_size_t]
        L.pgc_log_capture_end.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        check(L.pgc_log_capture_begin(self.ctx._h, self.v, self.max_rows, self.row_len))
        return self

    def __exit__(self, *exc):
        buf = np.zeros((max(self.max_rows, 1), self.row_len))
        n = C.c_size_t()
        rc = lib().pgc_log_capture_end(self.ctx._h, buf.ctypes.data, C.byref(n))
        if exc[0] is None:
            check(rc)
        self.rows = buf[:n.value]
        return False


def algo_desc(name: str, gens: int = 1, seed: int = 0, **overrides) -> AlgoDesc:
    """Reference-default constructor arguments of UDA `name`, with keyword overrides (e.g. variant=7, ftol=0.)."""
    d = AlgoDesc()
    check(lib().pgc_algo_defaults(ALGO[name], gens, seed, C.byref(d)))
    for k, v in overrides.items():
        if k == "allowed_variants":
            d.n_allowed = len(v)
            for i, a in enumerate(v):
                d.allowed_variants[i] = a
        else:
            if not hasattr(d, k):
                raise AttributeError(f"pgc_algo_desc has no field {k!r}")
            setattr(d, k, v)
    return d


def topology_connections(kind: str, n: int, i: int, weight: float = 1.0):
    """topology::get_connections(i): (sources of the edges into i, weights)."""
    idx = np.empty(max(n, 1), dtype=np.uint64)
    w = np.empty(max(n, 1))
    cnt = C.c_size_t()
    check(lib().pgc_topology_connections(TOPOLOGY[kind], n, i, weight, idx.ctypes.data_as(C.POINTER(C.c_size_t)),
                                         w.ctypes.data_as(C.POINTER(C.c_double)), C.byref(cnt)))
    return idx[:cnt.value].astype(np.int64), w[:cnt.value].copy()


def philox_u01(seed: int, tag: int, generation: int, index: int, slot: int) -> float:
    out = C.c_double()
    check(lib().pgc_philox_u01(seed, tag, generation, index, slot, C.byref(out)))
    return out.value


_lib = None


def lib():
    """Load libpgc.so (raises if it was not built: run `python -m pagmo2_b200.build`)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(f"{LIB_PATH} not found - build it with `python -m pagmo2_b200.build`; "
                                    "there is no CPU fallback")
        L = C.CDLL(str(LIB_PATH))
        L.pgc_version.restype = C.c_char_p
        L.pgc_last_error.restype = C.c_char_p
        vp, sz, dp = C.c_void_p, C.c_size_t, C.POINTER(C.c_double)
        szp = C.POINTER(C.c_size_t)
        L.pgc_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.pgc_ctx_destroy.argtypes = [vp]
        L.pgc_ctx_stream.argtypes = [vp, C.POINTER(vp)]
        L.pgc_ctx_synchronize.argtypes = [vp]
        L.pgc_ctx_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.pgc_problem_create.argtypes = [vp, C.POINTER(ProblemDesc), C.POINTER(vp)]
        L.pgc_problem_destroy.argtypes = [vp]
        L.pgc_problem_translate.argtypes = [vp, dp, sz, C.POINTER(vp)]
        L.pgc_problem_decompose.argtypes = [vp, dp, dp, sz, C.c_int, C.c_int, C.POINTER(vp)]
        L.pgc_problem_unconstrain.argtypes = [vp, C.c_int, dp, sz, C.POINTER(vp)]
        L.pgc_problem_set_c_tol.argtypes = [vp, dp, sz]
        L.pgc_problem_c_tol.argtypes = [vp, dp]
        L.pgc_feasibility_device.argtypes = [vp, vp, sz, vp, vp]
        for fn in ("pgc_problem_nx", "pgc_problem_nobj", "pgc_problem_nf", "pgc_problem_nec", "pgc_problem_nic"):
            getattr(L, fn).argtypes = [vp, C.POINTER(sz)]
        L.pgc_problem_bounds.argtypes = [vp, dp, dp]
        L.pgc_problem_name.argtypes = [vp, C.c_char_p, sz]
        L.pgc_problem_work.argtypes = [vp, dp, dp, dp]
        L.pgc_eval_device.argtypes = [vp, vp, sz, vp, vp]
        L.pgc_eval_host.argtypes = [vp, vp, sz, vp]
        L.pgc_debug_cec2014_phase_cycles.argtypes = [vp, vp, sz, vp, C.POINTER(C.c_uint64)]
        L.pgc_malloc_device.argtypes = [vp, sz, C.POINTER(vp)]
        L.pgc_free_device.argtypes = [vp, vp]
        L.pgc_malloc_pinned.argtypes = [vp, sz, C.POINTER(vp)]
        L.pgc_free_pinned.argtypes = [vp, vp]
        L.pgc_memcpy_h2d.argtypes = [vp, vp, vp, sz]
        L.pgc_memcpy_d2h.argtypes = [vp, vp, vp, sz]
        L.pgc_fnds_host.argtypes = [vp, vp, sz, sz, szp, szp, szp, szp, szp]
        L.pgc_crowding_distance_host.argtypes = [vp, vp, sz, sz, dp]
        L.pgc_select_best_N_mo_host.argtypes = [vp, vp, sz, sz, sz, szp, szp]
        L.pgc_sort_population_mo_host.argtypes = [vp, vp, sz, sz, szp]
        u32p = C.POINTER(C.c_uint32)
        L.pgc_fnds_device.argtypes = [vp, vp, sz, sz, vp, vp, vp, vp, u32p, vp]
        L.pgc_crowding_fronts_device.argtypes = [vp, vp, sz, sz, vp, vp, C.c_uint32, C.c_int, vp, vp]
        L.pgc_select_best_N_mo_device.argtypes = [vp, vp, sz, sz, sz, vp, u32p, vp]
        L.pgc_philox_u01.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, dp]
        L.pgc_philox_permutation_device.argtypes = [vp, sz, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp]
        L.pgc_nsga2_variation_device.argtypes = [vp, vp, vp, vp, sz, sz, vp, vp, vp, vp, C.c_double, C.c_double, C.c_double,
                                                 C.c_double, C.c_uint64, C.c_uint32, vp, vp]
        L.pgc_nsga2_evolve_device.argtypes = [vp, vp, vp, sz, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64,
                                              C.c_uint32, vp]
        L.pgc_pso_evolve_device.argtypes = [vp, vp, vp, vp, vp, sz, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint,
                                            C.c_uint, C.c_uint, C.c_uint64, C.c_uint32, vp]
        L.pgc_pso_shard_step_device.argtypes = [vp, vp, vp, vp, vp, sz, C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_double,
                                                C.c_uint, C.c_uint64, C.c_uint32, C.c_int, vp]
        L.pgc_pso_shard_step_gbest_device.argtypes = [vp, vp, vp, vp, vp, sz, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_double,
                                                      C.c_uint, C.c_uint64, C.c_uint32, C.c_int, vp, vp]
        L.pgc_de_evolve_device.argtypes = [vp, vp, vp, sz, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_double, vp, C.c_uint,
                                           C.c_double, C.c_double, vp, vp, vp, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint), vp]
        L.pgc_hv_compute_host.argtypes = [vp, vp, sz, sz, dp, dp]
        L.pgc_hv_contributions_host.argtypes = [vp, vp, sz, sz, dp, dp]
        L.pgc_hv_device.argtypes = [vp, vp, sz, sz, dp, C.c_int, vp, vp]
        L.pgc_cmaes_sample_device.argtypes = [vp, vp, vp, C.c_double, sz, sz, C.c_uint64, C.c_uint32, vp, vp, vp]
        L.pgc_weighted_gram_device.argtypes = [vp, vp, vp, vp, vp, sz, sz, C.c_double, vp, vp]
        L.pgc_weighted_mean_device.argtypes = [vp, vp, vp, vp, sz, sz, vp, vp]
        L.pgc_cmaes_evolve_device.argtypes = [vp, vp, vp, sz, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                              C.c_double, C.c_int, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint), dp, vp]
        L.pgc_algo_defaults.argtypes = [C.c_int, C.c_uint, C.c_uint64, C.POINTER(AlgoDesc)]
        L.pgc_algo_evolve_device.argtypes = [vp, C.POINTER(AlgoDesc), vp, vp, sz, C.c_uint32, C.POINTER(C.c_uint), vp]
        L.pgc_population_init_device.argtypes = [vp, sz, C.c_uint64, vp, vp, vp, vp]
        L.pgc_select_best_device.argtypes = [vp, vp, vp, vp, sz, sz, sz, C.c_int, C.c_double, vp, vp, vp, szp, vp]
        L.pgc_fair_replace_device.argtypes = [vp, vp, vp, vp, sz, sz, sz, C.c_int, C.c_double, vp, vp, vp, sz, vp]
        L.pgc_topology_connections.argtypes = [C.c_int, sz, sz, C.c_double, szp, dp, szp]
        u64p = C.POINTER(C.c_uint64)
        L.pgc_island_create.argtypes = [vp, sz, sz, sz, C.POINTER(vp)]
        L.pgc_island_destroy.argtypes = [vp]
        L.pgc_island_size.argtypes = [vp, szp, szp, szp]
        L.pgc_island_upload.argtypes = [vp, vp, vp, vp]
        L.pgc_island_download.argtypes = [vp, vp, vp, vp]
        L.pgc_island_init.argtypes = [vp, C.c_uint64]
        L.pgc_island_evolve.argtypes = [vp, C.POINTER(AlgoDesc), C.POINTER(C.c_uint)]
        L.pgc_island_generation.argtypes = [vp, u32p]
        L.pgc_island_set_generation.argtypes = [vp, C.c_uint32]
        L.pgc_island_select.argtypes = [vp, C.c_int, C.c_double, szp]
        L.pgc_island_clear_outbox.argtypes = [vp]
        L.pgc_island_outbox_download.argtypes = [vp, vp, vp, vp, szp]
        L.pgc_island_inbox_upload.argtypes = [vp, sz, vp, vp, vp, sz]
        L.pgc_island_replace.argtypes = [vp, C.c_int, C.c_double, sz, vp, vp, szp]
        L.pgc_island_champion.argtypes = [vp, vp, vp]
        L.pgc_island_pointers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        L.pgc_comm_nccl_version.argtypes = [C.POINTER(C.c_int)]
        L.pgc_comm_init.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]
        L.pgc_comm_unique_id.argtypes = [vp, sz]
        L.pgc_comm_init_rank.argtypes = [C.c_int, C.c_int, C.c_int, vp, sz, C.POINTER(vp)]
        L.pgc_comm_destroy.argtypes = [vp]
        L.pgc_comm_size.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.pgc_migrate.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int), sz, u32p, u32p, u32p, sz]
        L.pgc_measure_fp64_peak.argtypes = [vp, C.c_int, dp]
        L.pgc_measure_fp64_mma_peak.argtypes = [vp, C.c_int, dp]
        _lib = L
    return _lib


def check(rc: int):
    if rc != PGC_OK:
        raise PgcError(rc, lib().pgc_last_error().decode(errors="replace"))


def device_count() -> int:
    n = C.c_int()
    check(lib().pgc_device_count(C.byref(n)))
    return n.value


class Context:
    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().pgc_ctx_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            lib().pgc_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        s = C.c_void_p()
        check(lib().pgc_ctx_stream(self._h, C.byref(s)))
        return s.value or 0

    def synchronize(self):
        check(lib().pgc_ctx_synchronize(self._h))

    @property
    def launches(self) -> int:
        n = C.c_uint64()
        check(lib().pgc_ctx_launch_count(self._h, C.byref(n)))
        return n.value

    def malloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        check(lib().pgc_malloc_device(self._h, nbytes, C.byref(p)))
        return p.value

    def free(self, ptr: int):
        check(lib().pgc_free_device(self._h, C.c_void_p(ptr)))

    def pinned_array(self, shape, dtype=np.float64) -> np.ndarray:
        """numpy array backed by page-locked host memory (freed when the context closes... never, tiny leak ok)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        check(lib().pgc_malloc_pinned(self._h, n, C.byref(p)))
        buf = (C.c_char * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def to_device(self, a: np.ndarray) -> int:
        a = np.ascontiguousarray(a)
        d = self.malloc(a.nbytes)
        check(lib().pgc_memcpy_h2d(self._h, C.c_void_p(d), a.ctypes.data_as(C.c_void_p), a.nbytes))
        return d

    def from_device(self, ptr: int, shape, dtype=np.float64) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        check(lib().pgc_memcpy_d2h(self._h, out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), out.nbytes))
        return out

    # ---- CMA-ES / xNES contractions (host-array convenience wrappers over the device entry points) ----
    def cmaes_sample(self, mean, bd, sigma: float, lam: int, seed: int, generation: int):
        """(x [lam x D], z [lam x D]): x_i = mean + sigma * BD z_i."""
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        bd = np.ascontiguousarray(bd, dtype=np.float64)
        D = mean.size
        dm, db = self.to_device(mean), self.to_device(bd)
        dz, dx = self.malloc(8 * max(lam * D, 1)), self.malloc(8 * max(lam * D, 1))
        try:
            check(lib().pgc_cmaes_sample_device(self._h, dm, db, sigma, lam, D, seed, generation, dz, dx, None))
            self.synchronize()
            return self.from_device(dx, (lam, D)), self.from_device(dz, (lam, D))
        finally:
            for p in (dm, db, dz, dx):
                self.free(p)

    def weighted_gram(self, rows, w, idx=None, center=None, scale_div: float = 1.0):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        D = rows.shape[1]
        bufs = [self.to_device(rows), self.to_device(w), self.malloc(8 * D * D), self.malloc(8 * D)]
        di = self.to_device(np.ascontiguousarray(idx, dtype=np.uint32)) if idx is not None else None
        dc = self.to_device(np.ascontiguousarray(center, dtype=np.float64)) if center is not None else None
        try:
            check(lib().pgc_weighted_gram_device(self._h, bufs[0], di, dc, bufs[1], w.size, D, scale_div, bufs[2], None))
            check(lib().pgc_weighted_mean_device(self._h, bufs[0], di, bufs[1], w.size, D, bufs[3], None))
            self.synchronize()
            return self.from_device(bufs[2], (D, D)), self.from_device(bufs[3], (D,))
        finally:
            for p in bufs + [q for q in (di, dc) if q]:
                self.free(p)

    # ---- hypervolume (pagmo::hypervolume::compute / contributions for 2 and 3 objectives) ----
    def hv_compute(self, points: np.ndarray, r_point) -> float:
        f = np.ascontiguousarray(points, dtype=np.float64)
        r = np.ascontiguousarray(r_point, dtype=np.float64)
        out = C.c_double()
        check(lib().pgc_hv_compute_host(self._h, f.ctypes.data_as(C.c_void_p), f.shape[0], r.size, r.ctypes.data_as(C.POINTER(C.c_double)),
                                        C.byref(out)))
        return out.value

    def hv_fpras(self, f, r, eps=1e-2, delta=1e-2, seed=0) -> float:
        """bf_fpras::compute on the device (pgc_hv_fpras_host)."""
        f = np.ascontiguousarray(f, dtype=np.float64)
        r = np.ascontiguousarray(r, dtype=np.float64)
        n, m = f.shape if f.ndim == 2 else (0, r.size)
        out = C.c_double()
        L = lib()
        L.pgc_hv_fpras_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_double, C.c_double, C.c_uint64,
                                        C.POINTER(C.c_double)]
        check(L.pgc_hv_fpras_host(self._h, f.ctypes.data, n, m, r.ctypes.data, eps, delta, seed, C.byref(out)))
        return out.value

    def hv_approx_extreme(self, f, r, greatest=False, use_exact=True, trivial_subcase_size=1, eps=1e-2, delta=1e-6, delta_multiplier=0.775,
                          alpha=0.2, initial_delta_coeff=0.1, gamma=0.25, seed=0) -> int:
        """bf_approx::least_contributor / greatest_contributor on the device (pgc_hv_approx_extreme_host)."""
        f = np.ascontiguousarray(f, dtype=np.float64)
        r = np.ascontiguousarray(r, dtype=np.float64)
        out = C.c_size_t()
        L = lib()
        L.pgc_hv_approx_extreme_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_double,
                                                 C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64, C.POINTER(C.c_size_t)]
        check(L.pgc_hv_approx_extreme_host(self._h, f.ctypes.data, f.shape[0], f.shape[1], r.ctypes.data, int(greatest), int(use_exact),
                                           trivial_subcase_size, eps, delta, delta_multiplier, alpha, initial_delta_coeff, gamma, seed,
                                           C.byref(out)))
        return out.value

    def hv_contributions(self, points: np.ndarray, r_point) -> np.ndarray:
        f = np.ascontiguousarray(points, dtype=np.float64)
        r = np.ascontiguousarray(r_point, dtype=np.float64)
        out = np.empty(max(f.shape[0], 1))
        check(lib().pgc_hv_contributions_host(self._h, f.ctypes.data_as(C.c_void_p), f.shape[0], r.size,
                                              r.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out[:f.shape[0]]

    # ---- multi-objective utilities (host-vector entry points, pagmo signatures) ----
    def fnds(self, f: np.ndarray) -> dict:
        """fast_non_dominated_sorting -> {'rank', 'dom_count', 'fronts'} (fronts in the reference's order)."""
        f = np.ascontiguousarray(f, dtype=np.float64)
        n, m = f.shape
        rank, dc, fi = (np.empty(max(n, 1), dtype=np.uint64) for _ in range(3))
        fo = np.empty(n + 1, dtype=np.uint64)
        nf = C.c_size_t()
        szp = C.POINTER(C.c_size_t)
        check(lib().pgc_fnds_host(self._h, f.ctypes.data_as(C.c_void_p), n, m, rank.ctypes.data_as(szp), dc.ctypes.data_as(szp),
                                  fi.ctypes.data_as(szp), fo.ctypes.data_as(szp), C.byref(nf)))
        fronts = [fi[int(fo[k]):int(fo[k + 1])].astype(np.int64) for k in range(nf.value)]
        return {"rank": rank[:n].astype(np.int64), "dom_count": dc[:n].astype(np.int64), "fronts": fronts}

    def crowding_distance(self, f: np.ndarray) -> np.ndarray:
        f = np.ascontiguousarray(f, dtype=np.float64)
        n, m = f.shape
        out = np.empty(max(n, 1))
        check(lib().pgc_crowding_distance_host(self._h, f.ctypes.data_as(C.c_void_p), n, m, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out[:n]

    def select_best_N_mo(self, f: np.ndarray, N: int) -> np.ndarray:
        f = np.ascontiguousarray(f, dtype=np.float64)
        n, m = f.shape if f.ndim == 2 else (0, 0)
        out = np.empty(max(n, 1), dtype=np.uint64)
        nout = C.c_size_t()
        szp = C.POINTER(C.c_size_t)
        check(lib().pgc_select_best_N_mo_host(self._h, f.ctypes.data_as(C.c_void_p), n, m, N, out.ctypes.data_as(szp), C.byref(nout)))
        return out[: nout.value].astype(np.int64)

    def sort_population_mo(self, f: np.ndarray) -> np.ndarray:
        f = np.ascontiguousarray(f, dtype=np.float64)
        n, m = f.shape if f.ndim == 2 else (0, 0)
        out = np.empty(max(n, 1), dtype=np.uint64)
        check(lib().pgc_sort_population_mo_host(self._h, f.ctypes.data_as(C.c_void_p), n, m, out.ctypes.data_as(C.POINTER(C.c_size_t))))
        return out[:n].astype(np.int64)

    # ---- generation operators ----
    def philox_permutation(self, n: int, seed: int, tag: int, generation: int) -> np.ndarray:
        d = self.malloc(4 * max(n, 1))
        check(lib().pgc_philox_permutation_device(self._h, n, seed, tag, generation, C.c_void_p(d), None))
        out = self.from_device(d, (n,), dtype=np.uint32)
        self.free(d)
        return out.astype(np.int64)

    def nsga2_variation(self, x, rank, cd, lb, ub, sh1, sh2, cr, eta_c, m, eta_m, seed, generation) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        NP, nx = x.shape
        bufs = [self.to_device(x), self.to_device(np.ascontiguousarray(rank, dtype=np.uint32)),
                self.to_device(np.ascontiguousarray(cd, dtype=np.float64)), self.to_device(np.ascontiguousarray(lb, dtype=np.float64)),
                self.to_device(np.ascontiguousarray(ub, dtype=np.float64)), self.to_device(np.ascontiguousarray(sh1, dtype=np.uint32)),
                self.to_device(np.ascontiguousarray(sh2, dtype=np.uint32)), self.malloc(8 * NP * nx)]
        try:
            check(lib().pgc_nsga2_variation_device(self._h, bufs[0], bufs[1], bufs[2], NP, nx, bufs[3], bufs[4], bufs[5], bufs[6], cr, eta_c,
                                                   m, eta_m, seed, generation, bufs[7], None))
            self.synchronize()
            return self.from_device(bufs[7], (NP, nx))
        finally:
            for b in bufs:
                self.free(b)

    def fp64_peak_tflops(self, iters: int = 4096) -> float:
        t = C.c_double()
        check(lib().pgc_measure_fp64_peak(self._h, iters, C.byref(t)))
        return t.value

    def fp64_mma_peak_tflops(self, iters: int = 4096) -> float:
        t = C.c_double()
        check(lib().pgc_measure_fp64_mma_peak(self._h, iters, C.byref(t)))
        return t.value


DECOMPOSE_METHODS = {"weighted": 0, "tchebycheff": 1, "bi": 2}


class Problem:
    """Handle on a device-side UDP (pgc_problem).  Mirrors the accessors of pagmo::problem."""

    def __init__(self, ctx: Context, family: str, prob_id: int = 0, dim: int = 0, nobj: int = 0, param: int = 0,
                 rotation: np.ndarray | None = None, shift: np.ndarray | None = None, shuffle: np.ndarray | None = None):
        self.ctx = ctx
        d = ProblemDesc()
        d.family, d.prob_id, d.dim, d.nobj, d.param = FAMILY[family], prob_id, dim, nobj, param
        keep = []
        if rotation is not None:
            r = np.ascontiguousarray(rotation, dtype=np.float64); keep.append(r)
            d.rotation, d.rotation_len = r.ctypes.data_as(C.POINTER(C.c_double)), r.size
        if shift is not None:
            s = np.ascontiguousarray(shift, dtype=np.float64); keep.append(s)
            d.shift, d.shift_len = s.ctypes.data_as(C.POINTER(C.c_double)), s.size
        if shuffle is not None:
            p = np.ascontiguousarray(shuffle, dtype=np.int32); keep.append(p)
            d.shuffle, d.shuffle_len = p.ctypes.data_as(C.POINTER(C.c_int32)), p.size
        self._h = C.c_void_p()
        check(lib().pgc_problem_create(ctx._h, C.byref(d), C.byref(self._h)))
        self._inner = None
        self._read_sizes()

    def _read_sizes(self):
        n = C.c_size_t()
        check(lib().pgc_problem_nx(self._h, C.byref(n))); self.nx = n.value
        check(lib().pgc_problem_nobj(self._h, C.byref(n))); self.nobj = n.value
        check(lib().pgc_problem_nf(self._h, C.byref(n))); self.nf = n.value
        check(lib().pgc_problem_nec(self._h, C.byref(n))); self.nec = n.value
        check(lib().pgc_problem_nic(self._h, C.byref(n))); self.nic = n.value

    @classmethod
    def _wrap(cls, inner: "Problem", handle) -> "Problem":
        p = cls.__new__(cls)
        p.ctx, p._h, p._inner = inner.ctx, handle, inner  # the wrapper borrows the inner problem: keep it alive
        p._read_sizes()
        return p

    def translate(self, translation) -> "Problem":
        """pagmo::translate{self, translation} on the device (translate.cpp:100-153)."""
        t = np.ascontiguousarray(translation, dtype=np.float64)
        h = C.c_void_p()
        check(lib().pgc_problem_translate(self._h, t.ctypes.data_as(C.POINTER(C.c_double)), t.size, C.byref(h)))
        return Problem._wrap(self, h)

    def unconstrain(self, method: str = "death penalty", weights=()) -> "Problem":
        """pagmo::unconstrain{self, method, weights} on the device (unconstrain.cpp:66-97,136-223)."""
        if method not in UNCONSTRAIN_METHODS:
            raise PgcError(-1, f"The method {method} is not supported (did you misspell?)")
        w = np.ascontiguousarray(weights, dtype=np.float64)
        h = C.c_void_p()
        check(lib().pgc_problem_unconstrain(self._h, UNCONSTRAIN_METHODS[method], w.ctypes.data_as(C.POINTER(C.c_double)), C.c_size_t(w.size),
                                            C.byref(h)))
        return Problem._wrap(self, h)

    def set_c_tol(self, c_tol) -> None:
        """problem::set_c_tol (problem.cpp:620-660): a vector of nec + nic tolerances, or one value for all."""
        t = np.asarray(c_tol, dtype=np.float64)
        if t.ndim == 0:
            if np.isnan(t):
                raise PgcError(-1, "The tolerance cannot be set to be NaN.")
            if t < 0:
                raise PgcError(-1, "The tolerance cannot be negative.")
            t = np.full(self.nec + self.nic, float(t))
        t = np.ascontiguousarray(t)
        check(lib().pgc_problem_set_c_tol(self._h, t.ctypes.data_as(C.POINTER(C.c_double)), C.c_size_t(t.size)))

    def c_tol(self) -> np.ndarray:
        t = np.empty(self.nec + self.nic)
        check(lib().pgc_problem_c_tol(self._h, t.ctypes.data_as(C.POINTER(C.c_double))))
        return t

    def feasibility_device(self, d_f: int, n: int, d_out: int, stream: int = 0):
        """problem::feasibility_f per row of a device matrix [n x nf] -> n bytes (1 feasible, 0 not)."""
        check(lib().pgc_feasibility_device(self._h, C.c_void_p(d_f), C.c_size_t(n), C.c_void_p(d_out), C.c_void_p(stream)))

    def decompose(self, weight, z, method: str = "weighted", adapt_ideal: bool = False) -> "Problem":
        """pagmo::decompose{self, weight, z, method, adapt_ideal} on the device (decompose.cpp:66-154)."""
        if method not in DECOMPOSE_METHODS:
            raise PgcError(-1, f"Decomposition method requested is: {method} while only one of ['weighted', 'tchebycheff', 'bi'] "
                               "are allowed")
        w = np.ascontiguousarray(weight, dtype=np.float64)
        zz = np.ascontiguousarray(z, dtype=np.float64)
        if zz.size != w.size:
            raise PgcError(-1, "Reference point size must be equal to the number of objectives. The size of the reference point is "
                               f"{zz.size} while the problem has {self.nobj} objectives")
        h = C.c_void_p()
        dp = C.POINTER(C.c_double)
        check(lib().pgc_problem_decompose(self._h, w.ctypes.data_as(dp), zz.ctypes.data_as(dp), w.size, DECOMPOSE_METHODS[method],
                                          int(adapt_ideal), C.byref(h)))
        return Problem._wrap(self, h)

    def close(self):
        if self._h:
            lib().pgc_problem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def name(self) -> str:
        buf = C.create_string_buffer(256)
        check(lib().pgc_problem_name(self._h, buf, 256))
        return buf.value.decode()

    def bounds(self):
        lb, ub = np.empty(self.nx), np.empty(self.nx)
        dp = C.POINTER(C.c_double)
        check(lib().pgc_problem_bounds(self._h, lb.ctypes.data_as(dp), ub.ctypes.data_as(dp)))
        return lb, ub

    def set_strict(self, on: bool = True):
        """cec2013: rotations in the reference's summation order (pgc_problem_set_strict)."""
        lib().pgc_problem_set_strict.argtypes = [C.c_void_p, C.c_int]
        check(lib().pgc_problem_set_strict(self._h, int(on)))

    def work(self):
        """(fp64 flops, libm calls, algorithmic bytes) per evaluation."""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        check(lib().pgc_problem_work(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def eval_device(self, d_dvs: int, n: int, d_fvs: int, stream: int = 0):
        check(lib().pgc_eval_device(self._h, C.c_void_p(d_dvs), n, C.c_void_p(d_fvs), C.c_void_p(stream)))

    def phase_cycles(self, d_dvs: int, n: int, d_fvs: int) -> dict:
        """cec2014 only: per-phase cycle totals of the stage kernel (debug aid)."""
        out = (C.c_uint64 * 7)()
        check(lib().pgc_debug_cec2014_phase_cycles(self._h, C.c_void_p(d_dvs), n, C.c_void_p(d_fvs), out))
        names = ["load", "weight", "token_wait", "gemm", "store_z", "epilogue", "warp_tiles"]
        return dict(zip(names, [int(v) for v in out]))

    def nsga2_evolve(self, x: np.ndarray, f: np.ndarray, gens: int, cr=0.95, eta_c=10., m=0.01, eta_m=50., seed=0, first_generation=0):
        """nsga2::evolve on the device: returns the evolved (x, f)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64)
        NP = x.shape[0]
        dx, df = self.ctx.to_device(x), self.ctx.to_device(f)
        try:
            check(lib().pgc_nsga2_evolve_device(self._h, dx, df, NP, gens, cr, eta_c, m, eta_m, seed, first_generation, None))
            return self.ctx.from_device(dx, x.shape), self.ctx.from_device(df, f.shape)
        finally:
            self.ctx.free(dx)
            self.ctx.free(df)

    def pso_evolve(self, x, f, v=None, gens=1, omega=0.7298, eta1=2.05, eta2=2.05, max_vel=0.5, variant=5, neighb_type=2, neighb_param=4,
                   seed=0, first_generation=1):
        """pso_gen::evolve on the device: returns (lbX, lbfit, V or None, Xcur)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        n = x.shape[0]
        dx, df, dc = self.ctx.to_device(x), self.ctx.to_device(f), self.ctx.malloc(x.nbytes)
        dv = self.ctx.to_device(np.ascontiguousarray(v, dtype=np.float64)) if v is not None else None
        try:
            check(lib().pgc_pso_evolve_device(self._h, dx, df, dv, dc, n, gens, omega, eta1, eta2, max_vel, variant, neighb_type,
                                              neighb_param, seed, first_generation, None))
            return (self.ctx.from_device(dx, x.shape), self.ctx.from_device(df, f.shape),
                    self.ctx.from_device(dv, x.shape) if dv else None, self.ctx.from_device(dc, x.shape))
        finally:
            for b in (dx, df, dc, dv):
                if b:
                    self.ctx.free(b)

    def cmaes_evolve(self, x, f, gens=1, cc=-1., cs=-1., c1=-1., cmu=-1., sigma0=0.5, ftol=1e-6, xtol=1e-6, force_bounds=False, seed=0,
                     first_generation=1):
        """cmaes::evolve on the device: returns (x, f, gens_done, sigma)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        dx, df = self.ctx.to_device(x), self.ctx.to_device(f)
        done, sigma = C.c_uint(), C.c_double()
        try:
            check(lib().pgc_cmaes_evolve_device(self._h, dx, df, x.shape[0], gens, cc, cs, c1, cmu, sigma0, ftol, xtol, int(force_bounds), seed,
                                                first_generation, C.byref(done), C.byref(sigma), None))
            return self.ctx.from_device(dx, x.shape), self.ctx.from_device(df, f.shape), done.value, sigma.value
        finally:
            self.ctx.free(dx)
            self.ctx.free(df)

    def xnes_evolve(self, x, f, gens=1, eta_mu=-1., eta_sigma=-1., eta_b=-1., sigma0=-1., ftol=1e-6, xtol=1e-6, force_bounds=False, seed=0,
                    first_generation=1):
        """xnes::evolve on the device: returns (x, f, gens_done, sigma)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        dx, df = self.ctx.to_device(x), self.ctx.to_device(f)
        done, sigma = C.c_uint(), C.c_double()
        L = lib()
        L.pgc_xnes_evolve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_double,
                                             C.c_double, C.c_double, C.c_int, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint), C.POINTER(C.c_double),
                                             C.c_void_p]
        try:
            check(L.pgc_xnes_evolve_device(self._h, dx, df, x.shape[0], gens, eta_mu, eta_sigma, eta_b, sigma0, ftol, xtol, int(force_bounds), seed,
                                           first_generation, C.byref(done), C.byref(sigma), None))
            return self.ctx.from_device(dx, x.shape), self.ctx.from_device(df, f.shape), done.value, sigma.value
        finally:
            self.ctx.free(dx)
            self.ctx.free(df)

    def de_evolve(self, x, f, gens=1, algo="de1220", variant=2, variant_adptv=1, F=0.8, CR=0.9, allowed=(2, 3, 7, 10, 13, 14, 15, 16),
                  ftol=1e-6, xtol=1e-6, seed=0, first_generation=1):
        """generational de / sade / de1220 on the device: returns (x, f, gens_done, F, CR, variant)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        NP = x.shape[0]
        al = np.ascontiguousarray(allowed, dtype=np.uint32)
        code = {"de": 0, "sade": 1, "de1220": 2}[algo]
        dx, df = self.ctx.to_device(x), self.ctx.to_device(f)
        dF, dC, dV = None, None, None
        done = C.c_uint()
        try:
            check(lib().pgc_de_evolve_device(self._h, dx, df, NP, gens, code, variant, variant_adptv, F, CR, al.ctypes.data_as(C.c_void_p),
                                             al.size, ftol, xtol, dF, dC, dV, seed, first_generation, C.byref(done), None))
            return self.ctx.from_device(dx, x.shape), self.ctx.from_device(df, f.shape), done.value
        finally:
            self.ctx.free(dx)
            self.ctx.free(df)

    def nspso_evolve(self, x, f, gens=1, omega=0.6, c1=2.0, c2=2.0, chi=1.0, v_coeff=0.5, leader_selection_range=60,
                     diversity="crowding distance", seed=0, first_generation=1, vel=None, best_x=None, best_f=None):
        """nspso::evolve on the device: returns (x, f, vel, best_x, best_f); the last three are None unless the memory arrays were given."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(x.shape[0], -1)
        n = x.shape[0]
        dx, df = self.ctx.to_device(x), self.ctx.to_device(f)
        mem = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (vel, best_x, best_f)]
        dmem = [None if a is None else self.ctx.to_device(a) for a in mem]
        L = lib()
        L.pgc_nspso_evolve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_double, C.c_double, C.c_double,
                                              C.c_double, C.c_double, C.c_uint, C.c_uint, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
        try:
            check(L.pgc_nspso_evolve_device(self._h, dx, df, n, gens, omega, c1, c2, chi, v_coeff, leader_selection_range,
                                            NSPSO_DIVERSITY[diversity], seed, first_generation, *dmem, None))
            out = [None if a is None else self.ctx.from_device(d, a.shape) for a, d in zip(mem, dmem)]
            return (self.ctx.from_device(dx, x.shape), self.ctx.from_device(df, f.shape), *out)
        finally:
            for d in (dx, df, *dmem):
                if d is not None:
                    self.ctx.free(d)

    def gaco_evolve(self, x, f, gens=1, ker=63, q=1.0, oracle=0.0, acc=0.01, threshold=1, n_gen_mark=7, impstop=100000, evalstop=100000,
                    focus=0.0, seed=0, first_generation=1, state=None, memory=False):
        """gaco::evolve on the device: returns (x, f, state, gens_done); `state` (GacoState) = the algorithm's scalar members (and, with
        memory=True, its archive), pass it back in to continue with the same algorithm object."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(x.shape[0], -1)
        st = state if state is not None else GacoState()
        if state is None and memory:
            st.memory = 1
            st._archive = np.zeros(ker * (x.shape[1] + 2))  # lives as long as the state object
            st.h_archive, st.h_archive_len = st._archive.ctypes.data, st._archive.size
        dx, df = self.ctx.to_device(x), self.ctx.to_device(f)
        done = C.c_uint()
        L = lib()
        L.pgc_gaco_evolve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double,
                                             C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_uint64, C.c_uint32, C.c_void_p,
                                             C.POINTER(C.c_uint), C.c_void_p]
        try:
            check(L.pgc_gaco_evolve_device(self._h, dx, df, x.shape[0], gens, ker, q, oracle, acc, threshold, n_gen_mark, impstop, evalstop, focus,
                                           seed, first_generation, C.byref(st), C.byref(done), None))
            return self.ctx.from_device(dx, x.shape), self.ctx.from_device(df, f.shape), st, done.value
        finally:
            self.ctx.free(dx)
            self.ctx.free(df)

    def maco_evolve(self, x, f, gens=1, ker=63, q=1.0, threshold=1, n_gen_mark=7, evalstop=100000, focus=0.0, seed=0, first_generation=1,
                    state=None):
        """maco::evolve on the device: returns (x, f, state, gens_done); `state` (MacoState) = the algorithm object's members."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(x.shape[0], -1)
        st = state if state is not None else MacoState()
        dx, df = self.ctx.to_device(x), self.ctx.to_device(f)
        done = C.c_uint()
        L = lib()
        L.pgc_maco_evolve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_double, C.c_uint, C.c_uint,
                                             C.c_uint, C.c_double, C.c_uint64, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint), C.c_void_p]
        try:
            check(L.pgc_maco_evolve_device(self._h, dx, df, x.shape[0], gens, ker, q, threshold, n_gen_mark, evalstop, focus, seed,
                                           first_generation, C.byref(st), C.byref(done), None))
            return self.ctx.from_device(dx, x.shape), self.ctx.from_device(df, f.shape), st, done.value
        finally:
            self.ctx.free(dx)
            self.ctx.free(df)

    def moead_gen_evolve(self, x, f, weights, neigh, gens=1, decomposition="tchebycheff", CR=1.0, F=0.5, eta_m=20.0, realb=0.9, limit=2,
                         preserve_diversity=True, seed=0, first_generation=1):
        """moead_gen::evolve on the device with the given weight vectors [n x nobj] and neighbourhoods [n x T]: returns (x, f)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(x.shape[0], -1)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        nb = np.ascontiguousarray(neigh, dtype=np.uint32)
        n = x.shape[0]
        dx, df = self.ctx.to_device(x), self.ctx.to_device(f)
        L = lib()
        L.pgc_moead_gen_evolve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint,
                                                  C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint, C.c_int, C.c_uint64,
                                                  C.c_uint32, C.c_void_p]
        try:
            check(L.pgc_moead_gen_evolve_device(self._h, dx, df, n, gens, w.ctypes.data, nb.ctypes.data, nb.shape[1],
                                                {"weighted": 0, "tchebycheff": 1, "bi": 2}[decomposition], CR, F, eta_m, realb, limit,
                                                1 if preserve_diversity else 0, seed, first_generation, None))
            return self.ctx.from_device(dx, x.shape), self.ctx.from_device(df, f.shape)
        finally:
            self.ctx.free(dx)
            self.ctx.free(df)

    def evolve(self, algo: "AlgoDesc", x, f, first_generation=1):
        """pagmo::algorithm::evolve on host arrays: upload, `pgc_algo_evolve_device`, download.  Returns (x, f, gens_done)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(x.shape[0], -1)
        dx, df = self.ctx.to_device(x), self.ctx.to_device(f)
        done = C.c_uint()
        try:
            check(lib().pgc_algo_evolve_device(self._h, C.byref(algo), dx, df, x.shape[0], first_generation, C.byref(done), None))
            self.ctx.synchronize()
            return self.ctx.from_device(dx, x.shape), self.ctx.from_device(df, f.shape), done.value
        finally:
            self.ctx.free(dx)
            self.ctx.free(df)

    def evolve_logged(self, algo: "AlgoDesc", x, f, verbosity: int, first_generation=1):
        """evolve() with algorithm::set_verbosity(verbosity): returns (x, f, gens_done, log [rows x row_len]) - the lines the
        reference's get_log() would hold (`pgc_algo_evolve_logged_device`)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(x.shape[0], -1)
        L = lib()
        L.pgc_algo_log_row_len.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
        L.pgc_algo_evolve_logged_device.argtypes = [C.c_void_p, C.POINTER(AlgoDesc), C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32,
                                                    C.POINTER(C.c_uint), C.c_void_p, C.c_uint, C.c_void_p, C.c_size_t,
                                                    C.POINTER(C.c_size_t), C.c_void_p]
        row_len = C.c_size_t()
        check(L.pgc_algo_log_row_len(self._h, algo.algo, C.byref(row_len)))
        max_rows = (algo.gens - 1) // max(verbosity, 1) + 1 if algo.gens else 0
        rows = np.zeros((max(max_rows, 1), row_len.value))
        dx, df = self.ctx.to_device(x), self.ctx.to_device(f)
        done, n_rows = C.c_uint(), C.c_size_t()
        try:
            check(L.pgc_algo_evolve_logged_device(self._h, C.byref(algo), dx, df, x.shape[0], first_generation, C.byref(done), None, verbosity,
                                                  rows.ctypes.data, max_rows, C.byref(n_rows), None))
            self.ctx.synchronize()
            return self.ctx.from_device(dx, x.shape), self.ctx.from_device(df, f.shape), done.value, rows[:n_rows.value]
        finally:
            self.ctx.free(dx)
            self.ctx.free(df)

    def evolve_memory(self, algo: "AlgoDesc", x, f, first_generation=1, state=None):
        """evolve() of a UDA built with memory = true (`pgc_algo_evolve_memory_device`): `state` is what the previous call returned
        (None: the first call, the state is drawn as the reference's first evolve() draws it).  Returns (x, f, gens_done, state),
        state = {"a", "b", "c", "u"} host arrays (pgc_algo_memory: F / CR / variant, velocities, nspso's archive)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(x.shape[0], -1)
        n, nx, nf = x.shape[0], x.shape[1], f.shape[1]
        shapes = {"a": ((n, nx), np.float64), "b": ((n, nx), np.float64), "c": ((n, nf), np.float64), "u": ((n,), np.uint32)}
        host = {k: (np.zeros(sh, dt) if state is None else np.ascontiguousarray(state[k], dtype=dt).reshape(sh)) for k, (sh, dt) in shapes.items()}
        dev = {k: self.ctx.to_device(v) for k, v in host.items()}
        dx, df = self.ctx.to_device(x), self.ctx.to_device(f)
        es = None
        if algo.algo in (ALGO["cmaes"], ALGO["xnes"]):  # their state lives on the host
            L0 = lib()
            L0.pgc_es_state_len.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_size_t)]
            ln = C.c_size_t()
            check(L0.pgc_es_state_len(algo.algo, nx, C.byref(ln)))
            es = np.zeros(ln.value) if state is None else np.ascontiguousarray(state["es"], dtype=np.float64).copy()
        mem = AlgoMemory(dev["a"], dev["b"], dev["c"], dev["u"], 0 if state is None else 1, 0, es.ctypes.data if es is not None else None,
                         es.size if es is not None else 0)
        done = C.c_uint()
        L = lib()
        L.pgc_algo_evolve_memory_device.argtypes = [C.c_void_p, C.POINTER(AlgoDesc), C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32,
                                                    C.POINTER(C.c_uint), C.POINTER(AlgoMemory), C.c_void_p]
        try:
            check(L.pgc_algo_evolve_memory_device(self._h, C.byref(algo), dx, df, n, first_generation, C.byref(done), C.byref(mem), None))
            self.ctx.synchronize()
            out = {k: self.ctx.from_device(dev[k], host[k].shape, dtype=host[k].dtype) for k in host}
            if es is not None:
                out["es"] = es
            return self.ctx.from_device(dx, x.shape), self.ctx.from_device(df, f.shape), done.value, out
        finally:
            for d in (dx, df, *dev.values()):
                self.ctx.free(d)

    def eval_host_into(self, dvs: np.ndarray, fvs: np.ndarray):
        n = dvs.size // self.nx
        check(lib().pgc_eval_host(self._h, dvs.ctypes.data_as(C.c_void_p), n, fvs.ctypes.data_as(C.c_void_p)))

    def eval_host(self, dvs: np.ndarray) -> np.ndarray:
        dvs = np.ascontiguousarray(dvs, dtype=np.float64)
        if dvs.size % self.nx:
            raise ValueError("decision-vector batch size is not a multiple of nx")
        n = dvs.size // self.nx
        fvs = np.empty((n, self.nf))
        self.eval_host_into(dvs, fvs)
        return fvs


def _rate(rate):
    """migration rate as the C ABI takes it: (is_fraction, value) - a float is a fraction of the group, an int a count."""
    return (1, float(rate)) if isinstance(rate, float) else (0, float(rate))


class Island:
    """pgc_island: one island's population, migration outbox and inbox, resident on the problem's device (island.cu)."""

    def __init__(self, prob: "Problem", n: int, max_migrants: int = 1, max_in_edges: int = 1):
        self.prob, self.n, self.nx, self.nf = prob, n, prob.nx, prob.nf
        self.cap, self.slots = max(max_migrants, 1), max(max_in_edges, 1)
        h = C.c_void_p()
        check(lib().pgc_island_create(prob._h, n, max_migrants, max_in_edges, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().pgc_island_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self, seed: int):
        check(lib().pgc_island_init(self._h, seed))

    def upload(self, ids, x, f):
        ids, x, f = np.ascontiguousarray(ids, np.uint64), np.ascontiguousarray(x, np.float64), np.ascontiguousarray(f, np.float64)
        check(lib().pgc_island_upload(self._h, ids.ctypes.data, x.ctypes.data, f.ctypes.data))

    def download(self):
        ids, x, f = np.empty(self.n, np.uint64), np.empty((self.n, self.nx)), np.empty((self.n, self.nf))
        check(lib().pgc_island_download(self._h, ids.ctypes.data, x.ctypes.data, f.ctypes.data))
        return ids, x, f

    def evolve(self, algo: AlgoDesc) -> int:
        done = C.c_uint()
        check(lib().pgc_island_evolve(self._h, C.byref(algo), C.byref(done)))
        return done.value

    def select(self, rate) -> int:
        k = C.c_size_t()
        frac, r = _rate(rate)
        check(lib().pgc_island_select(self._h, frac, r, C.byref(k)))
        return k.value

    def outbox(self):
        ids, x, f, k = np.empty(self.cap, np.uint64), np.empty((self.cap, self.nx)), np.empty((self.cap, self.nf)), C.c_size_t()
        check(lib().pgc_island_outbox_download(self._h, ids.ctypes.data, x.ctypes.data, f.ctypes.data, C.byref(k)))
        return ids[:k.value], x[:k.value], f[:k.value]

    def inbox_upload(self, slot: int, ids, x, f):
        ids, x, f = np.ascontiguousarray(ids, np.uint64), np.ascontiguousarray(x, np.float64), np.ascontiguousarray(f, np.float64)
        k = ids.shape[0]
        check(lib().pgc_island_inbox_upload(self._h, slot, ids.ctypes.data if k else None, x.ctypes.data if k else None,
                                            f.ctypes.data if k else None, k))

    def replace(self, rate, n_slots: int, log: bool = True):
        """fair_replace with the rows of inbox slots [0, n_slots); returns [(migrant id, slot)] of the immigrants now in the population."""
        frac, r = _rate(rate)
        if not log:
            check(lib().pgc_island_replace(self._h, frac, r, n_slots, None, None, None))
            return []
        m = max(self.cap * self.slots, 1)
        ids, slot, na = np.empty(m, np.uint64), np.empty(m, np.uint32), C.c_size_t()
        check(lib().pgc_island_replace(self._h, frac, r, n_slots, ids.ctypes.data, slot.ctypes.data, C.byref(na)))
        return [(int(ids[i]), int(slot[i])) for i in range(na.value)]

    def replace_enqueue(self, rate, counts, log: bool = True):
        """the device half of replace(): asynchronous; `counts` = rows in each inbox slot used (known from the senders' policies)."""
        frac, r = _rate(rate)
        c = np.ascontiguousarray(counts, dtype=np.uint64)
        L = lib()
        L.pgc_island_replace_enqueue.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_size_t, C.c_void_p, C.c_int]
        check(L.pgc_island_replace_enqueue(self._h, frac, r, c.size, c.ctypes.data, int(log)))

    def replace_collect(self):
        """[(migrant id, slot)] of the last replace_enqueue(log=True): waits for its log copies only; [] when nothing is pending."""
        m = max(self.cap * self.slots, 1)
        ids, slot, na = np.empty(m, np.uint64), np.empty(m, np.uint32), C.c_size_t()
        L = lib()
        L.pgc_island_replace_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        check(L.pgc_island_replace_collect(self._h, ids.ctypes.data, slot.ctypes.data, C.byref(na)))
        return [(int(ids[i]), int(slot[i])) for i in range(na.value)]

    def champion(self):
        x, f = np.empty(self.nx), np.empty(1)
        check(lib().pgc_island_champion(self._h, x.ctypes.data, f.ctypes.data))
        return x, float(f[0])


class Comm:
    """pgc_comm: NCCL communicator(s) of this process.  Comm.all_local(devices): one process, several GPUs;
    Comm.from_unique_id(...): one process per GPU (rank 0 calls Comm.unique_id() and the application distributes the bytes)."""

    def __init__(self, handle):
        self._h = handle

    @staticmethod
    def nccl_version() -> int:
        v = C.c_int()
        check(lib().pgc_comm_nccl_version(C.byref(v)))
        return v.value

    @staticmethod
    def all_local(devices) -> "Comm":
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        check(lib().pgc_comm_init(len(devices), arr, C.byref(h)))
        return Comm(h)

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().pgc_comm_unique_id(buf, 128))
        return buf.raw

    @staticmethod
    def from_unique_id(device: int, nranks: int, rank: int, uid: bytes) -> "Comm":
        h = C.c_void_p()
        buf = C.create_string_buffer(uid, 128)
        check(lib().pgc_comm_init_rank(device, nranks, rank, buf, 128, C.byref(h)))
        return Comm(h)

    def close(self):
        if getattr(self, "_h", None):
            lib().pgc_comm_destroy(self._h)
            self._h = None


def migrate(comm, islands, owner_rank, edges):
    """pgc_migrate: islands = list with None for islands of other processes; edges = [(src, dst, slot)]."""
    n = len(islands)
    ptrs = (C.c_void_p * n)(*[(i._h if i is not None else None) for i in islands])
    own = (C.c_int * n)(*owner_rank)
    m = len(edges)
    es = (C.c_uint32 * max(m, 1))(*[e[0] for e in edges])
    ed = (C.c_uint32 * max(m, 1))(*[e[1] for e in edges])
    el = (C.c_uint32 * max(m, 1))(*[e[2] for e in edges])
    check(lib().pgc_migrate(comm._h if comm is not None else None, ptrs, own, n, es, ed, el, m))
