"""ctypes loaders for the two checkers.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* ``Oracle``    - oracle/liboracle.so, the plain-C restatement (oracle/restate_*.c + cec_synth.c).
* ``Reference`` - oracle/_ref/libpagmo_ref.so, the UNMODIFIED reference sources compiled against oracle/shim
                  (built in the authoring container only; the prebuilt .so travels to the GPU box).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "liboracle.so"
REF_SO = HERE / "_ref" / "libpagmo_ref.so"
REFERENCE_ROOT = Path("/root/reference")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_size_p = C.POINTER(C.c_size_t)


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def _ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int_p)


def _sp(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_size_p)


def build(ref: bool = True) -> None:
    """Compile the checkers (make -C oracle).  The reference library is only attempted when /root/reference exists."""
    targets = ["liboracle.so"]
    if ref and (REFERENCE_ROOT / "src" / "problem.cpp").exists():
        targets.append("_ref/libpagmo_ref.so")
    subprocess.run(["make", "-s", "-j8", "-C", str(HERE), *targets], check=True)


SIMPLE = {"rastrigin": 1, "ackley": 2, "griewank": 3, "schwefel": 4, "rosenbrock": 5}
CEC_NCOMP = 10


UNCONSTRAIN_METHODS = {"death penalty": 0, "kuri": 1, "weighted": 2, "ignore_c": 3, "ignore_o": 4}


class OracleProblem(C.Structure):
    _fields_ = [("family", C.c_int), ("prob_id", C.c_uint), ("dim", C.c_uint), ("nobj", C.c_uint), ("param", C.c_uint),
                ("rotation", c_double_p), ("shift", c_double_p), ("shuffle", c_int_p)]


FAMILY_ID = {"rastrigin": 1, "ackley": 2, "griewank": 3, "schwefel": 4, "rosenbrock": 5, "cec2014": 6, "cec2013": 7, "zdt": 8, "dtlz": 9,
             "lennard_jones": 11}


class Oracle:
    def __init__(self):
        if not ORACLE_SO.exists():
            build(ref=False)
        self.lib = L = C.CDLL(str(ORACLE_SO))
        L.oracle_cec2014_compact_shift.restype = C.c_size_t

    # ---- synthetic CEC tables (cec_synth.c) ----
    def cec2014_tables(self, func: int, dim: int):
        """Returns (Mr[10*dim*dim], Os_lines[10*100], S[10*dim] int32) exactly as the reference ctor sees them."""
        mr = np.empty(CEC_NCOMP * dim * dim)
        os_ = np.empty(CEC_NCOMP * 100)
        s = np.empty(CEC_NCOMP * dim, dtype=np.int32)
        self.lib.cec2014_synth_rotation(C.c_uint(func), C.c_uint(dim), _dp(mr))
        self.lib.cec2014_synth_shift(C.c_uint(func), _dp(os_))
        self.lib.cec2014_synth_shuffle(C.c_uint(func), C.c_uint(dim), _ip(s))
        return mr, os_, s

    def cec2014_compact_shift(self, lines: np.ndarray, dim: int) -> np.ndarray:
        out = np.empty(lines.size)
        k = self.lib.oracle_cec2014_compact_shift(_dp(lines), C.c_size_t(lines.size // 100), C.c_uint(dim), _dp(out))
        return out[:k].copy()

    def cec2014_problem_tables(self, func: int, dim: int):
        """(Mr, Os_compacted, S): the m_rotation_matrix / m_origin_shift / m_shuffle members of cec2014{func, dim}."""
        mr, lines, s = self.cec2014_tables(func, dim)
        return mr, self.cec2014_compact_shift(lines, dim), s

    def cec2013_tables(self, dim: int):
        mr = np.empty(CEC_NCOMP * dim * dim)
        os_ = np.empty(CEC_NCOMP * 100)
        self.lib.cec2013_synth_md(C.c_uint(dim), _dp(mr))
        self.lib.cec2013_synth_shift(_dp(os_))
        return mr, os_

    # ---- restated evaluators ----
    def simple(self, family: str, xs: np.ndarray) -> np.ndarray:
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        n, d = xs.shape
        out = np.empty(n)
        rc = self.lib.oracle_simple_batch(C.c_int(SIMPLE[family]), C.c_size_t(d), _dp(xs), C.c_size_t(n), _dp(out))
        if rc:
            raise ValueError(f"oracle_simple_batch failed rc={rc}")
        return out

    def zdt(self, prob_id: int, xs: np.ndarray) -> np.ndarray:
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        n, d = xs.shape
        out = np.empty((n, 2))
        if self.lib.oracle_zdt_batch(C.c_uint(prob_id), _dp(xs), C.c_size_t(n), C.c_size_t(d), _dp(out)):
            raise ValueError("oracle_zdt_batch failed")
        return out

    def dtlz(self, prob_id: int, xs: np.ndarray, fdim: int, alpha: int = 100) -> np.ndarray:
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        n, d = xs.shape
        out = np.empty((n, fdim))
        if self.lib.oracle_dtlz_batch(C.c_uint(prob_id), _dp(xs), C.c_size_t(n), C.c_size_t(d), C.c_size_t(fdim),
                                      C.c_uint(alpha), _dp(out)):
            raise ValueError("oracle_dtlz_batch failed")
        return out

    def wfg(self, prob_id: int, xs: np.ndarray, dim_obj: int, dim_k: int) -> np.ndarray:
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        n, d = xs.shape
        out = np.empty((n, dim_obj))
        if self.lib.oracle_wfg_batch(C.c_uint(prob_id), C.c_size_t(d), C.c_size_t(dim_obj), C.c_size_t(dim_k), _dp(xs), C.c_size_t(n), _dp(out)):
            raise ValueError(f"oracle_wfg_batch: invalid WFG{prob_id} configuration (dim_dvs={d}, dim_obj={dim_obj}, dim_k={dim_k})")
        return out

    def translate_rows(self, xs: np.ndarray, t: np.ndarray) -> np.ndarray:
        """translate::batch_fitness de-shifting (translate.cpp:137-150)."""
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        t = np.ascontiguousarray(t, dtype=np.float64)
        out = np.empty_like(xs)
        self.lib.oracle_translate_rows(_dp(xs), C.c_size_t(xs.shape[0]), C.c_size_t(xs.shape[1]), _dp(t), _dp(out))
        return out

    def decompose_rows(self, fs: np.ndarray, weight, z, method: str) -> np.ndarray:
        """decompose_objectives per row (multi_objective.cpp:582-638)."""
        fs = np.ascontiguousarray(fs, dtype=np.float64)
        w = np.ascontiguousarray(weight, dtype=np.float64)
        zz = np.ascontiguousarray(z, dtype=np.float64)
        out = np.empty(fs.shape[0])
        if self.lib.oracle_decompose_rows(_dp(fs), C.c_size_t(fs.shape[0]), C.c_size_t(fs.shape[1]), _dp(w), _dp(zz),
                                          C.c_int({"weighted": 0, "tchebycheff": 1, "bi": 2}[method]), _dp(out)):
            raise ValueError("oracle_decompose_rows failed")
        return out

    def hock_schittkowski_71(self, xs: np.ndarray) -> np.ndarray:
        """hock_schittkowski_71::fitness per row: [objective | equality | inequality] (hock_schittkowski_71.cpp:48-55)."""
        xs = np.ascontiguousarray(xs, dtype=np.float64).reshape(-1, 4)
        out = np.empty((xs.shape[0], 3))
        self.lib.oracle_hs71_batch(_dp(xs), C.c_size_t(xs.shape[0]), _dp(out))
        return out

    def luksan_vlcek1(self, xs: np.ndarray) -> np.ndarray:
        """luksan_vlcek1::fitness per row: [objective | dim - 2 equalities] (luksan_vlcek1.cpp:60-77)."""
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        n, d = xs.shape
        out = np.empty((n, d - 1))
        if self.lib.oracle_luksan_vlcek1_batch(C.c_size_t(d), _dp(xs), C.c_size_t(n), _dp(out)):
            raise ValueError("oracle_luksan_vlcek1_batch failed")
        return out

    def unconstrain_rows(self, fs: np.ndarray, nobj: int, nec: int, nic: int, c_tol, method: str, weights=None) -> np.ndarray:
        """unconstrain::penalize per row (unconstrain.cpp:136-223)."""
        fs = np.ascontiguousarray(fs, dtype=np.float64)
        tol = np.ascontiguousarray(c_tol, dtype=np.float64)
        w = np.ascontiguousarray(weights if weights is not None else np.zeros(nec + nic), dtype=np.float64)
        m = UNCONSTRAIN_METHODS[method]
        out = np.empty((fs.shape[0], 1 if m == 4 else nobj))
        if self.lib.oracle_unconstrain_rows(_dp(fs), C.c_size_t(fs.shape[0]), C.c_size_t(nobj), C.c_size_t(nec), C.c_size_t(nic), _dp(tol),
                                            C.c_int(m), _dp(w), _dp(out)):
            raise ValueError("oracle_unconstrain_rows failed")
        return out

    def lennard_jones(self, atoms: int, xs: np.ndarray) -> np.ndarray:
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        out = np.empty(xs.shape[0])
        if self.lib.oracle_lj_batch(C.c_uint(atoms), _dp(xs), C.c_size_t(xs.shape[0]), _dp(out)):
            raise ValueError("oracle_lj_batch failed")
        return out

    # ---- multi-objective utilities (restate_mo_utils.c) ----
    def fnds(self, f: np.ndarray, nolist: bool = False):
        """fast_non_dominated_sorting; nolist=True: the O(n)-memory, OpenMP form for full-size inputs (same results)."""
        f = np.ascontiguousarray(f, dtype=np.float64)
        n, m = f.shape
        rank, dc, fi = (np.empty(n, dtype=np.uint64) for _ in range(3))
        fo = np.empty(n + 1, dtype=np.uint64)
        nf = C.c_size_t()
        fn = self.lib.oracle_fnds_nolist if nolist else self.lib.oracle_fnds
        if fn(_dp(f), C.c_size_t(n), C.c_size_t(m), _sp(rank), _sp(dc), _sp(fi), _sp(fo), C.byref(nf)):
            raise ValueError("oracle_fnds failed")
        fronts = [fi[int(fo[k]):int(fo[k + 1])].astype(np.int64) for k in range(nf.value)]
        return {"rank": rank.astype(np.int64), "dom_count": dc.astype(np.int64), "fronts": fronts}

    def crowding_distance(self, f: np.ndarray) -> np.ndarray:
        f = np.ascontiguousarray(f, dtype=np.float64)
        out = np.empty(f.shape[0])
        if self.lib.oracle_crowding_distance(_dp(f), C.c_size_t(f.shape[0]), C.c_size_t(f.shape[1]), _dp(out)):
            raise ValueError("oracle_crowding_distance failed")
        return out

    def select_best_N_mo(self, f: np.ndarray, N: int) -> np.ndarray:
        f = np.ascontiguousarray(f, dtype=np.float64)
        out = np.empty(max(f.shape[0], 1), dtype=np.uint64)
        nout = C.c_size_t()
        if self.lib.oracle_select_best_N_mo(_dp(f), C.c_size_t(f.shape[0]), C.c_size_t(f.shape[1]), C.c_size_t(N), _sp(out),
                                            C.byref(nout)):
            raise ValueError("oracle_select_best_N_mo failed")
        return out[: nout.value].astype(np.int64)

    def sort_population_mo(self, f: np.ndarray) -> np.ndarray:
        f = np.ascontiguousarray(f, dtype=np.float64)
        out = np.empty(max(f.shape[0], 1), dtype=np.uint64)
        if self.lib.oracle_sort_population_mo(_dp(f), C.c_size_t(f.shape[0]), C.c_size_t(f.shape[1]), _sp(out)):
            raise ValueError("oracle_sort_population_mo failed")
        return out[: f.shape[0]].astype(np.int64)

    def problem(self, family, prob_id=0, dim=0, nobj=1, param=0, tables=None):
        """oracle_problem handle (keeps the table arrays alive)."""
        p = OracleProblem(FAMILY_ID[family], prob_id, dim, nobj, param, None, None, None)
        p._keep = tables
        if tables is not None:
            p.rotation, p.shift = _dp(tables[0]), _dp(tables[1])
            if len(tables) > 2:
                p.shuffle = _ip(tables[2])
        return p

    def pso_evolve(self, prob, lb, ub, x, f, v=None, gens=1, omega=0.7298, eta1=2.05, eta2=2.05, max_vel=0.5, variant=5, neighb_type=2,
                   neighb_param=4, seed=0, first_generation=1):
        """restated pso_gen::evolve: returns (lbX, lbfit, V, Xcur)."""
        x = np.array(x, dtype=np.float64, order="C")
        f = np.array(f, dtype=np.float64, order="C").reshape(-1)
        n, dim = x.shape
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        vv = None if v is None else np.array(v, dtype=np.float64, order="C")
        vout = vv if vv is not None else np.empty((n, dim))
        xcur = np.empty((n, dim))
        if vv is None:
            # velocities are drawn inside; fetch them by a second identical call is wasteful - expose through xcur only
            pass
        rc = self.lib.oracle_pso_evolve(C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), _dp(vv) if vv is not None else None, _dp(xcur),
                                        C.c_size_t(n), C.c_size_t(dim), C.c_uint(gens), C.c_double(omega), C.c_double(eta1),
                                        C.c_double(eta2), C.c_double(max_vel), C.c_uint(variant), C.c_uint(neighb_type),
                                        C.c_uint(neighb_param), C.c_uint64(seed), C.c_uint32(first_generation))
        if rc:
            raise ValueError("oracle_pso_evolve failed")
        return x, f, (vv if vv is not None else None), xcur

    def de_evolve(self, prob, lb, ub, x, f, gens=1, algo="de1220", variant=2, variant_adptv=1, F=0.8, CR=0.9,
                  allowed=(2, 3, 7, 10, 13, 14, 15, 16), ftol=1e-6, xtol=1e-6, seed=0, first_generation=1, sequential=False):
        """restated generational de / sade / de1220: returns (x, f, gens_done, F, CR, variant).
        sequential=True: same Philox draws, the reference's evaluate-and-select-one-at-a-time order."""
        x = np.array(x, dtype=np.float64, order="C")
        f = np.array(f, dtype=np.float64, order="C").reshape(-1)
        NP, dim = x.shape
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        al = np.ascontiguousarray(allowed, dtype=np.uint32)
        Fs, Cs, Vs = np.zeros(NP), np.zeros(NP), np.zeros(NP, dtype=np.uint32)
        done = C.c_uint()
        code = {"de": 0, "sade": 1, "de1220": 2}[algo]
        fn = self.lib.oracle_de_evolve_sequential if sequential else self.lib.oracle_de_evolve
        rc = fn(C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), C.c_size_t(NP), C.c_size_t(dim), C.c_uint(gens),
                                       C.c_uint(code), C.c_uint(variant), C.c_uint(variant_adptv), C.c_double(F), C.c_double(CR),
                                       al.ctypes.data_as(C.POINTER(C.c_uint)), C.c_uint(al.size), C.c_double(ftol), C.c_double(xtol),
                                       C.c_uint64(seed), C.c_uint32(first_generation), C.byref(done), _dp(Fs), _dp(Cs),
                                       Vs.ctypes.data_as(C.POINTER(C.c_uint)))
        if rc:
            raise ValueError("oracle_de_evolve failed")
        return x, f, done.value, Fs, Cs, Vs

    def sga_evolve(self, prob, lb, ub, x, f, gens=1, cr=0.9, eta_c=1.0, m=0.02, param_m=1.0, param_s=2, crossover="exponential",
                   mutation="polynomial", selection="tournament", seed=0, first_generation=1):
        """restated generational sga: returns (x, f)."""
        x = np.array(x, dtype=np.float64, order="C")
        f = np.array(f, dtype=np.float64, order="C").reshape(-1)
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        xo = {"exponential": 0, "binomial": 1, "single": 2, "sbx": 3}[crossover]
        mu = {"gaussian": 0, "uniform": 1, "polynomial": 2}[mutation]
        se = {"tournament": 0, "truncated": 1}[selection]
        if self.lib.oracle_sga_evolve(C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), C.c_size_t(x.shape[0]), C.c_size_t(x.shape[1]),
                                      C.c_uint(gens), C.c_double(cr), C.c_double(eta_c), C.c_double(m), C.c_double(param_m), C.c_uint(param_s),
                                      C.c_uint(xo), C.c_uint(mu), C.c_uint(se), C.c_uint64(seed), C.c_uint32(first_generation)):
            raise ValueError("oracle_sga_evolve failed")
        return x, f

    def cmaes_evolve(self, prob, lb, ub, x, f, gens=1, cc=-1., cs=-1., c1=-1., cmu=-1., sigma0=0.5, ftol=1e-6, xtol=1e-6, force_bounds=False,
                     seed=0, first_generation=1):
        """restated cmaes::evolve (memory = false) on the Philox normals: returns (x, f, gens_done, sigma)."""
        x = np.array(x, dtype=np.float64, order="C")
        f = np.array(f, dtype=np.float64, order="C").reshape(-1)
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        done, sigma = C.c_uint(), C.c_double()
        if self.lib.oracle_cmaes_evolve(C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), C.c_size_t(x.shape[0]), C.c_size_t(x.shape[1]),
                                        C.c_uint(gens), C.c_double(cc), C.c_double(cs), C.c_double(c1), C.c_double(cmu), C.c_double(sigma0),
                                        C.c_double(ftol), C.c_double(xtol), C.c_int(int(force_bounds)), C.c_uint64(seed),
                                        C.c_uint32(first_generation), C.byref(done), C.byref(sigma)):
            raise ValueError("oracle_cmaes_evolve failed")
        return x, f, done.value, sigma.value

    def xnes_evolve(self, prob, lb, ub, x, f, gens=1, eta_mu=-1., eta_sigma=-1., eta_b=-1., sigma0=-1., ftol=1e-6, xtol=1e-6, force_bounds=False,
                    seed=0, first_generation=1):
        """restated xnes::evolve (memory = false) on the Philox normals: returns (x, f, gens_done, sigma)."""
        x = np.array(x, dtype=np.float64, order="C")
        f = np.array(f, dtype=np.float64, order="C").reshape(-1)
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        done, sigma = C.c_uint(), C.c_double()
        if self.lib.oracle_xnes_evolve(C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), C.c_size_t(x.shape[0]), C.c_size_t(x.shape[1]),
                                       C.c_uint(gens), C.c_double(eta_mu), C.c_double(eta_sigma), C.c_double(eta_b), C.c_double(sigma0),
                                       C.c_double(ftol), C.c_double(xtol), C.c_int(int(force_bounds)), C.c_uint64(seed),
                                       C.c_uint32(first_generation), C.byref(done), C.byref(sigma)):
            raise ValueError("oracle_xnes_evolve failed")
        return x, f, done.value, sigma.value

    # ---- CMA-ES / xNES contractions (restate_cmaes.c) ----
    def weighted_gram(self, rows, w, idx=None, center=None, scale_div: float = 1.0):
        """(sum_i w_i (r_i - c)(r_i - c)^T / scale_div, sum_i w_i r_i) in the reference's order."""
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        D = rows.shape[1]
        ip = np.ascontiguousarray(idx, dtype=np.uint32).ctypes.data_as(C.POINTER(C.c_uint32)) if idx is not None else None
        cp = _dp(np.ascontiguousarray(center, dtype=np.float64)) if center is not None else None
        g, m = np.empty((D, D)), np.empty(D)
        self.lib.oracle_weighted_gram(_dp(rows), ip, cp, _dp(w), C.c_size_t(w.size), C.c_size_t(D), C.c_double(scale_div), _dp(g))
        self.lib.oracle_weighted_mean(_dp(rows), ip, _dp(w), C.c_size_t(w.size), C.c_size_t(D), _dp(m))
        return g, m

    def cmaes_sample(self, mean, bd, sigma, lam, seed, generation):
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        bd = np.ascontiguousarray(bd, dtype=np.float64)
        D = mean.size
        z, x = np.empty((lam, D)), np.empty((lam, D))
        self.lib.oracle_cmaes_sample(_dp(mean), _dp(bd), C.c_double(sigma), C.c_size_t(lam), C.c_size_t(D), C.c_uint64(seed),
                                     C.c_uint32(generation), _dp(z), _dp(x))
        return x, z

    # ---- hypervolume (restate_hv.c) ----
    def hv_compute(self, f, r) -> float:
        f = np.ascontiguousarray(f, dtype=np.float64)
        r = np.ascontiguousarray(r, dtype=np.float64)
        out = C.c_double()
        if self.lib.oracle_hv_compute(_dp(f), C.c_size_t(f.shape[0]), C.c_size_t(r.size), _dp(r), C.byref(out)):
            raise ValueError("Reference point is invalid, or the dimension is not 2 or 3")
        return out.value

    def hv_contributions(self, f, r) -> np.ndarray:
        f = np.ascontiguousarray(f, dtype=np.float64)
        r = np.ascontiguousarray(r, dtype=np.float64)
        out = np.empty(max(f.shape[0], 1))
        if self.lib.oracle_hv_contributions(_dp(f), C.c_size_t(f.shape[0]), C.c_size_t(r.size), _dp(r), _dp(out)):
            raise ValueError("Reference point is invalid, or the dimension is not 2 or 3")
        return out[:f.shape[0]]

    # ---- migration (restate_migration.c) ----
    @staticmethod
    def _group(ids, x, f, nf=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = x.shape[0]
        f = np.ascontiguousarray(f, dtype=np.float64)
        f = f.reshape(n, -1) if n else f.reshape(0, nf if nf is not None else (f.shape[1] if f.ndim == 2 else 1))
        return np.ascontiguousarray(ids, dtype=np.uint64), x, f

    def select_best(self, ids, x, f, rate):
        """select_best{rate}.select: (ids, x, f) of the selected individuals; float rate = fractional."""
        ids, x, f = self._group(ids, x, f)
        n, nx, nf = x.shape[0], x.shape[1], f.shape[1]
        io, xo, fo = np.empty(max(n, 1), dtype=np.uint64), np.empty((max(n, 1), nx)), np.empty((max(n, 1), nf))
        k = C.c_size_t()
        u64p = C.POINTER(C.c_uint64)
        if self.lib.oracle_select_best(ids.ctypes.data_as(u64p), _dp(x), _dp(f), C.c_size_t(n), C.c_size_t(nx), C.c_size_t(nf),
                                       C.c_int(isinstance(rate, float)), C.c_double(rate), io.ctypes.data_as(u64p), _dp(xo), _dp(fo), C.byref(k)):
            raise ValueError("oracle_select_best: invalid migration rate")
        return io[:k.value], xo[:k.value], fo[:k.value]

    def fair_replace(self, ids, x, f, rate, mids, mx, mf):
        """fair_replace{rate}.replace: the new (ids, x, f) of the island."""
        ids, x, f = self._group(ids, x, f)
        n, nx, nf = x.shape[0], x.shape[1], f.shape[1]
        mids, mx, mf = self._group(mids, np.asarray(mx, dtype=np.float64).reshape(-1, nx), mf, nf)
        io, xo, fo = np.empty(max(n, 1), dtype=np.uint64), np.empty((max(n, 1), nx)), np.empty((max(n, 1), nf))
        u64p = C.POINTER(C.c_uint64)
        if self.lib.oracle_fair_replace(ids.ctypes.data_as(u64p), _dp(x), _dp(f), C.c_size_t(n), C.c_size_t(nx), C.c_size_t(nf),
                                        C.c_int(isinstance(rate, float)), C.c_double(rate), mids.ctypes.data_as(u64p), _dp(mx), _dp(mf),
                                        C.c_size_t(mx.shape[0]), io.ctypes.data_as(u64p), _dp(xo), _dp(fo)):
            raise ValueError("oracle_fair_replace: invalid migration rate")
        return io[:n], xo[:n], fo[:n]

    def sort_population_con(self, f, nec, nic, tol) -> np.ndarray:
        f = np.ascontiguousarray(f, dtype=np.float64)
        tol = np.ascontiguousarray(tol, dtype=np.float64)
        out = np.empty(max(f.shape[0], 1), dtype=np.uint64)
        self.lib.oracle_sort_population_con(_dp(f), C.c_size_t(f.shape[0]), C.c_size_t(nec), C.c_size_t(nic), _dp(tol), _sp(out))
        return out[: f.shape[0]].astype(np.int64)

    def select_best_con(self, ids, x, f, rate, nec, nic, tol):
        """the constrained branch of select_best::select (select_best.cpp:137-152): rows of f are [objective | nec eq | nic ineq]."""
        ids, x, f = self._group(ids, x, f)
        tol = np.ascontiguousarray(tol, dtype=np.float64)
        n, nx, nf = x.shape[0], x.shape[1], f.shape[1]
        io, xo, fo = np.empty(max(n, 1), dtype=np.uint64), np.empty((max(n, 1), nx)), np.empty((max(n, 1), nf))
        k = C.c_size_t()
        u64p = C.POINTER(C.c_uint64)
        if self.lib.oracle_select_best_con(ids.ctypes.data_as(u64p), _dp(x), _dp(f), C.c_size_t(n), C.c_size_t(nx), C.c_size_t(nec), C.c_size_t(nic),
                                           _dp(tol), C.c_int(isinstance(rate, float)), C.c_double(rate), io.ctypes.data_as(u64p), _dp(xo), _dp(fo),
                                           C.byref(k)):
            raise ValueError("oracle_select_best_con: invalid migration rate")
        return io[:k.value], xo[:k.value], fo[:k.value]

    def fair_replace_con(self, ids, x, f, rate, mids, mx, mf, nec, nic, tol):
        """the constrained branch of fair_replace::replace (fair_replace.cpp:158-188)."""
        ids, x, f = self._group(ids, x, f)
        tol = np.ascontiguousarray(tol, dtype=np.float64)
        n, nx, nf = x.shape[0], x.shape[1], f.shape[1]
        mids, mx, mf = self._group(mids, np.asarray(mx, dtype=np.float64).reshape(-1, nx), mf, nf)
        io, xo, fo = np.empty(max(n, 1), dtype=np.uint64), np.empty((max(n, 1), nx)), np.empty((max(n, 1), nf))
        u64p = C.POINTER(C.c_uint64)
        if self.lib.oracle_fair_replace_con(ids.ctypes.data_as(u64p), _dp(x), _dp(f), C.c_size_t(n), C.c_size_t(nx), C.c_size_t(nec), C.c_size_t(nic),
                                            _dp(tol), C.c_int(isinstance(rate, float)), C.c_double(rate), mids.ctypes.data_as(u64p), _dp(mx),
                                            _dp(mf), C.c_size_t(mx.shape[0]), io.ctypes.data_as(u64p), _dp(xo), _dp(fo)):
            raise ValueError("oracle_fair_replace_con: invalid migration rate")
        return io[:n], xo[:n], fo[:n]

    def connections(self, kind: str, n: int, i: int) -> np.ndarray:
        out = np.empty(max(n, 1), dtype=np.uint64)
        cnt = C.c_size_t()
        fn = {"ring": self.lib.oracle_ring_connections, "fully_connected": self.lib.oracle_fully_connected_connections}[kind]
        if fn(C.c_size_t(n), C.c_size_t(i), out.ctypes.data_as(C.POINTER(C.c_size_t)), C.byref(cnt)):
            raise ValueError("invalid vertex index")
        return out[:cnt.value].astype(np.int64)

    def population_init(self, lb, ub, n: int, seed: int):
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        x, ids = np.empty((n, lb.size)), np.empty(n, dtype=np.uint64)
        if self.lib.oracle_population_init(_dp(lb), _dp(ub), C.c_size_t(n), C.c_size_t(lb.size), C.c_uint64(seed), _dp(x),
                                           ids.ctypes.data_as(C.POINTER(C.c_uint64))):
            raise ValueError("Cannot generate a random real if the bounds are not finite")
        return x, ids

    # ---- Philox draws and NSGA-II operators (restate_nsga2.c) ----
    def philox_raw(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self.lib.oracle_philox_raw(c, k, o)
        return list(o)

    def philox_u01(self, seed, tag, generation, index, slot) -> float:
        self.lib.oracle_philox_u01_at.restype = C.c_double
        return self.lib.oracle_philox_u01_at(C.c_uint64(seed), C.c_uint32(tag), C.c_uint32(generation), C.c_uint32(index), C.c_uint32(slot))

    def philox_perm(self, n, seed, tag, generation) -> np.ndarray:
        out = np.empty(max(n, 1), dtype=np.uint64)
        self.lib.oracle_philox_perm(C.c_size_t(n), C.c_uint64(seed), C.c_uint32(tag), C.c_uint32(generation), _sp(out))
        return out[:n].astype(np.int64)

    def nsga2_rank_crowding(self, f):
        f = np.ascontiguousarray(f, dtype=np.float64)
        rank = np.empty(f.shape[0], dtype=np.uint64)
        cd = np.empty(f.shape[0])
        if self.lib.oracle_nsga2_rank_crowding(_dp(f), C.c_size_t(f.shape[0]), C.c_size_t(f.shape[1]), _sp(rank), _dp(cd)):
            raise ValueError("oracle_nsga2_rank_crowding failed")
        return rank.astype(np.int64), cd

    def nsga2_variation(self, x, rank, cd, lb, ub, sh1, sh2, cr, eta_c, m, eta_m, seed, generation):
        x = np.ascontiguousarray(x, dtype=np.float64)
        NP, nx = x.shape
        r = np.ascontiguousarray(rank, dtype=np.uint64)
        s1, s2 = np.ascontiguousarray(sh1, dtype=np.uint64), np.ascontiguousarray(sh2, dtype=np.uint64)
        cd, lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (cd, lb, ub))
        out = np.empty((NP, nx))
        if self.lib.oracle_nsga2_variation(_dp(x), _sp(r), _dp(cd), C.c_size_t(NP), C.c_size_t(nx), _dp(lb), _dp(ub), _sp(s1), _sp(s2),
                                           C.c_double(cr), C.c_double(eta_c), C.c_double(m), C.c_double(eta_m), C.c_uint64(seed),
                                           C.c_uint32(generation), _dp(out)):
            raise ValueError("oracle_nsga2_variation failed")
        return out

    def nsga2_evolve(self, family, prob_id, nobj, alpha, lb, ub, x, f, gens, cr, eta_c, m, eta_m, seed, first_generation=0):
        x = np.array(x, dtype=np.float64, order="C")
        f = np.array(f, dtype=np.float64, order="C")
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        fam = {"zdt": 8, "dtlz": 9}[family]
        if self.lib.oracle_nsga2_evolve(C.c_int(fam), C.c_uint(prob_id), C.c_size_t(x.shape[1]), C.c_size_t(nobj), C.c_uint(alpha), _dp(lb),
                                        _dp(ub), _dp(x), _dp(f), C.c_size_t(x.shape[0]), C.c_uint(gens), C.c_double(cr), C.c_double(eta_c),
                                        C.c_double(m), C.c_double(eta_m), C.c_uint64(seed), C.c_uint32(first_generation)):
            raise ValueError("oracle_nsga2_evolve failed")
        return x, f

    # ---- sequential-mt19937 mode: the restatements on the reference's own draw stream (mt19937.h) ----
    def mt_sequence(self, seed: int, kind: str, n: int, a: int = 0, b: int = 0):
        k = {"raw": 0, "u01": 1, "int": 2, "normal": 3, "real": 4}[kind]
        r, i = np.empty(n), np.empty(n, dtype=np.uint64)
        if self.lib.oracle_mt_sequence(C.c_uint32(seed), C.c_int(k), C.c_uint64(a), C.c_uint64(b), C.c_size_t(n), _dp(r),
                                       i.ctypes.data_as(C.POINTER(C.c_uint64))):
            raise ValueError("oracle_mt_sequence failed")
        return i if k in (0, 2) else r

    def set_sort_mode(self, libstdcxx: bool) -> None:
        """True: index sorts leave ties as libstdc++'s std::sort does (= the compiled reference); False (default): stable."""
        self.lib.oracle_set_sort_mode(C.c_int(1 if libstdcxx else 0))

    def std_argsort(self, keys, desc: bool = False) -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=np.float64)
        out = np.empty(max(keys.size, 1), dtype=np.uint64)
        self.lib.oracle_std_argsort(_dp(keys), C.c_size_t(keys.size), C.c_int(int(desc)), _sp(out))
        return out[:keys.size].astype(np.int64)

    def mt_shuffles(self, seed: int, n: int, rounds: int = 1) -> np.ndarray:
        out = np.empty(max(n, 1), dtype=np.uint64)
        self.lib.oracle_mt_shuffles(C.c_uint32(seed), C.c_size_t(n), C.c_size_t(rounds), _sp(out))
        return out[:n].astype(np.int64)

    def mt_binomial(self, seed: int, t: int, p: float, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint64)
        if self.lib.oracle_mt_binomial_sequence(C.c_uint32(seed), C.c_uint64(t), C.c_double(p), C.c_size_t(n),
                                                out.ctypes.data_as(C.POINTER(C.c_uint64))):
            raise ValueError("oracle_mt_binomial_sequence: t*p >= 8 is not restated")
        return out

    def genetic_operators_mt(self, p1, p2, lb, ub, p_cr, eta_c, p_m, eta_m, rank, cd, seed):
        p1, p2, lb, ub, cd = (np.ascontiguousarray(a, dtype=np.float64) for a in (p1, p2, lb, ub, cd))
        rank = np.ascontiguousarray(rank, dtype=np.uint64)
        npairs = rank.size // 2
        c1, c2, w = np.empty_like(p1), np.empty_like(p1), np.empty(max(npairs, 1), dtype=np.uint64)
        self.lib.oracle_genetic_operators_mt(_dp(p1), _dp(p2), C.c_size_t(p1.size), _dp(lb), _dp(ub), C.c_double(p_cr), C.c_double(eta_c),
                                             C.c_double(p_m), C.c_double(eta_m), _sp(rank), _dp(cd), C.c_size_t(npairs), C.c_uint32(seed),
                                             _dp(c1), _dp(c2), _sp(w))
        return c1, c2, w[:npairs].astype(np.int64)

    def nsga2_evolve_mt(self, family, prob_id, nobj, alpha, lb, ub, x, f, gens, cr, eta_c, m, eta_m, seed):
        x = np.array(x, dtype=np.float64, order="C")
        f = np.array(f, dtype=np.float64, order="C")
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        fam = {"zdt": 8, "dtlz": 9}[family]
        if self.lib.oracle_nsga2_evolve_mt(C.c_int(fam), C.c_uint(prob_id), C.c_size_t(x.shape[1]), C.c_size_t(nobj), C.c_uint(alpha), _dp(lb),
                                           _dp(ub), _dp(x), _dp(f), C.c_size_t(x.shape[0]), C.c_uint(gens), C.c_double(cr), C.c_double(eta_c),
                                           C.c_double(m), C.c_double(eta_m), C.c_uint32(seed)):
            raise ValueError("oracle_nsga2_evolve_mt failed")
        return x, f

    def pso_evolve_mt(self, prob, lb, ub, x, f, gens=1, omega=0.7298, eta1=2.05, eta2=2.05, max_vel=0.5, variant=5, neighb_type=2,
                      neighb_param=4, seed=0):
        """restated pso_gen::evolve on the mt19937 stream: returns (lbX, lbfit)."""
        x = np.array(x, dtype=np.float64, order="C")
        f = np.array(f, dtype=np.float64, order="C").reshape(-1)
        n, dim = x.shape
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        if self.lib.oracle_pso_evolve_mt(C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), C.c_size_t(n), C.c_size_t(dim), C.c_uint(gens),
                                         C.c_double(omega), C.c_double(eta1), C.c_double(eta2), C.c_double(max_vel), C.c_uint(variant),
                                         C.c_uint(neighb_type), C.c_uint(neighb_param), C.c_uint32(seed)):
            raise ValueError("oracle_pso_evolve_mt failed")
        return x, f

    def set_nix(self, nix: int):
        """integer alleles at the end of the chromosome for the NSGA-II operators of this thread (problem::get_nix())."""
        self.lib.oracle_nsga2_set_nix.argtypes = [C.c_size_t]
        self.lib.oracle_nsga2_set_nix.restype = None
        self.lib.oracle_nsga2_set_nix(nix)

    DIVERSITY = {"crowding distance": 0, "niche count": 1, "max min": 2}

    def nspso_evolve(self, prob, lb, ub, x, f, gens=1, omega=0.6, c1=2.0, c2=2.0, chi=1.0, v_coeff=0.5, leader_selection_range=60,
                     diversity="crowding distance", seed=0, first_generation=1, vel=None, best_x=None, best_f=None, mt=False):
        """restated nspso::evolve (Philox draws, or the mt19937 stream with mt=True): returns (x, f, vel, best_x, best_f); vel / best_*
        are None unless the memory arrays were passed in."""
        x = np.array(x, dtype=np.float64, order="C")
        f = np.array(f, dtype=np.float64, order="C").reshape(x.shape[0], -1)
        n, dim = x.shape
        m = f.shape[1]
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        div = self.DIVERSITY[diversity]
        if mt:
            self.lib.oracle_nspso_evolve_mt.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, C.c_size_t, C.c_size_t,
                                                        C.c_size_t, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                                        C.c_uint, C.c_uint, C.c_uint32]
            rc = self.lib.oracle_nspso_evolve_mt(C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), n, dim, m, gens, omega, c1, c2, chi, v_coeff,
                                                 leader_selection_range, div, seed)
            if rc:
                raise ValueError("oracle_nspso_evolve_mt failed")
            return x, f, None, None, None
        mem = [None if a is None else np.array(a, dtype=np.float64, order="C") for a in (vel, best_x, best_f)]
        self.lib.oracle_nspso_evolve.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                                 C.c_uint, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint, C.c_uint,
                                                 C.c_uint64, C.c_uint32, c_double_p, c_double_p, c_double_p]
        rc = self.lib.oracle_nspso_evolve(C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), n, dim, m, gens, omega, c1, c2, chi, v_coeff,
                                          leader_selection_range, div, seed, first_generation, *[None if a is None else _dp(a) for a in mem])
        if rc:
            raise ValueError("oracle_nspso_evolve failed")
        return (x, f, *mem)

    class GacoState(C.Structure):
        _fields_ = [("oracle", C.c_double), ("q", C.c_double), ("n_evalstop", C.c_uint), ("n_impstop", C.c_uint), ("gen_mark", C.c_uint),
                    ("fevals", C.c_ulonglong), ("counter", C.c_uint), ("memory", C.c_int), ("archive", C.c_void_p),
                    ("has_champion", C.c_int), ("champion", C.c_double)]

    def gaco_evolve(self, prob, lb, ub, x, f, nix=0, gens=1, ker=63, q=1.0, oracle=0.0, acc=0.01, threshold=1, n_gen_mark=7, impstop=100000,
                    evalstop=100000, focus=0.0, seed=0, first_generation=1, mt=False, state=None, memory=False, calls=1):
        """restated gaco::evolve (Philox draws, or the mt19937 stream with mt=True): returns (x, f, state, gens_done); `state` = the
        scalar members that survive between evolve() calls (pass it back in to continue with the same algorithm object)."""
        x = np.array(x, dtype=np.float64, order="C")
        f = np.array(f, dtype=np.float64, order="C").reshape(-1)
        n, nx = x.shape
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        if mt:
            self.lib.oracle_gaco_evolve_mt.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, C.c_size_t, C.c_size_t,
                                                       C.c_size_t, C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_uint, C.c_uint,
                                                       C.c_uint, C.c_uint, C.c_double, C.c_uint32, C.c_int, C.c_uint]
            if self.lib.oracle_gaco_evolve_mt(C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), n, nx, nix, gens, ker, q, oracle, acc, threshold,
                                              n_gen_mark, impstop, evalstop, focus, seed, int(memory), calls):
                raise ValueError("oracle_gaco_evolve_mt failed")
            return x, f, None, gens
        st = state if state is not None else self.GacoState()
        if state is None:
            self.lib.oracle_gaco_state_init.argtypes = [C.c_void_p, C.c_double, C.c_double]
            self.lib.oracle_gaco_state_init.restype = None
            self.lib.oracle_gaco_state_init(C.byref(st), q, oracle)
            if memory:  # the archive lives as long as the state object does
                st.memory = 1
                st._archive = np.zeros(ker * (nx + 2))
                st.archive = st._archive.ctypes.data
        done = C.c_uint()
        self.lib.oracle_gaco_evolve.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                                C.c_uint, C.c_uint, C.c_double, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_uint64,
                                                C.c_uint32, C.c_void_p, C.POINTER(C.c_uint)]
        if self.lib.oracle_gaco_evolve(C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), n, nx, nix, gens, ker, acc, threshold, n_gen_mark,
                                       impstop, evalstop, focus, seed, first_generation, C.byref(st), C.byref(done)):
            raise ValueError("oracle_gaco_evolve failed")
        return x, f, st, done.value

    class MacoState(C.Structure):
        _fields_ = [("q", C.c_double), ("n_evalstop", C.c_uint), ("gen_mark", C.c_uint)]

    def maco_evolve(self, prob, lb, ub, x, f, nix=0, gens=1, ker=63, q=1.0, threshold=1, n_gen_mark=7, evalstop=100000, focus=0.0, seed=0,
                    first_generation=1, mt=False, state=None):
        """restated maco::evolve (Philox draws, or the mt19937 stream with mt=True): returns (x, f, state, gens_done)."""
        x = np.array(x, dtype=np.float64, order="C")
        f = np.array(f, dtype=np.float64, order="C").reshape(x.shape[0], -1)
        n, nx = x.shape
        m = f.shape[1]
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        if mt:
            self.lib.oracle_maco_evolve_mt.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, C.c_size_t, C.c_size_t,
                                                       C.c_size_t, C.c_size_t, C.c_uint, C.c_uint, C.c_double, C.c_uint, C.c_uint, C.c_uint,
                                                       C.c_double, C.c_uint32]
            if self.lib.oracle_maco_evolve_mt(C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), n, nx, nix, m, gens, ker, q, threshold, n_gen_mark,
                                              evalstop, focus, seed):
                raise ValueError("oracle_maco_evolve_mt failed")
            return x, f, None, gens
        st = state if state is not None else self.MacoState()
        if state is None:
            self.lib.oracle_maco_state_init.argtypes = [C.c_void_p, C.c_double]
            self.lib.oracle_maco_state_init.restype = None
            self.lib.oracle_maco_state_init(C.byref(st), q)
        done = C.c_uint()
        self.lib.oracle_maco_evolve.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                                C.c_size_t, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_uint64, C.c_uint32,
                                                C.c_void_p, C.POINTER(C.c_uint)]
        if self.lib.oracle_maco_evolve(C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), n, nx, nix, m, gens, ker, threshold, n_gen_mark, evalstop,
                                       focus, seed, first_generation, C.byref(st), C.byref(done)):
            raise ValueError("oracle_maco_evolve failed")
        return x, f, st, done.value

    DECOMPOSITION = {"weighted": 0, "tchebycheff": 1, "bi": 2}

    def moead_gen_evolve(self, prob, lb, ub, x, f, weights, neigh, gens=1, decomposition="tchebycheff", CR=1.0, F=0.5, eta_m=20.0, realb=0.9,
                         limit=2, preserve_diversity=True, seed=0, first_generation=1, mt=False, burn_draws=0):
        """restated moead_gen::evolve (Philox draws, or the mt19937 stream with mt=True) on given weight vectors [NP x m] and
        neighbourhoods [NP x T]: returns (x, f)."""
        x = np.array(x, dtype=np.float64, order="C")
        f = np.array(f, dtype=np.float64, order="C").reshape(x.shape[0], -1)
        NP, dim = x.shape
        m = f.shape[1]
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        w = np.ascontiguousarray(weights, dtype=np.float64)
        nb = np.ascontiguousarray(neigh, dtype=np.uint64)
        T = nb.shape[1]
        common = [C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), C.c_size_t(NP), C.c_size_t(dim), C.c_size_t(m), C.c_uint(gens), _dp(w),
                  nb.ctypes.data_as(C.POINTER(C.c_size_t)), C.c_size_t(T), C.c_int(self.DECOMPOSITION[decomposition]), C.c_double(CR),
                  C.c_double(F), C.c_double(eta_m), C.c_double(realb), C.c_uint(limit), C.c_int(1 if preserve_diversity else 0)]
        if mt:
            rc = self.lib.oracle_moead_gen_evolve_mt(*common, C.c_uint32(seed), C.c_size_t(burn_draws))
        else:
            rc = self.lib.oracle_moead_gen_evolve(*common, C.c_uint64(seed), C.c_uint32(first_generation), C.c_size_t(0))
        if rc:
            raise ValueError("oracle_moead_gen_evolve failed")
        return x, f

    def de_evolve_mt(self, prob, lb, ub, x, f, gens=1, algo="de1220", variant=2, variant_adptv=1, F=0.8, CR=0.9,
                     allowed=(2, 3, 7, 10, 13, 14, 15, 16), ftol=1e-6, xtol=1e-6, seed=0):
        """restated de / sade / de1220 on the mt19937 stream, in the reference's one-at-a-time order: returns (x, f, gens_done)."""
        x = np.array(x, dtype=np.float64, order="C")
        f = np.array(f, dtype=np.float64, order="C").reshape(-1)
        NP, dim = x.shape
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        al = np.ascontiguousarray(allowed, dtype=np.uint32)
        done = C.c_uint()
        code = {"de": 0, "sade": 1, "de1220": 2}[algo]
        rc = self.lib.oracle_de_evolve_mt(C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), C.c_size_t(NP), C.c_size_t(dim), C.c_uint(gens),
                                          C.c_uint(code), C.c_uint(variant), C.c_uint(variant_adptv), C.c_double(F), C.c_double(CR),
                                          al.ctypes.data_as(C.POINTER(C.c_uint)), C.c_uint(al.size), C.c_double(ftol), C.c_double(xtol),
                                          C.c_uint32(seed), C.byref(done))
        if rc:
            raise ValueError("oracle_de_evolve_mt failed")
        return x, f, done.value

    def sga_evolve_mt(self, prob, lb, ub, x, f, gens=1, cr=0.9, eta_c=1.0, m=0.02, param_m=1.0, param_s=2, crossover="exponential",
                      mutation="polynomial", selection="tournament", seed=0):
        x = np.array(x, dtype=np.float64, order="C")
        f = np.array(f, dtype=np.float64, order="C").reshape(-1)
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        xo = {"exponential": 0, "binomial": 1, "single": 2, "sbx": 3}[crossover]
        mu = {"gaussian": 0, "uniform": 1, "polynomial": 2}[mutation]
        se = {"tournament": 0, "truncated": 1}[selection]
        if self.lib.oracle_sga_evolve_mt(C.byref(prob), _dp(lb), _dp(ub), _dp(x), _dp(f), C.c_size_t(x.shape[0]), C.c_size_t(x.shape[1]),
                                         C.c_uint(gens), C.c_double(cr), C.c_double(eta_c), C.c_double(m), C.c_double(param_m), C.c_uint(param_s),
                                         C.c_uint(xo), C.c_uint(mu), C.c_uint(se), C.c_uint32(seed)):
            raise ValueError("oracle_sga_evolve_mt failed")
        return x, f

    def population_init_mt(self, lb, ub, n: int, seed: int):
        """population(prob, n, seed)'s decision vectors and ids on the mt19937 stream (population.cpp:62-80, 570-596)."""
        lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (lb, ub))
        x, ids = np.empty((n, lb.size)), np.empty(n, dtype=np.uint64)
        if self.lib.oracle_population_init_mt(_dp(lb), _dp(ub), C.c_size_t(n), C.c_size_t(lb.size), C.c_uint32(seed), _dp(x),
                                              ids.ctypes.data_as(C.POINTER(C.c_uint64))):
            raise ValueError("oracle_population_init_mt failed")
        return x, ids

    def cec2014(self, func: int, xs: np.ndarray, tables=None, nthreads: int = 1) -> np.ndarray:
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        n, d = xs.shape
        mr, os_, s = tables if tables is not None else self.cec2014_problem_tables(func, d)
        out = np.empty(n)
        rc = self.lib.oracle_cec2014_batch(C.c_uint(func), C.c_uint(d), _dp(mr), _dp(os_), _ip(s), _dp(xs),
                                           C.c_size_t(n), _dp(out), C.c_int(nthreads))
        if rc:
            raise ValueError(f"oracle_cec2014_batch failed rc={rc} (func={func}, dim={d})")
        return out


    def cec2013(self, func: int, xs: np.ndarray, tables=None) -> np.ndarray:
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        n, d = xs.shape
        mr, os_ = tables if tables is not None else self.cec2013_tables(d)
        out = np.empty(n)
        if self.lib.oracle_cec2013_batch(C.c_uint(func), C.c_uint(d), _dp(mr), _dp(os_), _dp(xs), C.c_size_t(n), _dp(out)):
            raise ValueError(f"oracle_cec2013_batch failed (func={func}, dim={d})")
        return out


class RefProblem:
    def __init__(self, ref: "Reference", handle):
        self._ref, self._h = ref, handle
        L = ref.lib
        self.nx = L.ref_problem_nx(handle)
        self.nf = L.ref_problem_nf(handle)
        self.nobj = L.ref_problem_nobj(handle)
        self.nec = L.ref_problem_nec(handle)
        self.nic = L.ref_problem_nic(handle)

    def __del__(self):
        try:
            self._ref.lib.ref_problem_destroy(self._h)
        except Exception:
            pass

    @property
    def name(self) -> str:
        buf = C.create_string_buffer(256)
        self._ref._check(self._ref.lib.ref_problem_name(self._h, buf, C.c_size_t(256)))
        return buf.value.decode()

    @property
    def fevals(self) -> int:
        return self._ref.lib.ref_problem_fevals(self._h)

    def set_c_tol(self, tol) -> None:
        """problem::set_c_tol (problem.cpp:620-644)"""
        tol = np.ascontiguousarray(tol, dtype=np.float64)
        self._ref._check(self._ref.lib.ref_problem_set_c_tol(self._h, _dp(tol), C.c_size_t(tol.size)))

    def bounds(self):
        lb, ub = np.empty(self.nx), np.empty(self.nx)
        self._ref._check(self._ref.lib.ref_problem_bounds(self._h, _dp(lb), _dp(ub)))
        return lb, ub

    def fitness(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = np.empty(self.nf)
        self._ref._check(self._ref.lib.ref_problem_fitness(self._h, _dp(x), _dp(f)))
        return f

    def fitness_loop(self, xs: np.ndarray) -> np.ndarray:
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        n = xs.size // self.nx
        f = np.empty((n, self.nf))
        self._ref._check(self._ref.lib.ref_problem_fitness_loop(self._h, _dp(xs), C.c_size_t(n), _dp(f)))
        return f

    def thread_bfe(self, xs: np.ndarray, nthreads: int = 0) -> np.ndarray:
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        n = xs.size // self.nx
        f = np.empty((n, self.nf))
        self._ref._check(self._ref.lib.ref_thread_bfe(self._h, _dp(xs), C.c_size_t(n), _dp(f), C.c_int(nthreads)))
        return f

    def evolve(self, algo: str, pop_size: int, gens: int, pop_seed: int = 1, algo_seed: int = 2, use_bfe: bool = False,
               nthreads: int = 0):
        """Unmodified reference algorithm on a fresh population: returns (seconds of evolve(), x, f, fevals)."""
        import os
        x = np.empty((pop_size, self.nx))
        f = np.empty((pop_size, self.nf))
        secs = C.c_double()
        fe = C.c_ulonglong()
        if nthreads > 0:
            os.environ["ORACLE_TBB_THREADS"] = str(nthreads)
        self._ref.lib.ref_evolve.argtypes = [C.c_void_p, C.c_char_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_int,
                                             C.POINTER(C.c_double), c_double_p, c_double_p, C.POINTER(C.c_ulonglong)]
        self._ref._check(self._ref.lib.ref_evolve(self._h, algo.encode(), pop_size, gens, pop_seed, algo_seed, int(use_bfe),
                                                  C.byref(secs), _dp(x), _dp(f), C.byref(fe)))
        return secs.value, x, f, fe.value

    def default_bfe(self, xs: np.ndarray, nthreads: int = 0) -> np.ndarray:
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        n = xs.size // self.nx
        f = np.empty((n, self.nf))
        self._ref._check(self._ref.lib.ref_default_bfe(self._h, _dp(xs), C.c_size_t(n), _dp(f), C.c_int(nthreads)))
        return f


class Reference:
    """The unmodified reference behind a C handle API (oracle/ref_capi.h)."""

    @staticmethod
    def available() -> bool:
        return REF_SO.exists()

    def __init__(self):
        if not REF_SO.exists():
            if (REFERENCE_ROOT / "src" / "problem.cpp").exists():
                build(ref=True)
            else:
                raise FileNotFoundError(f"{REF_SO} is missing and /root/reference is not available to build it")
        self.lib = L = C.CDLL(str(REF_SO))
        L.ref_last_error.restype = C.c_char_p
        L.ref_problem_nx.restype = C.c_size_t
        L.ref_problem_nf.restype = C.c_size_t
        L.ref_problem_nobj.restype = C.c_size_t
        L.ref_problem_fevals.restype = C.c_ulonglong
        L.ref_problem_nec.restype = C.c_size_t
        L.ref_problem_nic.restype = C.c_size_t
        L.ref_problem_unconstrain.argtypes = [C.c_void_p, C.c_char_p, c_double_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.ref_problem_set_c_tol.argtypes = [C.c_void_p, c_double_p, C.c_size_t]
        for fn in ("ref_problem_nx", "ref_problem_nf", "ref_problem_nobj", "ref_problem_nec", "ref_problem_nic", "ref_problem_fevals",
                   "ref_problem_destroy"):
            getattr(L, fn).argtypes = [C.c_void_p]
        L.ref_problem_destroy.restype = None
        L.ref_problem_bounds.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.ref_problem_translate.argtypes = [C.c_void_p, c_double_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.ref_problem_decompose.argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_size_t, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        L.ref_decompose_objectives.argtypes = [c_double_p, c_double_p, c_double_p, C.c_size_t, C.c_char_p, C.POINTER(C.c_double)]
        L.ref_problem_name.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.ref_problem_fitness.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.ref_problem_fitness_loop.argtypes = [C.c_void_p, c_double_p, C.c_size_t, c_double_p]
        L.ref_thread_bfe.argtypes = [C.c_void_p, c_double_p, C.c_size_t, c_double_p, C.c_int]
        L.ref_default_bfe.argtypes = [C.c_void_p, c_double_p, C.c_size_t, c_double_p, C.c_int]
        L.ref_cec2014_origin_shift.argtypes = [C.c_void_p, c_double_p, C.c_size_t, c_size_p]

    def _check(self, rc: int):
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())

    def problem(self, family: str, p0=0, p1=0, p2=0, p3=0) -> RefProblem:
        h = C.c_void_p()
        self._check(self.lib.ref_problem_create(family.encode(), C.c_uint(p0), C.c_uint(p1), C.c_uint(p2), C.c_uint(p3),
                                                C.byref(h)))
        return RefProblem(self, h)

    def translate(self, inner: RefProblem, t) -> RefProblem:
        """pagmo::problem{pagmo::translate{inner, t}}"""
        t = np.ascontiguousarray(t, dtype=np.float64)
        h = C.c_void_p()
        self._check(self.lib.ref_problem_translate(inner._h, _dp(t), C.c_size_t(t.size), C.byref(h)))
        return RefProblem(self, h)

    def decompose(self, inner: RefProblem, weight, z, method: str = "weighted", adapt_ideal: bool = False) -> RefProblem:
        """pagmo::problem{pagmo::decompose{inner, weight, z, method, adapt_ideal}}"""
        w = np.ascontiguousarray(weight, dtype=np.float64)
        zz = np.ascontiguousarray(z, dtype=np.float64)
        h = C.c_void_p()
        self._check(self.lib.ref_problem_decompose(inner._h, _dp(w), _dp(zz), C.c_size_t(w.size), method.encode(),
                                                   C.c_int(int(adapt_ideal)), C.byref(h)))
        return RefProblem(self, h)

    def unconstrain(self, inner: RefProblem, method: str = "death penalty", weights=()) -> RefProblem:
        """pagmo::problem{pagmo::unconstrain{inner, method, weights}}"""
        w = np.ascontiguousarray(weights, dtype=np.float64)
        h = C.c_void_p()
        self._check(self.lib.ref_problem_unconstrain(inner._h, method.encode(), _dp(w), C.c_size_t(w.size), C.byref(h)))
        return RefProblem(self, h)

    def decompose_objectives(self, f, weight, z, method: str) -> float:
        f = np.ascontiguousarray(f, dtype=np.float64)
        w = np.ascontiguousarray(weight, dtype=np.float64)
        zz = np.ascontiguousarray(z, dtype=np.float64)
        out = C.c_double()
        self._check(self.lib.ref_decompose_objectives(_dp(f), _dp(w), _dp(zz), C.c_size_t(f.size), method.encode(), C.byref(out)))
        return out.value

    def cec2014_tables(self, func: int, dim: int):
        mr = np.empty(CEC_NCOMP * dim * dim)
        os_ = np.empty(CEC_NCOMP * 100)
        s = np.empty(CEC_NCOMP * dim, dtype=np.int32)
        self._check(self.lib.ref_cec2014_tables(C.c_uint(func), C.c_uint(dim), _dp(mr), _dp(os_), _ip(s)))
        return mr, os_, s

    def cec2013_tables(self, dim: int):
        """(MD[dim], shift_data) as the reference constructors saw them."""
        mr = np.empty(CEC_NCOMP * dim * dim)
        os_ = np.empty(CEC_NCOMP * 100)
        self.lib.ref_cec2013_tables.argtypes = [C.c_uint, c_double_p, c_double_p]
        self._check(self.lib.ref_cec2013_tables(C.c_uint(dim), _dp(mr), _dp(os_)))
        return mr, os_

    # ---- pin entry points (ref_pin.cpp) ----
    def std_sequence(self, seed: int, kind: str, n: int, a: int = 0, b: int = 0):
        k = {"raw": 0, "u01": 1, "int": 2, "normal": 3, "real": 4}[kind]
        r, i = np.empty(n), np.empty(n, dtype=np.uint64)
        self._check(self.lib.ref_std_sequence(C.c_uint(seed), C.c_int(k), C.c_ulonglong(a), C.c_ulonglong(b), C.c_size_t(n), _dp(r),
                                              i.ctypes.data_as(C.POINTER(C.c_ulonglong))))
        return i if k in (0, 2) else r

    def std_argsort(self, keys, desc: bool = False) -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=np.float64)
        out = np.empty(max(keys.size, 1), dtype=np.uint64)
        self._check(self.lib.ref_std_argsort(_dp(keys), C.c_size_t(keys.size), C.c_int(int(desc)), _sp(out)))
        return out[:keys.size].astype(np.int64)

    def std_shuffles(self, seed: int, n: int, rounds: int = 1) -> np.ndarray:
        out = np.empty(max(n, 1), dtype=np.uint64)
        self._check(self.lib.ref_std_shuffles(C.c_uint(seed), C.c_size_t(n), C.c_size_t(rounds), _sp(out)))
        return out[:n].astype(np.int64)

    def std_binomial(self, seed: int, t: int, p: float, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint64)
        self._check(self.lib.ref_std_binomial(C.c_uint(seed), C.c_ulonglong(t), C.c_double(p), C.c_size_t(n),
                                              out.ctypes.data_as(C.POINTER(C.c_ulonglong))))
        return out

    def genetic_operators(self, p1, p2, lb, ub, p_cr, eta_c, p_m, eta_m, rank, cd, seed):
        """sbx_crossover_impl, polynomial_mutation_impl x2, mo_tournament_selection_impl on pairs - one engine."""
        p1, p2, lb, ub, cd = (np.ascontiguousarray(a, dtype=np.float64) for a in (p1, p2, lb, ub, cd))
        rank = np.ascontiguousarray(rank, dtype=np.uint64)
        npairs = rank.size // 2
        c1, c2, w = np.empty_like(p1), np.empty_like(p1), np.empty(max(npairs, 1), dtype=np.uint64)
        self._check(self.lib.ref_genetic_operators(_dp(p1), _dp(p2), C.c_size_t(p1.size), _dp(lb), _dp(ub), C.c_double(p_cr), C.c_double(eta_c),
                                                   C.c_double(p_m), C.c_double(eta_m), _sp(rank), _dp(cd), C.c_size_t(npairs), C.c_uint(seed),
                                                   _dp(c1), _dp(c2), _sp(w)))
        return c1, c2, w[:npairs].astype(np.int64)

    def evolve_from(self, prob: "RefProblem", algo: str, par, x0, gens: int, seed: int, strategies: str | None = None):
        """The unmodified reference UDA on the population with decision vectors x0: returns (x, f)."""
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        par = np.ascontiguousarray(par, dtype=np.float64)
        n = x0.shape[0]
        x, f = np.empty((n, prob.nx)), np.empty((n, prob.nf))
        self.lib.ref_evolve_from.argtypes = [C.c_void_p, C.c_char_p, c_double_p, C.c_size_t, C.c_char_p, c_double_p, C.c_size_t, C.c_uint,
                                             C.c_uint, c_double_p, c_double_p]
        self._check(self.lib.ref_evolve_from(prob._h, algo.encode(), _dp(par), par.size, strategies.encode() if strategies else None, _dp(x0), n,
                                             gens, seed, _dp(x), _dp(f)))
        return x, f

    def decomposition_weights(self, n_f: int, n_w: int, method: str, seed: int = 0) -> np.ndarray:
        """pagmo::decomposition_weights with a fresh std::mt19937(seed)."""
        out = np.empty((n_w, n_f))
        self.lib.ref_decomposition_weights.argtypes = [C.c_size_t, C.c_size_t, C.c_char_p, C.c_uint, c_double_p]
        self._check(self.lib.ref_decomposition_weights(n_f, n_w, method.encode(), seed, _dp(out)))
        return out

    def knn(self, points, k: int) -> np.ndarray:
        """pagmo::kNN: [n x k] indices of the k nearest other points of every point."""
        points = np.ascontiguousarray(points, dtype=np.float64)
        n, m = points.shape
        out = np.empty((n, k), dtype=np.uint64)
        self.lib.ref_knn.argtypes = [c_double_p, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t)]
        self._check(self.lib.ref_knn(_dp(points), n, m, k, out.ctypes.data_as(C.POINTER(C.c_size_t))))
        return out

    def population_init(self, prob: "RefProblem", n: int, seed: int):
        x, ids = np.empty((n, prob.nx)), np.empty(n, dtype=np.uint64)
        self.lib.ref_population_init.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, c_double_p, C.POINTER(C.c_ulonglong)]
        self._check(self.lib.ref_population_init(prob._h, n, seed, _dp(x), ids.ctypes.data_as(C.POINTER(C.c_ulonglong))))
        return x, ids

    def fair_replace(self, ids, x, f, rate, mids, mx, mf):
        """unmodified fair_replace{rate}.replace on flat groups."""
        ids, x, f = Oracle._group(ids, x, f)
        n, nx, nf = x.shape[0], x.shape[1], f.shape[1]
        mids, mx, mf = Oracle._group(mids, np.asarray(mx, dtype=np.float64).reshape(-1, nx), mf, nf)
        io, xo, fo = np.empty(max(n, 1), dtype=np.uint64), np.empty((max(n, 1), nx)), np.empty((max(n, 1), nf))
        u64p = C.POINTER(C.c_ulonglong)
        self.lib.ref_fair_replace.argtypes = [u64p, c_double_p, c_double_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_double, u64p,
                                              c_double_p, c_double_p, C.c_size_t, u64p, c_double_p, c_double_p]
        self._check(self.lib.ref_fair_replace(ids.ctypes.data_as(u64p), _dp(x), _dp(f), n, nx, nf, int(isinstance(rate, float)), float(rate),
                                              mids.ctypes.data_as(u64p), _dp(mx), _dp(mf), mx.shape[0], io.ctypes.data_as(u64p), _dp(xo),
                                              _dp(fo)))
        return io[:n], xo[:n], fo[:n]

    def select_best(self, ids, x, f, rate):
        ids, x, f = Oracle._group(ids, x, f)
        n, nx, nf = x.shape[0], x.shape[1], f.shape[1]
        io, xo, fo = np.empty(max(n, 1), dtype=np.uint64), np.empty((max(n, 1), nx)), np.empty((max(n, 1), nf))
        k = C.c_size_t()
        u64p = C.POINTER(C.c_ulonglong)
        self.lib.ref_select_best.argtypes = [u64p, c_double_p, c_double_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_double, u64p,
                                             c_double_p, c_double_p, C.POINTER(C.c_size_t)]
        self._check(self.lib.ref_select_best(ids.ctypes.data_as(u64p), _dp(x), _dp(f), n, nx, nf, int(isinstance(rate, float)), float(rate),
                                             io.ctypes.data_as(u64p), _dp(xo), _dp(fo), C.byref(k)))
        return io[:k.value], xo[:k.value], fo[:k.value]

    def hv_fpras(self, f, r, eps=1e-2, delta=1e-2, seed=0) -> float:
        f = np.ascontiguousarray(f, dtype=np.float64)
        r = np.ascontiguousarray(r, dtype=np.float64)
        out = C.c_double()
        self.lib.ref_hv_fpras.argtypes = [c_double_p, C.c_size_t, C.c_size_t, c_double_p, C.c_double, C.c_double, C.c_uint, C.POINTER(C.c_double)]
        self._check(self.lib.ref_hv_fpras(_dp(f), f.shape[0], r.size, _dp(r), eps, delta, seed, C.byref(out)))
        return out.value

    def hv_approx_extreme(self, f, r, greatest=False, use_exact=True, eps=1e-2, delta=1e-6, seed=0) -> int:
        f = np.ascontiguousarray(f, dtype=np.float64)
        r = np.ascontiguousarray(r, dtype=np.float64)
        out = C.c_size_t()
        self.lib.ref_hv_approx_extreme.argtypes = [c_double_p, C.c_size_t, C.c_size_t, c_double_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                                   C.c_uint, C.POINTER(C.c_size_t)]
        self._check(self.lib.ref_hv_approx_extreme(_dp(f), f.shape[0], r.size, _dp(r), int(greatest), int(use_exact), eps, delta, seed,
                                                   C.byref(out)))
        return out.value

    def select_best_con(self, ids, x, f, rate, nec, nic, tol):
        ids, x, f = Oracle._group(ids, x, f)
        tol = np.ascontiguousarray(tol, dtype=np.float64)
        n, nx, nf = x.shape[0], x.shape[1], f.shape[1]
        io, xo, fo = np.empty(max(n, 1), dtype=np.uint64), np.empty((max(n, 1), nx)), np.empty((max(n, 1), nf))
        k = C.c_size_t()
        u64p = C.POINTER(C.c_ulonglong)
        self.lib.ref_select_best_con.argtypes = [u64p, c_double_p, c_double_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, c_double_p, C.c_int,
                                                 C.c_double, u64p, c_double_p, c_double_p, C.POINTER(C.c_size_t)]
        self._check(self.lib.ref_select_best_con(ids.ctypes.data_as(u64p), _dp(x), _dp(f), n, nx, nec, nic, _dp(tol), int(isinstance(rate, float)),
                                                 float(rate), io.ctypes.data_as(u64p), _dp(xo), _dp(fo), C.byref(k)))
        return io[:k.value], xo[:k.value], fo[:k.value]

    def fair_replace_con(self, ids, x, f, rate, mids, mx, mf, nec, nic, tol):
        ids, x, f = Oracle._group(ids, x, f)
        tol = np.ascontiguousarray(tol, dtype=np.float64)
        n, nx, nf = x.shape[0], x.shape[1], f.shape[1]
        mids, mx, mf = Oracle._group(mids, np.asarray(mx, dtype=np.float64).reshape(-1, nx), mf, nf)
        io, xo, fo = np.empty(max(n, 1), dtype=np.uint64), np.empty((max(n, 1), nx)), np.empty((max(n, 1), nf))
        u64p = C.POINTER(C.c_ulonglong)
        self.lib.ref_fair_replace_con.argtypes = [u64p, c_double_p, c_double_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, c_double_p, C.c_int,
                                                  C.c_double, u64p, c_double_p, c_double_p, C.c_size_t, u64p, c_double_p, c_double_p]
        self._check(self.lib.ref_fair_replace_con(ids.ctypes.data_as(u64p), _dp(x), _dp(f), n, nx, nec, nic, _dp(tol), int(isinstance(rate, float)),
                                                  float(rate), mids.ctypes.data_as(u64p), _dp(mx), _dp(mf), mx.shape[0], io.ctypes.data_as(u64p),
                                                  _dp(xo), _dp(fo)))
        return io[:n], xo[:n], fo[:n]

    def hv_compute(self, f, r) -> float:
        f = np.ascontiguousarray(f, dtype=np.float64)
        r = np.ascontiguousarray(r, dtype=np.float64)
        out = C.c_double()
        self.lib.ref_hv_compute.argtypes = [c_double_p, C.c_size_t, C.c_size_t, c_double_p, C.POINTER(C.c_double)]
        self._check(self.lib.ref_hv_compute(_dp(f), f.shape[0], r.size, _dp(r), C.byref(out)))
        return out.value

    def hv_contributions(self, f, r) -> np.ndarray:
        f = np.ascontiguousarray(f, dtype=np.float64)
        r = np.ascontiguousarray(r, dtype=np.float64)
        out = np.empty(max(f.shape[0], 1))
        self.lib.ref_hv_contributions.argtypes = [c_double_p, C.c_size_t, C.c_size_t, c_double_p, c_double_p]
        self._check(self.lib.ref_hv_contributions(_dp(f), f.shape[0], r.size, _dp(r), _dp(out)))
        return out[:f.shape[0]]

    def cec2014_origin_shift(self, prob: RefProblem) -> np.ndarray:
        out = np.empty(CEC_NCOMP * 100)
        n = C.c_size_t()
        self._check(self.lib.ref_cec2014_origin_shift(prob._h, _dp(out), C.c_size_t(out.size), C.byref(n)))
        return out[: n.value].copy()

    # ---- multi-objective utilities ----
    def fnds(self, f: np.ndarray, with_dom_list: bool = False):
        f = np.ascontiguousarray(f, dtype=np.float64)
        n, m = f.shape
        rank = np.empty(n, dtype=np.uint64)
        dc = np.empty(n, dtype=np.uint64)
        fidx = np.empty(n, dtype=np.uint64)
        foff = np.empty(n + 1, dtype=np.uint64)
        nfr = C.c_size_t()
        dl_off = np.empty(n + 1, dtype=np.uint64) if with_dom_list else None
        dl_cap = n * n if with_dom_list else 0
        dl_idx = np.empty(dl_cap, dtype=np.uint64) if with_dom_list else None
        self._check(self.lib.ref_fnds(_dp(f), C.c_size_t(n), C.c_size_t(m), _sp(rank), _sp(dc), _sp(fidx), _sp(foff),
                                      C.byref(nfr), _sp(dl_idx) if with_dom_list else None,
                                      _sp(dl_off) if with_dom_list else None, C.c_size_t(dl_cap)))
        k = nfr.value
        fronts = [fidx[int(foff[i]):int(foff[i + 1])].astype(np.int64) for i in range(k)]
        out = {"rank": rank.astype(np.int64), "dom_count": dc.astype(np.int64), "fronts": fronts}
        if with_dom_list:
            out["dom_list"] = [dl_idx[int(dl_off[i]):int(dl_off[i + 1])].astype(np.int64) for i in range(n)]
        return out

    def crowding_distance(self, f: np.ndarray) -> np.ndarray:
        f = np.ascontiguousarray(f, dtype=np.float64)
        out = np.empty(f.shape[0])
        self._check(self.lib.ref_crowding_distance(_dp(f), C.c_size_t(f.shape[0]), C.c_size_t(f.shape[1]), _dp(out)))
        return out

    def sort_population_mo(self, f: np.ndarray) -> np.ndarray:
        f = np.ascontiguousarray(f, dtype=np.float64)
        out = np.empty(f.shape[0], dtype=np.uint64)
        self._check(self.lib.ref_sort_population_mo(_dp(f), C.c_size_t(f.shape[0]), C.c_size_t(f.shape[1]), _sp(out)))
        return out.astype(np.int64)

    def select_best_N_mo(self, f: np.ndarray, N: int) -> np.ndarray:
        f = np.ascontiguousarray(f, dtype=np.float64)
        out = np.empty(f.shape[0], dtype=np.uint64)
        nout = C.c_size_t()
        self._check(self.lib.ref_select_best_N_mo(_dp(f), C.c_size_t(f.shape[0]), C.c_size_t(f.shape[1]), C.c_size_t(N),
                                                  _sp(out), C.byref(nout)))
        return out[: nout.value].astype(np.int64)


_ORACLE = None
_REFERENCE = None


def oracle() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE


def reference() -> Reference:
    global _REFERENCE
    if _REFERENCE is None:
        _REFERENCE = Reference()
    return _REFERENCE
