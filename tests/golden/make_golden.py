"""Regenerate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libpagmo_ref.so).

Run in the authoring container (needs /root/reference to build oracle/_ref):  python tests/golden/make_golden.py
The fixtures let the oracle restatement and the CUDA path be checked against reference outputs on boxes where the
reference itself cannot be built.  Inputs are seeded; outputs are whatever the reference code computes.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import reference  # noqa: E402

OUT = Path(__file__).resolve().parent


def cec2013_fixture(R):
    # ---- CEC2013: every function at D in {10, 30, 50}: 4 points in the box, 3 near the component-0 shift, the shift, origin
    data = {}
    rng13 = np.random.default_rng(20131)
    for dim in (10, 30, 50):
        _, os13 = R.cec2013_tables(dim)
        for func in range(1, 29):
            p = R.problem("cec2013", func, dim)
            xs = np.vstack([rng13.uniform(-100, 100, (4, dim)), os13[:dim] + rng13.normal(0, 1.0, (3, dim)), os13[None, :dim],
                            np.zeros((1, dim))])
            data[f"x_f{func}_d{dim}"] = xs
            data[f"f_f{func}_d{dim}"] = p.fitness_loop(xs)[:, 0]
    np.savez_compressed(OUT / "cec2013_ref.npz", **data)


WFG_CONFIGS = ((9, 5, 8), (10, 5, 8), (5, 3, 4), (12, 3, 4), (30, 4, 6), (4, 2, 2), (24, 2, 4))


def wfg_fixture(R):
    # ---- WFG1..9: random points in the box, the corners, and the point of the reference's own tests (x = 2, tests/wfg.cpp:75-181)
    data = {}
    rng = np.random.default_rng(20161)
    for pid in range(1, 10):
        for n, m, k in WFG_CONFIGS:
            if pid in (2, 3) and (n - k) % 2:
                continue
            ub = 2.0 * (np.arange(n) + 1)
            xs = np.vstack([rng.uniform(0, 1, (6, n)) * ub, np.zeros((1, n)), ub[None, :], np.full((1, n), 2.0)])
            data[f"x_wfg{pid}_{n}_{m}_{k}"] = xs
            data[f"f_wfg{pid}_{n}_{m}_{k}"] = R.problem("wfg", pid, n, m, k).fitness_loop(xs)
    np.savez_compressed(OUT / "wfg_ref.npz", **data)


def hv_wfg_fixture(R):
    """four and more objectives (hvwfg): (a) the reference's own WFG fixtures (tests/hypervolume_test_data/testcases_list.txt:28,33 and
    the 7-objective file next to them) with the answers they carry; (b) hypervolume and contributions of seeded random sets computed by
    the compiled reference (hvwfg::compute / hvwfg::contributions)."""
    base = Path("/root/reference/tests/hypervolume_test_data/testcases")
    data = {}

    def parse(name, kind, limit):
        tok = (base / name).read_text().split()
        pos = 1
        for case in range(min(int(tok[0]), limit)):
            d, n = int(tok[pos]), int(tok[pos + 1])
            pos += 2
            data[f"{kind}_{name}_{case}_r"] = np.array(tok[pos:pos + d], dtype=np.float64)
            pos += d
            data[f"{kind}_{name}_{case}_p"] = np.array(tok[pos:pos + d * n], dtype=np.float64).reshape(n, d)
            pos += d * n
            na = {"compute": 1, "exclusive": 2}[kind]
            data[f"{kind}_{name}_{case}_a"] = np.array(tok[pos:pos + na], dtype=np.float64)
            pos += na

    parse("c_max_t1_d5_n1024", "compute", 1)
    parse("c_max_t1_d7_n64", "compute", 1)  # not in testcases_list.txt; the answer it carries is good to ~2e-4 only
    for name in ("c_max_t1_d5_n1024", "c_max_t1_d7_n64"):  # what the compiled hvwfg returns for them
        data[f"compute_{name}_0_hv"] = np.array([R.hv_compute(data[f"compute_{name}_0_p"], data[f"compute_{name}_0_r"])])
    parse("e_max_d5", "exclusive", 10**6)
    rng = np.random.default_rng(20172)
    for m in (4, 5, 6):
        for n, kind in ((1, "random"), (2, "random"), (3, "random"), (37, "random"), (150, "random"), (100, "front"), (80, "dups")):
            f = rng.uniform(0, 1, (n, m))
            if kind == "front":
                f = f / np.linalg.norm(f, axis=1, keepdims=True)
            if kind == "dups":
                f[1] = f[0]
                f[2] = f[0] + 0.1
                f[5:20, m - 1] = f[4, m - 1]
                f[20:30, 0] = f[19, 0]
            r = np.full(m, 1.3)
            key = f"ref_m{m}_n{n}_{kind}"
            data[key + "_p"], data[key + "_r"] = f, r
            data[key + "_hv"] = np.array([R.hv_compute(f, r)])
            data[key + "_c"] = R.hv_contributions(f, r)
    np.savez_compressed(OUT / "hv_wfg_ref.npz", **data)


def hv_fixture(R):
    """(a) the reference's own hypervolume fixtures (tests/hypervolume_test_data, format: tests/hypervolume.cpp:132-163): the first fronts
    of the 2D / 3D `compute` files, and the complete `exclusive` and `least_contributor` files for 2 and 3 objectives, with the answers
    they carry; (b) contributions of seeded random fronts (with dominated points and duplicates) computed by the compiled reference."""
    base = Path("/root/reference/tests/hypervolume_test_data/testcases")
    data = {}

    def parse(name, kind, limit):
        tok = (base / name).read_text().split()
        pos = 0

        def take(k):
            nonlocal pos
            out = tok[pos:pos + k]
            pos += k
            return out
        t = int(take(1)[0])
        for case in range(min(t, limit)):
            d, n = (int(v) for v in take(2))
            r = np.array(take(d), dtype=np.float64)
            pts = np.array(take(d * n), dtype=np.float64).reshape(n, d)
            ans = [float(v) for v in take({"compute": 1, "exclusive": 2, "least": 1}[kind])]
            data[f"{kind}_{name}_{case}_r"] = r
            data[f"{kind}_{name}_{case}_p"] = pts
            data[f"{kind}_{name}_{case}_a"] = np.array(ans)

    parse("c_max_t100_d2_n128", "compute", 12)
    parse("c_max_t100_d3_n128", "compute", 12)
    parse("c_max_t1_d3_n2048", "compute", 1)
    for name in ("e_max_d2", "e_max_d3"):
        parse(name, "exclusive", 10**6)
    for name in ("lc_max_d2", "lc_max_d3"):
        parse(name, "least", 10**6)
    rng = np.random.default_rng(20171)
    for m in (2, 3):
        for n, kind in ((1, "random"), (2, "random"), (37, "random"), (300, "random"), (300, "front"), (120, "dups")):
            f = rng.uniform(0, 1, (n, m))
            if kind == "front":
                f = f / np.linalg.norm(f, axis=1, keepdims=True)
            if kind == "dups":
                f[1] = f[0]
                f[2] = f[0] + 0.1
                f[5:20, m - 1] = f[4, m - 1]  # ties in the sweep coordinate
            r = np.full(m, 1.3)
            key = f"ref_m{m}_n{n}_{kind}"
            data[key + "_p"], data[key + "_r"] = f, r
            data[key + "_hv"] = np.array([R.hv_compute(f, r)])
            data[key + "_c"] = R.hv_contributions(f, r)
    np.savez_compressed(OUT / "hv_ref.npz", **data)


def constrained_fixture(R):
    """hock_schittkowski_71, luksan_vlcek1 and pagmo::unconstrain of both (five methods, zero and non-zero tolerances)."""
    rng = np.random.default_rng(20171)
    data = {}
    for fam, p0 in (("hock_schittkowski_71", 0), ("luksan_vlcek1", 3), ("luksan_vlcek1", 10), ("luksan_vlcek1", 33)):
        p = R.problem(fam, p0)
        lb, ub = p.bounds()
        xs = rng.uniform(lb, ub, (24, p.nx)) if fam != "luksan_vlcek1" else rng.uniform(-1.5, 1.5, (24, p.nx))
        key = f"{fam}_{p0}"
        nc = p.nec + p.nic
        data[f"x_{key}"] = xs
        data[f"f_{key}"] = p.fitness_loop(xs)
        for ti, tol in enumerate((np.zeros(nc), rng.uniform(1.0, 6.0, nc))):
            p.set_c_tol(tol)
            w = rng.uniform(0.1, 2.0, nc)
            data[f"tol{ti}_{key}"] = tol
            data[f"w{ti}_{key}"] = w
            for method in ("death penalty", "kuri", "weighted", "ignore_c", "ignore_o"):
                u = R.unconstrain(p, method, w if method == "weighted" else ())
                data[f"u{ti}_{method.replace(' ', '_')}_{key}"] = u.fitness_loop(xs)
    np.savez_compressed(OUT / "constrained_ref.npz", **data)


def main():
    R = reference()
    if "--constrained-only" in sys.argv:
        return constrained_fixture(R)
    constrained_fixture(R)
    if "--hv-only" in sys.argv:
        return hv_fixture(R)
    if "--hv-wfg-only" in sys.argv:
        return hv_wfg_fixture(R)
    if "--wfg-only" in sys.argv:  # add one fixture without rewriting the others
        return wfg_fixture(R)
    cec2013_fixture(R)
    if "--cec2013-only" in sys.argv:
        return
    wfg_fixture(R)
    hv_fixture(R)
    hv_wfg_fixture(R)
    rng = np.random.default_rng(20141)
    # ---- CEC2014: every function at D in {10, 30, 100}, 6 points in the box + the shift itself + origin
    data = {}
    for dim in (10, 30, 100):
        for func in range(1, 31):
            p = R.problem("cec2014", func, dim)
            shift = R.cec2014_origin_shift(p)[:dim]
            xs = np.vstack([rng.uniform(-100, 100, (6, dim)), shift[None, :], np.zeros((1, dim))])
            data[f"x_f{func}_d{dim}"] = xs
            data[f"f_f{func}_d{dim}"] = p.fitness_loop(xs)[:, 0]
    np.savez_compressed(OUT / "cec2014_ref.npz", **data)
    # ---- simple UDPs: random points + the known answers of the reference's own tests
    data = {}
    for fam in ("rastrigin", "ackley", "griewank", "schwefel", "rosenbrock"):
        for dim in (2, 5, 10, 100):
            p = R.problem(fam, dim)
            lb, ub = p.bounds()
            xs = np.vstack([rng.uniform(lb, ub, (6, dim)), np.ones((1, dim)), np.zeros((1, dim))])
            data[f"x_{fam}_d{dim}"] = xs
            data[f"f_{fam}_d{dim}"] = p.fitness_loop(xs)[:, 0]
    np.savez_compressed(OUT / "simple_ref.npz", **data)
    # ---- ZDT / DTLZ: random points in the box
    data = {}
    for pid in range(1, 7):
        for param in (2, 11, 30):
            p = R.problem("zdt", pid, param)
            lb, ub = p.bounds()
            xs = rng.uniform(lb, ub, (8, p.nx))
            data[f"x_zdt{pid}_p{param}"] = xs
            data[f"f_zdt{pid}_p{param}"] = p.fitness_loop(xs)
    for pid in range(1, 8):
        for dim, fdim in ((5, 3), (12, 3), (7, 2), (30, 5)):
            p = R.problem("dtlz", pid, dim, fdim, 100)
            xs = rng.uniform(0, 1, (8, dim))
            data[f"x_dtlz{pid}_d{dim}_m{fdim}"] = xs
            data[f"f_dtlz{pid}_d{dim}_m{fdim}"] = p.fitness_loop(xs)
    np.savez_compressed(OUT / "mo_ref.npz", **data)
    print("wrote", [p.name for p in OUT.glob("*.npz")])


if __name__ == "__main__":
    main()
