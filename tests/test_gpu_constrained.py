"""GPU parity of the constrained UDPs and the `unconstrain` meta-problem against the restated oracle, through
the C ABI.

Reference: src/problems/hock_schittkowski_71.cpp:48-67, src/problems/luksan_vlcek1.cpp:46-88, src/problems/unconstrain.cpp:66-97,
136-223,269-276, src/problem.cpp:620-644,709-721, include/pagmo/utils/constrained.hpp:49-80.  The oracle side is pinned bit for bit
against the unmodified reference in tests/test_oracle.py::test_constrained_udps_and_unconstrain_are_bit_exact_vs_reference.
Tolerances: hock_schittkowski_71 is products and sums only - bit-exact; luksan_vlcek1's constraints are sums of terms of mixed sign
(3 x^3, x exp(.), sin sin ...), so the bound is 1e-12 of the sum of the terms' magnitudes (1e-12 relative on the objective); the
penalties are the reference's operations in the reference's order on those rows."""
import numpy as np
import pytest

from pagmo2_b200 import capi

pytestmark = pytest.mark.gpu
RTOL = 1e-12
METHODS = ("death penalty", "kuri", "weighted", "ignore_c", "ignore_o")


def lv_scale(xs):
    """per-element magnitude of the summed terms of luksan_vlcek1 (objective column: the objective itself)."""
    x0, x1, x2 = xs[:, :-2], xs[:, 1:-1], xs[:, 2:]
    con = 3 * np.abs(x1) ** 3 + 2 * np.abs(x2) + 5 + 1 + 4 * np.abs(x1) + np.abs(x0) * np.exp(x0 - x1) + 3
    return con


def test_hock_schittkowski_71_is_bit_exact(ctx, orc):
    p = capi.Problem(ctx, "hock_schittkowski_71")
    assert (p.nx, p.nobj, p.nec, p.nic, p.nf) == (4, 1, 1, 1, 3) and p.name == "Hock Schittkowski 71"
    lb, ub = p.bounds()
    assert np.array_equal(lb, np.ones(4)) and np.array_equal(ub, np.full(4, 5.0))
    assert np.array_equal(p.c_tol(), np.zeros(2))
    rng = np.random.default_rng(5)
    for n in (1, 255, 4097):
        xs = rng.uniform(1, 5, (n, 4))
        assert np.array_equal(p.eval_host(xs), orc.hock_schittkowski_71(xs))
    assert p.eval_host(np.empty((0, 4))).shape == (0, 3)


@pytest.mark.parametrize("dim", [3, 4, 10, 33, 257])
def test_luksan_vlcek1_matches_oracle(ctx, orc, dim):
    p = capi.Problem(ctx, "luksan_vlcek1", dim=dim)
    assert (p.nx, p.nobj, p.nec, p.nic, p.nf) == (dim, 1, dim - 2, 0, dim - 1) and p.name == "luksan_vlcek1"
    lb, ub = p.bounds()
    assert np.array_equal(lb, np.full(dim, -5.0)) and np.array_equal(ub, np.full(dim, 5.0))
    rng = np.random.default_rng(dim)
    xs = rng.uniform(-5, 5, (1031, dim))
    got, want = p.eval_host(xs), orc.luksan_vlcek1(xs)
    assert np.all(np.abs(got[:, 0] - want[:, 0]) <= RTOL * np.abs(want[:, 0]))
    assert np.all(np.abs(got[:, 1:] - want[:, 1:]) <= RTOL * lv_scale(xs))


def test_luksan_vlcek1_rejects_small_dimensions(ctx):
    with pytest.raises(capi.PgcError, match="minimum 3 dimension"):
        capi.Problem(ctx, "luksan_vlcek1", dim=2)


@pytest.mark.parametrize("method", METHODS)
def test_unconstrain_matches_oracle(ctx, orc, method):
    rng = np.random.default_rng(11)
    # hock_schittkowski_71: exact inner rows, so the penalized rows are bit-exact; tolerances 0 (everything infeasible), moderate, large
    inner = capi.Problem(ctx, "hock_schittkowski_71")
    xs = rng.uniform(1, 5, (3000, 4))
    fin = orc.hock_schittkowski_71(xs)
    w = rng.uniform(0.1, 2.0, 2)
    for tol in (np.zeros(2), np.array([3.0, 1.5]), np.array([60.0, 700.0])):
        inner.set_c_tol(tol)
        assert np.array_equal(inner.c_tol(), tol)
        p = inner.unconstrain(method, w if method == "weighted" else ())
        assert (p.nobj, p.nec, p.nic, p.nf, p.nx) == (1, 0, 0, 1, 4) and p.name == "Hock Schittkowski 71 [unconstrained]"
        want = orc.unconstrain_rows(fin, 1, 1, 1, tol, method, w)
        assert np.array_equal(p.eval_host(xs), want), (method, tol)
    # the wrapper keeps the tolerances it was created with (unconstrain copies its inner problem)
    inner.set_c_tol(np.array([3.0, 1.5]))
    p = inner.unconstrain(method, w if method == "weighted" else ())
    inner.set_c_tol(np.zeros(2))
    assert np.array_equal(p.eval_host(xs), orc.unconstrain_rows(fin, 1, 1, 1, [3.0, 1.5], method, w))
    # luksan_vlcek1: 8 equalities; rows near the feasible region (x = 1 satisfies nothing exactly: use a tolerance)
    inner = capi.Problem(ctx, "luksan_vlcek1", dim=10)
    xs = rng.uniform(-1.5, 1.5, (2000, 10))
    tol = rng.uniform(1.0, 6.0, 8)
    inner.set_c_tol(tol)
    w = rng.uniform(0.1, 2.0, 8)
    p = inner.unconstrain(method, w if method == "weighted" else ())
    got = p.eval_host(xs)
    fin_dev = inner.eval_host(xs)
    # the penalty applied to the device's own inner rows is the oracle's, bit for bit
    assert np.array_equal(got, orc.unconstrain_rows(fin_dev, 1, 8, 0, tol, method, w))
    # and against the oracle's inner rows within the evaluator's tolerance, wherever the feasibility decisions agree
    fin = orc.luksan_vlcek1(xs)
    want = orc.unconstrain_rows(fin, 1, 8, 0, tol, method, w)
    margin = np.abs(np.abs(fin[:, 1:]) - tol).min(axis=1)
    safe = margin > 1e-9
    assert safe.sum() > 1900
    bound = RTOL * np.maximum(1.0, np.abs(want[safe, 0])) * (1 + (w.sum() if method == "weighted" else 0) + 1e3)
    assert np.all(np.abs(got[safe, 0] - want[safe, 0]) <= bound)


def test_unconstrain_constructor_errors(ctx):
    hs = capi.Problem(ctx, "hock_schittkowski_71")
    with pytest.raises(capi.PgcError, match="can only be applied to constrained problems, the instance of Rastrigin Function is not one"):
        capi.Problem(ctx, "rastrigin", dim=5).unconstrain()
    with pytest.raises(capi.PgcError, match="Length of weight vector is: 1 while the problem constraints are: 2"):
        hs.unconstrain("weighted", [1.0])
    with pytest.raises(capi.PgcError, match="is not supported"):
        hs.unconstrain("mispelled")
    with pytest.raises(capi.PgcError, match="needs to be empty"):
        hs.unconstrain("kuri", [1.0, 1.0])
    # problem::set_c_tol, problem.cpp:620-660
    with pytest.raises(capi.PgcError, match="The tolerance vector size should be: 2, while a size of: 3 was detected"):
        hs.set_c_tol([0.0, 0.0, 0.0])
    with pytest.raises(capi.PgcError, match="NaN value at the index 1"):
        hs.set_c_tol([0.0, np.nan])
    with pytest.raises(capi.PgcError, match="negative value at the index 0"):
        hs.set_c_tol([-1.0, 0.0])
    with pytest.raises(capi.PgcError, match="cannot be negative"):
        hs.set_c_tol(-1.0)
    hs.set_c_tol(0.25)
    assert np.array_equal(hs.c_tol(), [0.25, 0.25])
    # an unconstrained wrapper of an unconstrained wrapper is refused like any unconstrained problem
    with pytest.raises(capi.PgcError, match="can only be applied to constrained problems"):
        hs.unconstrain().unconstrain()


def test_feasibility_rows(ctx, orc):
    rng = np.random.default_rng(3)
    hs = capi.Problem(ctx, "hock_schittkowski_71")
    tol = np.array([4.0, 30.0])
    hs.set_c_tol(tol)
    xs = rng.uniform(1, 5, (5000, 4))
    f = orc.hock_schittkowski_71(xs)
    f[7, 1] = np.nan  # a NaN constraint is never satisfied
    d_f = ctx.to_device(f)
    d_o = ctx.malloc(f.shape[0])
    try:
        hs.feasibility_device(d_f, f.shape[0], d_o)
        got = ctx.from_device(d_o, (f.shape[0],), dtype=np.uint8)
    finally:
        ctx.free(d_f)
        ctx.free(d_o)
    want = (np.maximum(np.abs(f[:, 1]) - tol[0], 0) <= 0) & (np.maximum(f[:, 2] - tol[1], 0) <= 0)
    assert np.array_equal(got.astype(bool), want) and 0 < want.sum() < want.size and not got[7]
    # death penalty == "objective where feasible, DBL_MAX elsewhere"
    dp = orc.unconstrain_rows(np.nan_to_num(f, nan=1e300), 1, 1, 1, tol, "death penalty")[:, 0]
    assert np.array_equal(dp == np.finfo(float).max, ~want)


def test_algorithms_refuse_constraints_and_run_on_the_unconstrained_problem(ctx, orc):
    hs = capi.Problem(ctx, "hock_schittkowski_71")
    rng = np.random.default_rng(9)
    x = rng.uniform(1, 5, (64, 4))
    with pytest.raises(capi.PgcError, match="Non linear constraints detected in Hock Schittkowski 71 instance"):
        hs.de_evolve(x.copy(), np.zeros(64), gens=2, algo="de")
    # de on unconstrain{hs71, "weighted"}: the penalized objective improves and the final rows re-evaluate to the oracle's values
    hs.set_c_tol([1e-3, 1e-3])
    w = [10.0, 10.0]
    p = hs.unconstrain("weighted", w)
    f = p.eval_host(x).reshape(-1, 1)
    x2, f2 = p.de_evolve(x.copy(), f.copy(), gens=60, algo="de", variant=2, seed=3)[:2]
    assert f2.min() < f.min()
    assert np.array_equal(f2.reshape(-1), orc.unconstrain_rows(orc.hock_schittkowski_71(x2), 1, 1, 1, [1e-3, 1e-3], "weighted", w)[:, 0])
    lb, ub = p.bounds()
    assert (x2 >= lb).all() and (x2 <= ub).all()
    # translate of a constrained problem keeps its constraints (translate forwards nec / nic to the inner problem)
    t = np.array([0.5, -0.25, 0.125, 1.0])
    pt = hs.translate(t)
    assert (pt.nec, pt.nic, pt.nf) == (1, 1, 3)
    assert np.array_equal(pt.eval_host(x), orc.hock_schittkowski_71(orc.translate_rows(x, t)))


def test_constrained_golden_fixture_on_device(ctx):
    """the device path against outputs of the compiled reference itself (tests/golden/constrained_ref.npz)."""
    from pathlib import Path
    g = np.load(Path(__file__).resolve().parent / "golden" / "constrained_ref.npz")
    for key, fam, dim in (("hock_schittkowski_71_0", "hock_schittkowski_71", 0), ("luksan_vlcek1_3", "luksan_vlcek1", 3),
                          ("luksan_vlcek1_10", "luksan_vlcek1", 10), ("luksan_vlcek1_33", "luksan_vlcek1", 33)):
        p = capi.Problem(ctx, fam, dim=dim)
        xs, f = g[f"x_{key}"], g[f"f_{key}"]
        got = p.eval_host(xs)
        if fam == "hock_schittkowski_71":
            assert np.array_equal(got, f)
        else:
            assert np.all(np.abs(got[:, 0] - f[:, 0]) <= RTOL * np.abs(f[:, 0])) and np.all(np.abs(got[:, 1:] - f[:, 1:]) <= RTOL * lv_scale(xs))
        for ti in (0, 1):
            tol, w = g[f"tol{ti}_{key}"], g[f"w{ti}_{key}"]
            p.set_c_tol(tol)
            margin = np.abs(np.abs(f[:, 1:1 + p.nec]) - tol[:p.nec]).min(axis=1)  # rows whose feasibility decisions are not borderline
            safe = margin > 1e-9
            for method in METHODS:
                want = g[f"u{ti}_{method.replace(' ', '_')}_{key}"]
                out = p.unconstrain(method, w if method == "weighted" else ()).eval_host(xs)
                if fam == "hock_schittkowski_71":
                    assert np.array_equal(out, want), (key, ti, method)
                else:
                    bound = RTOL * np.maximum(1.0, np.abs(want[safe, 0])) * (1e3 + w.sum())
                    assert safe.sum() >= 20 and np.all(np.abs(out[safe, 0] - want[safe, 0]) <= bound), (key, ti, method)
