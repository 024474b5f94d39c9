"""GPU parity of the device meta-problems (translate, decompose) against the restated oracle, through the C ABI.

Reference: src/problems/translate.cpp:100-153,175-181; src/problems/decompose.cpp:66-154; decompose_objectives
src/utils/multi_objective.cpp:582-638.  The oracle side is pinned bit-exactly against the unmodified reference in
tests/test_oracle.py::test_meta_restatement_is_bit_exact_vs_reference.  Tolerance: 1e-12 relative (north_star) on the inner
fitness; the decomposition and the de-shifting themselves are the reference's operations in the reference's order."""
import numpy as np
import pytest

from pagmo2_b200 import capi

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def close(a, b, rtol=RTOL):
    return np.all(np.abs(a - b) <= rtol * np.maximum(1.0, np.abs(b)))


def test_translate_matches_oracle_and_moves_bounds(ctx, orc):
    rng = np.random.default_rng(5)
    for fam, dim in (("rastrigin", 10), ("ackley", 7), ("rosenbrock", 5)):
        inner = capi.Problem(ctx, fam, dim=dim)
        t = rng.uniform(-1, 1, dim)
        p = inner.translate(t)
        lb, ub = inner.bounds()
        plb, pub = p.bounds()
        assert np.array_equal(plb, lb + t) and np.array_equal(pub, ub + t)
        assert p.name.endswith("[translated]") and p.nx == dim and p.nobj == 1
        xs = rng.uniform(lb + t, ub + t, (1000, dim))
        want = orc.simple(fam, orc.translate_rows(xs, t))
        got = p.eval_host(xs).reshape(-1)
        assert close(got, want)
        # de-shifting twice by opposite vectors gives back the inner problem (tests/translate.cpp:79-83)
        back = p.translate(-t)
        assert close(back.eval_host(xs).reshape(-1), orc.simple(fam, orc.translate_rows(orc.translate_rows(xs, t), -t)))
    with pytest.raises(capi.PgcError, match="Length of shift vector is: 2 while the problem dimension is: 5"):
        inner.translate([1.0, 2.0])


def test_translate_cec2014(ctx, orc):
    rng = np.random.default_rng(6)
    dim = 30
    for func in (1, 9, 17, 24):
        mr, os_c, s = orc.cec2014_problem_tables(func, dim)
        inner = capi.Problem(ctx, "cec2014", prob_id=func, dim=dim, rotation=mr, shift=os_c, shuffle=s)
        t = rng.uniform(-3, 3, dim)
        p = inner.translate(t)
        xs = rng.uniform(-100, 100, (257, dim))
        want = orc.cec2014(func, orc.translate_rows(xs, t), tables=(mr, os_c, s))
        assert close(p.eval_host(xs).reshape(-1), want)


@pytest.mark.parametrize("method", ["weighted", "tchebycheff", "bi"])
def test_decompose_matches_oracle(ctx, orc, method):
    rng = np.random.default_rng(7)
    # zdt1 (2 objectives), dtlz2 (3 objectives), incl. a zero weight (tchebycheff's 1e-4 substitution)
    xs = rng.uniform(0, 1, (2049, 30))
    inner = capi.Problem(ctx, "zdt", prob_id=1, dim=30)
    for w in ([0.3, 0.7], [1.0, 0.0]):
        z = [0.1, -0.2]
        p = inner.decompose(w, z, method)
        assert p.nobj == 1 and p.nx == 30 and p.name.endswith("[decomposed]")
        assert close(p.eval_host(xs).reshape(-1), orc.decompose_rows(orc.zdt(1, xs), w, z, method))
    xs = rng.uniform(0, 1, (999, 12))
    inner3 = capi.Problem(ctx, "dtlz", prob_id=2, dim=12, nobj=3, param=100)
    w, z = rng.dirichlet(np.ones(3)), rng.normal(0, 0.1, 3)
    p = inner3.decompose(w, z, method)
    assert close(p.eval_host(xs).reshape(-1), orc.decompose_rows(orc.dtlz(2, xs, 3), w, z, method))


def test_decompose_known_answer_and_errors(ctx):
    # tests/decompose.cpp:116-141: zdt{1, 2} at (1, 1)
    inner = capi.Problem(ctx, "zdt", prob_id=1, dim=2)
    x = np.array([[1.0, 1.0]])
    f = inner.eval_host(x)[0]
    lam, z = np.array([0.5, 0.5]), np.zeros(2)
    assert inner.decompose(lam, z, "weighted").eval_host(x)[0] == pytest.approx(f @ lam, rel=1e-10)
    assert inner.decompose(lam, z, "tchebycheff").eval_host(x)[0] == pytest.approx(np.max(lam * np.abs(f - z)), rel=1e-10)
    il = lam / np.sqrt(lam @ lam)
    d1 = (f - z) @ il
    d2 = np.sqrt(np.sum((f - (z + d1 * il)) ** 2))
    assert inner.decompose(lam, z, "bi").eval_host(x)[0] == pytest.approx(d1 + 5.0 * d2, rel=1e-10)
    # constructor checks, decompose.cpp:68-124
    with pytest.raises(capi.PgcError, match="multi-objective"):
        capi.Problem(ctx, "rastrigin", dim=3).decompose([0.5, 0.5], [0.0, 0.0])
    with pytest.raises(capi.PgcError, match="Weight vector size"):
        inner.decompose([0.2, 0.3, 0.5], [0.0, 0.0, 0.0])
    with pytest.raises(capi.PgcError, match="must sum to 1"):
        inner.decompose([0.5, 0.6], [0.0, 0.0])
    with pytest.raises(capi.PgcError, match="non negative"):
        inner.decompose([1.5, -0.5], [0.0, 0.0])
    with pytest.raises(capi.PgcError, match="non finite"):
        inner.decompose([0.5, 0.5], [0.0, np.inf])
    with pytest.raises(capi.PgcError, match="Decomposition method requested is: pippo"):
        inner.decompose([0.5, 0.5], [0.0, 0.0], "pippo")
    with pytest.raises(capi.PgcError, match="ideal-point adaptation"):
        inner.decompose([0.5, 0.5], [0.0, 0.0], "weighted", adapt_ideal=True)


def test_meta_problem_drives_an_algorithm(ctx, orc):
    # a decomposed problem is single-objective: the device DE runs on it unchanged and improves the decomposed fitness
    inner = capi.Problem(ctx, "zdt", prob_id=1, dim=30)
    p = inner.decompose([0.5, 0.5], [0.0, 0.0], "tchebycheff")
    rng = np.random.default_rng(9)
    x = rng.uniform(0, 1, (64, 30))
    f = p.eval_host(x).reshape(-1, 1)
    x2, f2 = p.de_evolve(x.copy(), f.copy(), gens=30, algo="de", variant=2, seed=3)[:2]
    assert f2.min() < f.min()
    assert close(f2.reshape(-1), orc.decompose_rows(orc.zdt(1, x2), [0.5, 0.5], [0.0, 0.0], "tchebycheff"))
