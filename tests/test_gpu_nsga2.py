"""GPU parity of the NSGA-II generation operators (shuffles, tournament + SBX + polynomial mutation, the generation
loop) against the restated operators consuming the SAME Philox draws (parity on injected draws)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from pagmo2_b200 import capi as m
    return m


def test_philox_matches_oracle(capi, ctx, orc):
    out = C.c_double()
    for args in ((0, 1, 0, 0, 0), (123456789012345, 3, 7, 4242, 9), (2**63 + 5, 5, 1000, 65535, 31)):
        capi.check(capi.lib().pgc_philox_u01(*args, C.byref(out)))
        assert out.value == orc.philox_u01(*args)
    for n in (1, 5, 1000, 65536):
        assert np.array_equal(ctx.philox_permutation(n, 99, 1, 3), orc.philox_perm(n, 99, 1, 3))


@pytest.mark.parametrize("NP,nx,m", [(8, 5, 2), (64, 30, 2), (1024, 12, 3)])
def test_variation_parity(capi, ctx, orc, NP, nx, m):
    rng = np.random.default_rng(NP)
    lb, ub = np.zeros(nx), np.ones(nx)
    lb[0], ub[0] = -2.0, 3.0
    x = rng.uniform(lb, ub, (NP, nx))
    x[1] = x[0]  # identical parents: the |p1-p2| > 1e-14 guard (genetic_operators.cpp:94)
    f = rng.uniform(0, 1, (NP, m))
    f[5] = f[4]  # equal rank and crowding -> the tournament consumes a draw (:209-210)
    rank, cd = orc.nsga2_rank_crowding(f)
    sh1, sh2 = orc.philox_perm(NP, 11, 1, 2), orc.philox_perm(NP, 11, 2, 2)
    for cr, mm in ((0.95, 0.01), (0.5, 0.5), (0.0, 1.0)):
        want = orc.nsga2_variation(x, rank, cd, lb, ub, sh1, sh2, cr, 10.0, mm, 50.0, 11, 2)
        got = ctx.nsga2_variation(x, rank, cd, lb, ub, sh1, sh2, cr, 10.0, mm, 50.0, 11, 2)
        assert (got >= lb).all() and (got <= ub).all()
        # identical control flow (same draws, same comparisons): genes either copied bit-exactly or recomputed through pow()
        assert np.allclose(got, want, rtol=1e-12, atol=1e-14), np.abs(got - want).max()
        assert (got == want).mean() > 0.5 or mm == 1.0


def test_one_generation_matches_oracle(capi, ctx, orc):
    """One full generation (shuffle, FNDS, crowding, variation, evaluation, select_best_N_mo) on ZDT1 and DTLZ2: the
    survivors are the same individuals in the same order; values agree within 1e-12."""
    rng = np.random.default_rng(5)
    for fam, pid, nx, nobj in (("zdt", 1, 30, 2), ("dtlz", 2, 12, 3)):
        NP = 256
        prob = capi.Problem(ctx, fam, prob_id=pid, dim=nx, nobj=nobj, param=100)
        lb, ub = prob.bounds()
        x = rng.uniform(lb, ub, (NP, nx))
        f = orc.zdt(pid, x) if fam == "zdt" else orc.dtlz(pid, x, nobj, 100)
        xo, fo = orc.nsga2_evolve(fam, pid, nobj, 100, lb, ub, x, f, gens=1, cr=0.95, eta_c=10, m=0.05, eta_m=50, seed=17)
        # each side starts from fitness values of ITS OWN evaluator: an offspring that escaped crossover and mutation is a clone
        # of its parent, and only then does it tie with it exactly on both sides (device and libm cos/sin differ in the last bits)
        xg, fg = prob.nsga2_evolve(x, prob.eval_host(x), gens=1, cr=0.95, eta_c=10, m=0.05, eta_m=50, seed=17)
        assert np.allclose(xg, xo, rtol=1e-12, atol=1e-14) and np.allclose(fg, fo, rtol=1e-12, atol=1e-14)
        prob.close()


def test_evolve_converges_and_checks_arguments(capi, ctx, orc):
    rng = np.random.default_rng(1)
    NP, nx = 256, 30
    prob = capi.Problem(ctx, "zdt", prob_id=1, dim=nx)
    x = rng.uniform(0, 1, (NP, nx))
    f = prob.eval_host(x)
    x2, f2 = prob.nsga2_evolve(x, f, gens=100, seed=2)
    assert (x2 >= 0).all() and (x2 <= 1).all()
    assert np.allclose(prob.eval_host(x2), f2, rtol=1e-12)
    g0, g1 = 1 + 9 * x[:, 1:].sum(1) / (nx - 1), 1 + 9 * x2[:, 1:].sum(1) / (nx - 1)
    assert g1.mean() < 0.35 * g0.mean()  # moved towards the Pareto front (g = 1)
    assert len(ctx.fnds(f2)["fronts"]) < len(ctx.fnds(f)["fronts"])
    for NPbad in (4, 10):  # nsga2.cpp:121-126
        with pytest.raises(capi.PgcError):
            prob.nsga2_evolve(x[:NPbad], f[:NPbad], gens=1)
    with pytest.raises(capi.PgcError):
        prob.nsga2_evolve(x, f, gens=1, cr=1.0)
    so = capi.Problem(ctx, "rastrigin", dim=5)  # single objective: nsga2.cpp:117-120
    with pytest.raises(capi.PgcError):
        so.nsga2_evolve(np.zeros((8, 5)), np.zeros((8, 1)), gens=1)


def test_carried_ranking_equals_fresh_sorting(capi, ctx, orc):
    """From the second generation on, the ranks / fronts of the population are derived from the previous generation's
    select_best_N_mo instead of a new fast_non_dominated_sorting.  Five generations in one call (carried ranking) must give exactly
    the population of five one-generation calls (fresh sorting every time), and follow the restated reference loop."""
    rng = np.random.default_rng(12)
    for fam, pid, nx, nobj, NP in (("zdt", 1, 30, 2, 256), ("zdt", 3, 10, 2, 1000), ("dtlz", 2, 12, 3, 512), ("dtlz", 1, 7, 3, 64)):
        prob = capi.Problem(ctx, fam, prob_id=pid, dim=nx, nobj=nobj, param=100)
        lb, ub = prob.bounds()
        x = rng.uniform(lb, ub, (NP, nx))
        f = prob.eval_host(x)
        kw = dict(cr=0.9, eta_c=10, m=0.05, eta_m=50, seed=5)
        xa, fa = prob.nsga2_evolve(x, f, gens=5, first_generation=0, **kw)
        xb, fb = x, f
        for g in range(5):
            xb, fb = prob.nsga2_evolve(xb, fb, gens=1, first_generation=g, **kw)
        assert np.array_equal(xa, xb) and np.array_equal(fa, fb), (fam, pid)
        if fam == "zdt" and pid == 1:  # zdt1 is evaluated bit-exactly on the device: the restated loop must agree over 5 generations
            xo, fo = orc.nsga2_evolve(fam, pid, nobj, 100, lb, ub, x, orc.zdt(pid, x), gens=5, first_generation=0, **kw)
            assert np.allclose(xa, xo, rtol=1e-12, atol=1e-14) and np.allclose(fa, fo, rtol=1e-12, atol=1e-14)
        prob.close()


@pytest.mark.parametrize("fam,pid,nx,nobj", [("zdt", 1, 30, 2), ("dtlz", 2, 12, 3)])
def test_generation_at_cfg3_size_matches_restated_reference(capi, ctx, orc, fam, pid, nx, nobj):
    """BASELINE cfg3 at its own size: ONE whole nsga2 generation at pop 65 536 (ranking of N, tournament + SBX + mutation, batch
    evaluation, select_best_N_mo over 2N = 131 072) against the restated reference loop on the same Philox draws - the same
    survivors in the same order.  The restated loop is the one pinned bit for bit to nsga2.cpp (tests/test_oracle_pin.py);
    ~1 minute of host time per case for its two O(N^2) sorts."""
    rng = np.random.default_rng(65536 + pid)
    NP = 65536
    prob = capi.Problem(ctx, fam, prob_id=pid, dim=nx, nobj=nobj, param=100)
    lb, ub = prob.bounds()
    x = rng.uniform(lb, ub, (NP, nx))
    f = orc.zdt(pid, x) if fam == "zdt" else orc.dtlz(pid, x, nobj, 100)
    kw = dict(gens=1, cr=0.95, eta_c=10, m=1.0 / nx, eta_m=50, seed=23)
    xo, fo = orc.nsga2_evolve(fam, pid, nobj, 100, lb, ub, x, f, **kw)
    xg, fg = prob.nsga2_evolve(x, prob.eval_host(x), **kw)
    assert np.allclose(xg, xo, rtol=1e-12, atol=1e-14) and np.allclose(fg, fo, rtol=1e-12, atol=1e-14)
    # survivors that are untouched parents or unmutated clones must be bit-identical rows in the same positions
    assert (xg == xo).all(axis=1).mean() > 0.3
    prob.close()


@pytest.mark.parametrize("param,NP", [(3, 24), (11, 64), (5, 400)])
def test_nsga2_on_zdt5_integer_alleles(capi, ctx, orc, param, NP):
    """ZDT5 is all-integer: the device loop applies the reference's two-point crossover and uniform integer mutation to the integer
    alleles (genetic_operators.cpp:125-137, :187-195) - same trajectory as the restated loop (pinned to the compiled reference on
    the mt19937 stream in tests/test_oracle_pin.py), integer decision vectors throughout; the other evolve entry points still
    refuse integer genes."""
    prob = capi.Problem(ctx, "zdt", prob_id=5, dim=param)
    lb, ub = prob.bounds()
    nx = prob.nx
    x = np.floor(np.random.default_rng(param).uniform(lb, ub + 1, (NP, nx))).clip(lb, ub)
    f = prob.eval_host(x)
    orc.set_nix(nx)
    try:
        for cr, m, gens in ((0.95, 0.01, 6), (0.5, 0.2, 4), (0.9, 1.0 / nx, 5)):
            xo, fo = orc.nsga2_evolve("zdt", 5, 2, 0, lb, ub, x, f, gens, cr, 10.0, m, 50.0, 17, 2)
            xg, fg = prob.nsga2_evolve(x, f, gens, cr=cr, eta_c=10.0, m=m, eta_m=50.0, seed=17, first_generation=2)
            assert np.array_equal(xg, xo) and np.allclose(fg, fo, rtol=1e-12, atol=0)
            assert np.array_equal(xg, np.round(xg)) and (xg >= lb).all() and (xg <= ub).all()
    finally:
        orc.set_nix(0)
    with pytest.raises(capi.PgcError):
        prob.de_evolve(x, f[:, 0], gens=1, algo="de")
    prob.close()
