"""GPU parity of gaco::evolve (pgc_gaco_evolve_device, gaco.cu) against the restated loop consuming the same Philox draws.  The
restatement itself is pinned bit for bit to the compiled reference on the mt19937 stream (tests/test_oracle_pin.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from pagmo2_b200 import capi as m
    return m


# (n, gens, keyword arguments)
CASES = [
    (40, 10, dict(ker=13)),
    (63, 6, dict()),  # reference defaults: ker 63 = the whole population
    (50, 14, dict(ker=20, oracle=1e9, acc=0.0, threshold=5, n_gen_mark=3)),
    (32, 9, dict(ker=5, q=0.5, acc=0.5, threshold=3, focus=4.0)),
    (48, 25, dict(ker=8, oracle=50.0, threshold=2, n_gen_mark=2, impstop=3)),
    (48, 30, dict(ker=8, evalstop=4)),
    (24, 8, dict(ker=2, q=2.0, oracle=-5.0, threshold=4, n_gen_mark=1, focus=100.0)),
]


@pytest.mark.parametrize("family,dim", [("rastrigin", 8), ("rosenbrock", 5), ("ackley", 12), ("schwefel", 4)])
def test_gaco_matches_oracle(capi, ctx, orc, family, dim):
    rng = np.random.default_rng(dim)
    prob = capi.Problem(ctx, family, dim=dim)
    op = orc.problem(family, dim=dim)
    lb, ub = prob.bounds()
    for n, gens, kw in CASES:
        x = rng.uniform(lb, ub, (n, dim))
        f = prob.eval_host(x)[:, 0]
        args = dict(gens=gens, seed=n + gens, first_generation=3, **kw)
        xo, fo, so, done_o = orc.gaco_evolve(op, lb, ub, x, f, **args)
        xg, fg, sg, done_g = prob.gaco_evolve(x, f, **args)
        assert done_g == done_o, (n, kw)
        assert np.allclose(xg, xo, rtol=1e-9, atol=1e-12), (n, kw, np.abs(xg - xo).max())
        assert np.allclose(fg[:, 0], fo, rtol=1e-9, atol=1e-12)
        assert (sg.n_evalstop, sg.n_impstop, sg.gen_mark, sg.fevals) == (so.n_evalstop, so.n_impstop, so.gen_mark, so.fevals), (n, kw)
        assert np.isclose(sg.oracle, so.oracle, rtol=1e-12) and sg.q == so.q
        assert (xg >= lb).all() and (xg <= ub).all() and np.allclose(prob.eval_host(xg)[:, 0], fg[:, 0], rtol=1e-12, atol=1e-15)
    prob.close()


def test_gaco_state_continues_and_stops(capi, ctx, orc):
    """the algorithm object's members travel in the state: a second evolve() starts from the oracle parameter and the counters the
    first one left (gaco.cpp:60-64: they are members, not locals), on the device as in the restated loop; a stopping criterion returns
    the population untouched by the archive write-back (:193-205)."""
    prob = capi.Problem(ctx, "rastrigin", dim=6)
    op = orc.problem("rastrigin", dim=6)
    lb, ub = prob.bounds()
    x = np.random.default_rng(0).uniform(lb, ub, (30, 6))
    f = prob.eval_host(x)[:, 0]
    kw = dict(ker=10, seed=4, oracle=1e6)
    xa, fa, st, _ = prob.gaco_evolve(x, f, gens=5, first_generation=1, **kw)
    assert st.oracle < 1e6 and st.fevals == 150
    xb, fb, st, _ = prob.gaco_evolve(xa, fa, gens=5, first_generation=6, state=st, **kw)
    xo, fo, so, _ = orc.gaco_evolve(op, lb, ub, x, f, gens=5, first_generation=1, **kw)
    xo, fo, so, _ = orc.gaco_evolve(op, lb, ub, xo, fo, gens=5, first_generation=6, state=so, **kw)
    assert np.allclose(xb, xo, rtol=1e-9, atol=1e-12) and st.fevals == so.fevals == 300 and st.gen_mark == so.gen_mark
    # evalstop = 1: the counter starts at 1, so the very first generation returns the population as it came
    xs, fs, _, done = prob.gaco_evolve(x, f, gens=5, ker=10, evalstop=1, seed=1)
    assert done == 0 and np.array_equal(xs, x) and np.array_equal(fs[:, 0], f)
    prob.close()


def test_gaco_memory_follows_the_restated_loop(capi, ctx, orc):
    """memory = true (gaco.cpp:106-108, :223-250, :732-752, :778-784): one algorithm object evolved call after call - the archive and the
    call counter carry over, the kernel weights switch when the COUNTER reaches the threshold, the archive is never written back.  The
    restated loop is pinned to the compiled reference in this mode too (tests/test_oracle_pin.py)."""
    prob = capi.Problem(ctx, "rastrigin", dim=7)
    op = orc.problem("rastrigin", dim=7)
    lb, ub = prob.bounds()
    rng = np.random.default_rng(4)
    for n, gens, calls, kw in ((20, 1, 6, dict(ker=8, oracle=1e9, threshold=3)), (24, 3, 4, dict(ker=6, threshold=2, n_gen_mark=3, focus=5.0))):
        x = rng.uniform(lb, ub, (n, 7))
        f = prob.eval_host(x)[:, 0]
        xg, fg, sg = x, f, None
        xo, fo, so = x, f, None
        for c in range(calls):
            xg, fg, sg, _ = prob.gaco_evolve(xg, fg, gens=gens, seed=9, first_generation=1 + c * gens, state=sg, memory=True, **kw)
            xo, fo, so, _ = orc.gaco_evolve(op, lb, ub, xo, fo, gens=gens, seed=9, first_generation=1 + c * gens, state=so, memory=True, **kw)
            fg = fg[:, 0]
            assert np.allclose(xg, xo, rtol=1e-9, atol=1e-12) and np.allclose(fg, fo, rtol=1e-9, atol=1e-12), (n, c)
            assert (sg.counter, sg.n_evalstop, sg.n_impstop, sg.gen_mark, sg.q) == (so.counter, so.n_evalstop, so.n_impstop, so.gen_mark, so.q)
            assert np.isclose(sg.champion_f, so.champion) and np.allclose(sg._archive, so._archive, rtol=1e-9, atol=1e-12)
        # the population that comes back holds the last ants, not the archive: the champion the state tracks is the better of everything seen
        assert sg.champion_f <= min(fg.min(), sg._archive[7 + 1]) + 1e-12
    prob.close()


def test_gaco_integer_tail_and_argument_checks(capi, ctx):
    """zdt5-like integer variables are rounded (:869-873) - exercised through a single-objective decomposition of zdt5; the
    reference's constructor / evolve checks (gaco.cpp:62-94, :157-171)."""
    z5 = capi.Problem(ctx, "zdt", prob_id=5, dim=11)
    dec = z5.decompose([0.5, 0.5], [0.0, 0.0], "weighted")
    lb, ub = dec.bounds()
    x = np.round(np.random.default_rng(2).uniform(lb, ub, (30, dec.nx)))
    f = dec.eval_host(x)
    xg, fg, _, _ = dec.gaco_evolve(x, f, gens=6, ker=8, seed=3)
    assert np.array_equal(xg, np.round(xg)) and (xg >= lb).all() and (xg <= ub).all()
    assert np.allclose(dec.eval_host(xg), fg, rtol=1e-12, atol=1e-15)
    prob = capi.Problem(ctx, "rastrigin", dim=4)
    x = np.random.default_rng(1).uniform(-5, 5, (16, 4))
    f = prob.eval_host(x)
    for bad in (dict(acc=-1.0), dict(focus=-1.0), dict(threshold=0), dict(threshold=9), dict(q=-1.0), dict(ker=1), dict(ker=17)):
        with pytest.raises(capi.PgcError):
            prob.gaco_evolve(x, f, gens=3, **{"ker": 8, **bad})
    with pytest.raises(capi.PgcError):
        z5.gaco_evolve(np.zeros((8, z5.nx)), np.zeros((8, 2)), gens=1, ker=4)
    # gen = 0 and the empty population return at once (:147-156)
    x0, f0, _, done = prob.gaco_evolve(x, f, gens=0, ker=8)
    assert done == 0 and np.array_equal(x0, x)
    dec.close(); z5.close(); prob.close()


# ---- maco (pgc_maco_evolve_device): the same ants on an archive ordered by hypervolume contributions -------------------------------
MACO_CASES = [
    (24, 8, dict(ker=10)),
    (30, 7, dict(ker=30, threshold=3, n_gen_mark=3)),
    (20, 9, dict(ker=4, q=0.5, threshold=2, focus=5.0)),
    (28, 15, dict(ker=12, n_gen_mark=2, evalstop=3)),
    (63, 5, dict()),
]


@pytest.mark.parametrize("family,kw", [("zdt", dict(prob_id=1, dim=8)), ("zdt", dict(prob_id=2, dim=6)), ("zdt", dict(prob_id=3, dim=7)),
                                       ("dtlz", dict(prob_id=2, dim=7, nobj=3, param=100)), ("dtlz", dict(prob_id=1, dim=6, nobj=3, param=100)),
                                       ("dtlz", dict(prob_id=2, dim=8, nobj=4, param=100))])
def test_maco_matches_oracle(capi, ctx, orc, family, kw):
    """GPU parity of maco::evolve against the restated loop on the same Philox draws (the restatement reproduces the compiled
    reference bit for bit on the mt19937 stream, tests/test_oracle_pin.py).  4 objectives take the device WFG for the contributions."""
    rng = np.random.default_rng(kw["dim"])
    prob = capi.Problem(ctx, family, **kw)
    op = orc.problem(family, **kw)
    lb, ub = prob.bounds()
    for n, gens, args in MACO_CASES:
        x = rng.uniform(lb, ub, (n, prob.nx))
        f = prob.eval_host(x)
        a = dict(gens=gens, seed=n + gens, first_generation=2, **args)
        xo, fo, so, done_o = orc.maco_evolve(op, lb, ub, x, f, **a)
        xg, fg, sg, done_g = prob.maco_evolve(x, f, **a)
        assert done_g == done_o and (sg.n_evalstop, sg.gen_mark, sg.q) == (so.n_evalstop, so.gen_mark, so.q), (n, args)
        assert np.allclose(xg, xo, rtol=1e-9, atol=1e-12), (n, args, np.abs(xg - xo).max())
        assert np.allclose(fg, fo, rtol=1e-9, atol=1e-12)
        assert (xg >= lb).all() and (xg <= ub).all() and np.allclose(prob.eval_host(xg), fg, rtol=1e-12, atol=1e-15)
    prob.close()


def test_maco_scale_and_argument_checks(capi, ctx):
    """a colony of 4096 ants on ZDT1: the archive the run leaves in the first `ker` rows is mutually non-dominated and closer to the
    front than the start; the reference's checks (maco.cpp:64-83, :126-152)."""
    prob = capi.Problem(ctx, "zdt", prob_id=1, dim=30)
    lb, ub = prob.bounds()
    n, ker = 4096, 63
    x = np.random.default_rng(0).uniform(lb, ub, (n, prob.nx))
    f = prob.eval_host(x)
    xg, fg, st, done = prob.maco_evolve(x, f, gens=30, ker=ker, seed=7)
    assert done == 30 and np.allclose(prob.eval_host(xg), fg, rtol=1e-12, atol=1e-15)
    a = fg[:ker]
    dominated = ((a[:, None, :] >= a[None, :, :]).all(axis=2) & (a[:, None, :] > a[None, :, :]).any(axis=2)).any(axis=1)
    assert not dominated.any()
    g = lambda ff: (ff[:, 1] + np.sqrt(np.maximum(ff[:, 0], 0)) - 1).mean()  # distance-like measure to ZDT1's front f2 = 1 - sqrt(f1)
    assert g(a) < 0.5 * g(f)
    small = x[:16], f[:16]
    for bad in (dict(focus=-1.0), dict(threshold=0), dict(threshold=9), dict(ker=17), dict(ker=1)):
        with pytest.raises(capi.PgcError):
            prob.maco_evolve(*small, gens=3, **{"ker": 8, **bad})
    so = capi.Problem(ctx, "rastrigin", dim=4)
    with pytest.raises(capi.PgcError):
        so.maco_evolve(np.zeros((8, 4)), np.zeros((8, 1)), gens=1, ker=4)
    x0, f0, _, done = prob.maco_evolve(*small, gens=0, ker=8)
    assert done == 0 and np.array_equal(x0, small[0])
    so.close(); prob.close()
