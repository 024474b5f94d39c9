"""GPU parity of nspso::evolve (pgc_nspso_evolve_device, nspso.cu) against the restated loop consuming the same Philox draws.  The
restatement itself is pinned bit for bit to the compiled reference on the mt19937 stream (tests/test_oracle_pin.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from pagmo2_b200 import capi as m
    return m


CASES = [("zdt", dict(prob_id=1, dim=8), 2, 40), ("zdt", dict(prob_id=3, dim=6), 2, 37), ("dtlz", dict(prob_id=2, dim=7, nobj=3, param=100), 3, 48),
         ("dtlz", dict(prob_id=1, dim=6, nobj=3, param=100), 3, 21)]


@pytest.mark.parametrize("family,kw,m,n", CASES)
@pytest.mark.parametrize("diversity", ("crowding distance", "niche count", "max min"))
def test_nspso_matches_oracle(capi, ctx, orc, family, kw, m, n, diversity):
    rng = np.random.default_rng(n + m)
    prob = capi.Problem(ctx, family, **kw)
    op = orc.problem(family, **kw)
    lb, ub = prob.bounds()
    x = rng.uniform(lb, ub, (n, prob.nx))
    x[5] = x[2]  # duplicates: ties in every order
    f = prob.eval_host(x)
    for lsr, gens in ((60, 7), (5, 3), (100, 3)):
        args = dict(gens=gens, leader_selection_range=lsr, diversity=diversity, seed=11 + lsr, first_generation=3)
        xo, fo, *_ = orc.nspso_evolve(op, lb, ub, x, f, **args)
        xg, fg, *_ = prob.nspso_evolve(x, f, **args)
        assert np.allclose(xg, xo, rtol=1e-9, atol=1e-12), (diversity, lsr, np.abs(xg - xo).max())
        assert np.allclose(fg, fo, rtol=1e-9, atol=1e-12)
        assert (xg >= lb).all() and (xg <= ub).all() and np.allclose(prob.eval_host(xg), fg, rtol=1e-12, atol=1e-15)
    prob.close()


def test_nspso_memory_continues_a_run(capi, ctx, orc):
    """velocities and the archive handed back and in again (memory = true, nspso.cpp:127-152): two calls of 4 generations equal one
    call of 8, on the device and in the restated loop, and the archive keeps only non-dominated-or-best rows of (moved | archive)."""
    prob = capi.Problem(ctx, "zdt", prob_id=2, dim=10)
    op = orc.problem("zdt", prob_id=2, dim=10)
    lb, ub = prob.bounds()
    n = 64
    x = np.random.default_rng(8).uniform(lb, ub, (n, prob.nx))
    f = prob.eval_host(x)
    kw = dict(seed=5, diversity="crowding distance")
    x8, f8, v8, bx8, bf8 = prob.nspso_evolve(x, f, gens=8, first_generation=1, vel=np.zeros_like(x), best_x=x, best_f=f, **kw)
    xa, fa, va, bxa, bfa = prob.nspso_evolve(x, f, gens=4, first_generation=1, vel=np.zeros_like(x), best_x=x, best_f=f, **kw)
    xb, fb, vb, bxb, bfb = prob.nspso_evolve(xa, fa, gens=4, first_generation=5, vel=va, best_x=bxa, best_f=bfa, **kw)
    assert np.array_equal(xb, x8) and np.array_equal(fb, f8) and np.array_equal(vb, v8) and np.array_equal(bxb, bx8) and np.array_equal(bfb, bf8)
    xo, fo, vo, bxo, bfo = orc.nspso_evolve(op, lb, ub, x, f, gens=8, first_generation=1, vel=np.zeros_like(x), best_x=x, best_f=f, **kw)
    assert np.allclose(x8, xo, rtol=1e-9, atol=1e-12) and np.allclose(bx8, bxo, rtol=1e-9, atol=1e-12) and np.allclose(bf8, bfo, rtol=1e-9, atol=1e-12)
    assert np.allclose(prob.eval_host(bx8), bf8, rtol=1e-12, atol=1e-15)
    prob.close()


def test_nspso_argument_checks_and_scale(capi, ctx):
    """the reference's constructor / evolve checks (nspso.cpp:50-82, :99-119), and a swarm of 16 384 on ZDT1: the archive's
    hypervolume-free sanity measure - the share of archive rows in the first front - grows."""
    prob = capi.Problem(ctx, "zdt", prob_id=1, dim=30)
    lb, ub = prob.bounds()
    x = np.random.default_rng(1).uniform(lb, ub, (16384, 30))
    f = prob.eval_host(x)
    for bad in (dict(omega=1.5), dict(c1=0.0), dict(chi=-1.0), dict(v_coeff=0.0), dict(v_coeff=1.5), dict(leader_selection_range=101)):
        with pytest.raises(capi.PgcError):
            prob.nspso_evolve(x[:8], f[:8], gens=1, **bad)
    with pytest.raises(capi.PgcError):
        prob.nspso_evolve(x[:1], f[:1], gens=1)
    so = capi.Problem(ctx, "rastrigin", dim=5)
    with pytest.raises(capi.PgcError):
        so.nspso_evolve(np.zeros((8, 5)), np.zeros((8, 1)), gens=1)
    so.close()
    xg, fg, v, bx, bf = prob.nspso_evolve(x, f, gens=15, seed=3, vel=np.zeros_like(x), best_x=x, best_f=f)
    r0 = ctx.fnds(f)["rank"]
    r1 = ctx.fnds(bf)["rank"]
    assert (r1 == 0).mean() > (r0 == 0).mean() and bf[:, 1].mean() < f[:, 1].mean()
    prob.close()
