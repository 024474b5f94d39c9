"""GPU parity of the simple genetic algorithm (all crossover / mutation / selection strategies) against the restated loop consuming the
same Philox draws."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from pagmo2_b200 import capi as m
    return m


CASES = list(itertools.product(("exponential", "binomial", "single", "sbx"), ("gaussian", "uniform", "polynomial"), ("tournament", "truncated")))


@pytest.mark.parametrize("crossover,mutation,selection", CASES)
def test_sga_matches_oracle(capi, ctx, orc, crossover, mutation, selection):
    rng = np.random.default_rng(abs(hash((crossover, mutation, selection))) % 2**32)
    NP, dim = 40, 10
    prob = capi.Problem(ctx, "rastrigin", dim=dim)
    op = orc.problem("rastrigin", dim=dim)
    lb, ub = prob.bounds()
    x = rng.uniform(lb, ub, (NP, dim))
    f = orc.simple("rastrigin", x)
    pm = 20.0 if mutation == "polynomial" else 0.1
    kw = dict(gens=6, cr=0.8, eta_c=5.0, m=0.15, param_m=pm, param_s=4, crossover=crossover, mutation=mutation, selection=selection, seed=31)
    xo, fo = orc.sga_evolve(op, lb, ub, x, f, first_generation=1, **kw)
    d = capi.algo_desc("sga", gens=6, seed=31, cr=0.8, eta_c=5.0, m=0.15, param_m=pm, param_s=4, crossover=capi.SGA_CROSSOVER[crossover],
                       mutation=capi.SGA_MUTATION[mutation], selection=capi.SGA_SELECTION[selection])
    xg, fg, done = prob.evolve(d, x, f, first_generation=1)
    assert done == 6
    assert np.allclose(xg, xo, rtol=1e-9, atol=1e-12), np.abs(xg - xo).max()
    assert np.allclose(fg[:, 0], fo, rtol=1e-9)
    assert (np.diff(fg[:, 0]) >= 0).all() and fg[0, 0] <= f.min() and (xg >= lb).all() and (xg <= ub).all()
    prob.close()


def test_sga_defaults_and_errors(capi, ctx):
    rng = np.random.default_rng(4)
    prob = capi.Problem(ctx, "rastrigin", dim=10)
    lb, ub = prob.bounds()
    x = rng.uniform(lb, ub, (1024, 10))
    f = prob.eval_host(x)
    xg, fg, _ = prob.evolve(capi.algo_desc("sga", gens=100, seed=1), x, f)
    assert fg.min() < 0.5 * f.min() and np.allclose(prob.eval_host(xg), fg, rtol=1e-12)
    for bad in (dict(cr=1.5), dict(eta_c=0.5), dict(m=-0.1), dict(param_s=0), dict(param_m=0.5), dict(mutation=0, param_m=2.0), dict(crossover=4),
                dict(param_s=2000)):
        with pytest.raises(capi.PgcError):
            prob.evolve(capi.algo_desc("sga", gens=1, seed=1, **bad), x, f)
    with pytest.raises(capi.PgcError):  # sbx needs an even population (sga.cpp:211-215)
        prob.evolve(capi.algo_desc("sga", gens=1, seed=1, crossover=3), x[:33], f[:33])
    zp = capi.Problem(ctx, "zdt", prob_id=1, dim=30)
    with pytest.raises(capi.PgcError):  # multi-objective (sga.cpp:198-201)
        zp.evolve(capi.algo_desc("sga", gens=1, seed=1), np.zeros((8, 30)), np.zeros((8, 2)))
