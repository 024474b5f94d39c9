"""GPU parity of moead_gen::evolve (pgc_moead_gen_evolve_device, moead.cu) against the restated loop consuming the same Philox draws.
The restatement is pinned bit for bit to the compiled reference on the mt19937 stream (tests/test_oracle_pin.py).  Weight vectors
and neighbourhoods come from the compiled reference's own decomposition_weights / kNN where it is present (the authoring
container) and from small committed stand-ins otherwise (the GPU box has no reference)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from pagmo2_b200 import capi as m
    return m


def weights_and_neighbours(NP, m, T, seed):
    """simplex weights (corners first, as pagmo does) and their T nearest neighbours; plain numpy so that the GPU box needs no reference"""
    rng = np.random.default_rng(seed)
    w = np.vstack([np.eye(m), rng.dirichlet(np.ones(m), NP - m)])
    d = np.sqrt(((w[:, None, :] - w[None, :, :]) ** 2).sum(-1))
    np.fill_diagonal(d, np.inf)
    return w, np.argsort(d, axis=1, kind="stable")[:, :T].astype(np.uint32)


CASES = [("zdt", dict(prob_id=1, dim=8), 2, 40), ("zdt", dict(prob_id=4, dim=6), 2, 33), ("dtlz", dict(prob_id=2, dim=7, nobj=3, param=100), 3, 45)]


@pytest.mark.parametrize("family,kw,m,NP", CASES)
@pytest.mark.parametrize("decomposition", ("tchebycheff", "weighted", "bi"))
def test_moead_gen_matches_oracle(capi, ctx, orc, family, kw, m, NP, decomposition):
    prob = capi.Problem(ctx, family, **kw)
    op = orc.problem(family, **kw)
    lb, ub = prob.bounds()
    x = np.random.default_rng(NP + m).uniform(lb, ub, (NP, prob.nx))
    f = prob.eval_host(x)
    for T, CR, F, realb, limit, preserve, gens in ((6, 1.0, 0.5, 0.9, 2, True, 6), (9, 0.6, 0.8, 0.4, 1, True, 4), (5, 0.9, 0.5, 0.9, 2, False, 4),
                                                   (4, 1.0, 0.5, 0.5, 1000, True, 3), (4, 1.0, 0.5, 0.5, 0, True, 3)):
        w, nb = weights_and_neighbours(NP, m, T, NP + T)
        args = dict(gens=gens, decomposition=decomposition, CR=CR, F=F, eta_m=20.0, realb=realb, limit=limit, preserve_diversity=preserve,
                    seed=5 + T, first_generation=2)
        xo, fo = orc.moead_gen_evolve(op, lb, ub, x, f, w, nb, **args)
        xg, fg = prob.moead_gen_evolve(x, f, w, nb, **args)
        assert np.allclose(xg, xo, rtol=1e-9, atol=1e-12), (decomposition, T, limit, preserve, np.abs(xg - xo).max())
        assert np.allclose(fg, fo, rtol=1e-9, atol=1e-12)
        assert (xg >= lb).all() and (xg <= ub).all() and np.allclose(prob.eval_host(xg), fg, rtol=1e-12, atol=1e-15)
    prob.close()


def test_moead_gen_checks_and_progress(capi, ctx):
    """the reference's checks (moead_gen.cpp:60-104, :140-166), and on ZDT1 with 500 sub-problems the average decomposed fitness
    (the quantity the reference logs, :196-199) falls"""
    prob = capi.Problem(ctx, "zdt", prob_id=1, dim=30)
    lb, ub = prob.bounds()
    NP, T = 500, 20
    w, nb = weights_and_neighbours(NP, 2, T, 1)
    x = np.random.default_rng(2).uniform(lb, ub, (NP, 30))
    f = prob.eval_host(x)
    for bad in (dict(CR=1.5), dict(F=-0.1), dict(eta_m=-1.0), dict(realb=2.0)):
        with pytest.raises(capi.PgcError):
            prob.moead_gen_evolve(x, f, w, nb, gens=1, **bad)
    with pytest.raises(capi.PgcError):
        prob.moead_gen_evolve(x, f, w, nb[:, :1], gens=1)          # T < 2
    with pytest.raises(capi.PgcError):
        prob.moead_gen_evolve(x[:10], f[:10], w[:10], nb[:10] % 10, gens=1)   # T > NP - 1
    bad_nb = nb.copy()
    bad_nb[3, 2] = NP
    with pytest.raises(capi.PgcError):
        prob.moead_gen_evolve(x, f, w, bad_nb, gens=1)
    xg, fg = prob.moead_gen_evolve(x, f, w, nb, gens=40, seed=3)

    def adf(ff):
        ideal = np.minimum(f.min(0), fg.min(0))
        return np.max(np.where(w == 0, 1e-4, w) * np.abs(ff - ideal), axis=1).mean()
    assert adf(fg) < 0.5 * adf(f)
    prob.close()
