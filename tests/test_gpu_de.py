"""GPU parity of the differential-evolution family (de, sade, de1220; generational form) against the restated loop consuming
the same Philox draws."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from pagmo2_b200 import capi as m
    return m


CASES = [("de", dict(variant=v)) for v in range(1, 11)] + [("sade", dict(variant=v, variant_adptv=a)) for v in (1, 2, 7, 11, 12, 13, 16, 17, 18)
                                                           for a in (1, 2)] + [("de1220", dict(variant_adptv=a)) for a in (1, 2)]


@pytest.mark.parametrize("algo,kw", CASES)
def test_de_family_matches_oracle(capi, ctx, orc, algo, kw):
    rng = np.random.default_rng(hash((algo, tuple(kw.items()))) % 2**32)
    NP, dim = 40, 10
    prob = capi.Problem(ctx, "rastrigin", dim=dim)
    op = orc.problem("rastrigin", dim=dim)
    lb, ub = prob.bounds()
    x = rng.uniform(lb, ub, (NP, dim))
    f = orc.simple("rastrigin", x)
    args = dict(gens=6, algo=algo, seed=77, first_generation=1, ftol=0.0, xtol=0.0, **kw)
    xo, fo, go, *_ = orc.de_evolve(op, lb, ub, x, f, **args)
    xg, fg, gg = prob.de_evolve(x, f, **args)
    assert gg == go == 6
    assert np.allclose(xg, xo, rtol=1e-9, atol=1e-12), np.abs(xg - xo).max()
    assert np.allclose(fg, fo, rtol=1e-9)
    assert (fg <= f).all() and (xg >= lb).all() and (xg <= ub).all()
    prob.close()


def test_cfg1_de1220_rastrigin(capi, ctx, orc):
    """BASELINE cfg1: rastrigin D=10, pop 1024, de1220, 100 generations - the device loop improves like the restated one and
    stops on the reference's ftol/xtol exit conditions when asked."""
    rng = np.random.default_rng(23)
    NP, dim = 1024, 10
    prob = capi.Problem(ctx, "rastrigin", dim=dim)
    lb, ub = prob.bounds()
    x = rng.uniform(lb, ub, (NP, dim))
    f = prob.eval_host(x)[:, 0]
    xg, fg, gg = prob.de_evolve(x, f, gens=100, algo="de1220", seed=41)
    assert gg == 100 and fg.min() < 0.25 * f.min() and np.allclose(prob.eval_host(xg)[:, 0], fg, rtol=1e-12)
    xs, fs, gs = prob.de_evolve(np.tile(x[:1], (16, 1)), np.tile(f[:1], 16), gens=50, algo="de", seed=1)  # flat population: exit at once
    assert gs == 1
    for bad in (dict(algo="de", variant=11), dict(algo="sade", variant=19), dict(algo="sade", variant_adptv=3), dict(algo="de1220", allowed=(0,)),
                dict(algo="de", F=1.5)):
        with pytest.raises(capi.PgcError):
            prob.de_evolve(x, f, gens=1, **bad)
    with pytest.raises(capi.PgcError):
        prob.de_evolve(x[:4], f[:4], gens=1, algo="de")


def test_large_population_path_matches_oracle(capi, ctx, orc):
    """populations of 16384 and more spread the global-best scan over many CTAs: same trajectory as the restated loop."""
    rng = np.random.default_rng(99)
    NP, dim = 16384, 6
    prob = capi.Problem(ctx, "rastrigin", dim=dim)
    op = orc.problem("rastrigin", dim=dim)
    lb, ub = prob.bounds()
    x = rng.uniform(lb, ub, (NP, dim))
    x[5000] = x[77]  # ties in the fitness: best / worst / accepted tie rules across CTA slices
    x[16000] = x[77]
    f = orc.simple("rastrigin", x)
    for algo, kw in (("de", dict(variant=1)), ("sade", dict(variant=3, variant_adptv=2)), ("de1220", dict(variant_adptv=1))):
        args = dict(gens=3, algo=algo, seed=5, first_generation=1, ftol=0.0, xtol=0.0, **kw)
        xo, fo, go, *_ = orc.de_evolve(op, lb, ub, x, f, **args)
        xg, fg, gg = prob.de_evolve(x, f, **args)
        assert gg == go == 3
        assert np.allclose(xg, xo, rtol=1e-9, atol=1e-12) and np.allclose(fg, fo, rtol=1e-9), algo
    prob.close()
