"""memory = true of the reference UDAs (sade.cpp:137-156, de1220.cpp:147-165, pso_gen.cpp:193-201, nspso.cpp:127-152) behind
`pgc_algo_evolve_memory_device` and inside a resident island (`pgc_island_evolve` with `algo.memory`): the adaptation state, the
velocities and nspso's archive survive between evolve() calls.  The draws are keyed by (seed, generation, individual), so a run
split into two calls must equal the uninterrupted run bit for bit when the state is kept, and must differ when it is re-drawn."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from pagmo2_b200 import capi as m
    return m


SO_CASES = [
    ("sade", dict(variant=2, variant_adptv=1)),
    ("sade", dict(variant=7, variant_adptv=2)),
    ("de1220", dict(variant_adptv=1, allowed_variants=[2, 3, 7, 10, 13])),
    ("de1220", dict(variant_adptv=2)),
    ("pso_gen", dict(variant=5, neighb_type=2)),
    ("pso_gen", dict(variant=1, neighb_type=1)),
    ("pso_gen", dict(variant=6, neighb_type=3)),
]


def _so_problem(capi, ctx, n=48):
    prob = capi.Problem(ctx, "rosenbrock", dim=9)
    lb, ub = prob.bounds()
    x = np.random.default_rng(3).uniform(lb, ub, (n, prob.nx))
    return prob, x, prob.eval_host(x)


@pytest.mark.parametrize("name,kw", SO_CASES)
def test_split_run_with_memory(capi, ctx, name, kw):
    prob, x, f = _so_problem(capi, ctx)
    tol = {} if name == "pso_gen" else dict(ftol=0.0, xtol=0.0)
    whole = capi.algo_desc(name, gens=6, seed=17, **tol, **kw)
    half = capi.algo_desc(name, gens=3, seed=17, **tol, **kw)
    x6, f6, done = prob.evolve(whole, x, f, first_generation=1)
    assert done == 6
    xa, fa, _, st = prob.evolve_memory(half, x, f, first_generation=1)
    x3, f3, _ = prob.evolve(half, x, f, first_generation=1)
    assert np.array_equal(xa, x3) and np.array_equal(fa, f3)  # a first call with memory draws what a memory-less call draws
    xb, fb, _, st2 = prob.evolve_memory(half, xa, fa, first_generation=4, state=st)
    # memory-less: the second call re-draws F / CR / the velocities from the streams of generation 4
    xc, fc, _ = prob.evolve(half, xa, fa, first_generation=4)
    assert not np.array_equal(xc, xb)
    assert np.allclose(prob.eval_host(xb)[:, 0], fb.ravel(), rtol=1e-12, atol=1e-15) and fb.min() <= fa.min()
    if name == "pso_gen":
        # the reference restarts from the particles' best positions with the kept velocities (pso_gen.cpp:187-201, :524-527)
        lb, ub = prob.bounds()
        xe, fe, ve, _ = prob.pso_evolve(xa, fa.ravel(), v=st["a"], gens=3, seed=17, first_generation=4, variant=kw["variant"],
                                        neighb_type=kw["neighb_type"])
        assert np.array_equal(xb, xe) and np.array_equal(fb.ravel(), fe) and np.array_equal(st2["a"], ve)
        assert (np.abs(st2["a"]) <= 0.5 * (ub - lb) + 1e-12).all() and np.abs(st2["a"]).max() > 0
    elif kw.get("variant_adptv") == 1:
        # jDE (variant_adptv 1) reads nothing but F[i] / CR[i] (/ variant[i]): with the state kept, 3 + 3 generations are the
        # uninterrupted 6.  (variant_adptv 2 also reads the F / CR of the iteration's best, which every evolve() resets to
        # F[0] / CR[0] - sade.cpp:158-162 - so there the split run differs in the reference as well.)
        assert np.array_equal(xb, x6) and np.array_equal(fb, f6)
        n = x.shape[0]
        F, CR = st2["a"].ravel()[:n], st2["b"].ravel()[:n]
        assert (F >= 0.1).all() and (F <= 1.0).all() and (CR >= 0).all() and (CR <= 1).all()
    else:
        assert not np.array_equal(st2["a"].ravel()[:x.shape[0]], st["a"].ravel()[:x.shape[0]])
    if name == "de1220":
        allowed = kw.get("allowed_variants", [2, 3, 7, 10, 13, 14, 15, 16])
        assert set(st2["u"].tolist()) <= set(allowed)
    prob.close()


def test_nspso_memory_through_the_descriptor(capi, ctx):
    prob = capi.Problem(ctx, "zdt", prob_id=1, dim=12)
    lb, ub = prob.bounds()
    x = np.random.default_rng(1).uniform(lb, ub, (40, prob.nx))
    f = prob.eval_host(x)
    whole, half = capi.algo_desc("nspso", gens=6, seed=9), capi.algo_desc("nspso", gens=3, seed=9)
    x6, f6, _ = prob.evolve(whole, x, f, first_generation=1)
    xa, fa, _, st = prob.evolve_memory(half, x, f, first_generation=1)
    xb, fb, _, st = prob.evolve_memory(half, xa, fa, first_generation=4, state=st)
    assert np.array_equal(xb, x6) and np.array_equal(fb, f6)
    assert np.allclose(prob.eval_host(st["b"]), st["c"], rtol=1e-12, atol=1e-15)  # the archive is consistent
    prob.close()


@pytest.mark.parametrize("name,kw", [("sade", dict(variant_adptv=1, ftol=0.0, xtol=0.0)), ("pso_gen", dict(variant=5)),
                                     ("de1220", dict(ftol=0.0, xtol=0.0))])
def test_island_keeps_the_state_in_hbm(capi, ctx, name, kw):
    prob, x, f = _so_problem(capi, ctx, n=32)
    ids = np.arange(1, 33, dtype=np.uint64)
    half = capi.algo_desc(name, gens=3, seed=5, memory=1, **kw)
    isl = capi.Island(prob, 32)
    isl.upload(ids, x, f)
    isl.evolve(half)
    isl.evolve(half)
    _, xi, fi = isl.download()
    isl.close()
    # the same two calls with the state travelling through the host
    xa, fa, _, st = prob.evolve_memory(half, x, f, first_generation=1)
    xb, fb, _, _ = prob.evolve_memory(half, xa, fa, first_generation=4, state=st)
    assert np.array_equal(xi, xb) and np.array_equal(fi, fb)
    # and without memory the island re-draws it
    isl = capi.Island(prob, 32)
    isl.upload(ids, x, f)
    d0 = capi.algo_desc(name, gens=3, seed=5, memory=0, **kw)
    isl.evolve(d0)
    isl.evolve(d0)
    _, xn, _ = isl.download()
    isl.close()
    assert not np.array_equal(xn, xi)
    prob.close()


@pytest.mark.parametrize("name", ("cmaes", "xnes"))
def test_evolution_strategies_keep_their_distribution(capi, ctx, name):
    """cmaes (cmaes.cpp:201-228) and xnes (xnes.cpp:163-175) with memory = true keep sigma, the mean and the covariance factors: 3 + 3
    generations are the uninterrupted 6 (the exit tests aside, nothing else reads the population); without memory the second call starts
    again from the bounds' box.  cmaes restarts when the population size changes, xnes does not."""
    prob, x, f = _so_problem(capi, ctx, n=24)
    whole = capi.algo_desc(name, gens=6, seed=17, ftol=0.0, xtol=0.0)
    half = capi.algo_desc(name, gens=3, seed=17, ftol=0.0, xtol=0.0)
    x6, f6, done = prob.evolve(whole, x, f, first_generation=1)
    assert done == 6
    xa, fa, _, st = prob.evolve_memory(half, x, f, first_generation=1)
    x3, f3, _ = prob.evolve(half, x, f, first_generation=1)
    assert np.array_equal(xa, x3) and np.array_equal(fa, f3)
    xb, fb, _, st2 = prob.evolve_memory(half, xa, fa, first_generation=4, state=st)
    assert np.array_equal(xb, x6) and np.array_equal(fb, f6)
    xc, _, _ = prob.evolve(half, xa, fa, first_generation=4)
    assert not np.array_equal(xc, x6)
    assert st2["es"][0] == 1 and st2["es"][1] == prob.nx and not np.array_equal(st2["es"], st["es"])
    # a population of another size: cmaes starts afresh (= a memory-less call), xnes carries on
    xs, fs = xa[:16], fa[:16]
    xd, fd, _, _ = prob.evolve_memory(half, xs, fs, first_generation=4, state={k: (v if k == "es" else v[:16]) for k, v in st.items()})
    xe, fe, _ = prob.evolve(half, xs, fs, first_generation=4)
    assert np.array_equal(xd, xe) == (name == "cmaes")
    # inside a resident island the state stays with the island
    isl = capi.Island(prob, 24)
    isl.upload(np.arange(1, 25, dtype=np.uint64), x, f)
    hm = capi.algo_desc(name, gens=3, seed=17, ftol=0.0, xtol=0.0, memory=1)
    isl.evolve(hm)
    isl.evolve(hm)
    _, xi, fi = isl.download()
    isl.close()
    assert np.array_equal(xi, x6) and np.array_equal(fi, f6)
    prob.close()


def test_memory_argument_checks(capi, ctx):
    prob, x, f = _so_problem(capi, ctx, n=16)
    with pytest.raises(capi.PgcError):  # de keeps no state
        prob.evolve_memory(capi.algo_desc("de", gens=1, seed=1), x, f)
    with pytest.raises(capi.PgcError):  # sade.cpp:66-69
        prob.evolve_memory(capi.algo_desc("sade", gens=1, seed=1, variant_adptv=3), x, f)
    prob.close()
