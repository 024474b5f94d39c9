"""GPU parity of the CMA-ES / xNES contractions (sampling, weighted mean, rank-mu / weighted Gram matrix) against the restated loops."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D", (1, 2, 7, 10, 30, 50, 100, 128, 129, 200, 444))
def test_rank_mu_and_mean(ctx, orc, D):
    rng = np.random.default_rng(D)
    for n, mu in ((8, 4), (200, 100), (4099, 2049), (65536, 32768 if D <= 50 else 4096))[: 4 if D <= 128 else 3]:
        x = rng.normal(size=(n, D)) * rng.uniform(0.1, 10, D)
        idx = rng.permutation(n)[:mu]
        w = np.log(mu + 0.5) - np.log(np.arange(mu) + 1.0)  # cmaes.cpp:166-169
        w /= w.sum()
        m_old, sigma = x.mean(axis=0), 0.37
        C, m = ctx.weighted_gram(x, w, idx=idx, center=m_old, scale_div=sigma * sigma)
        Co, mo = orc.weighted_gram(x, w, idx=idx, center=m_old, scale_div=sigma * sigma)
        assert np.array_equal(m, mo)  # same order, no fused multiply-add: bit-exact
        # entries are sums of signed products: bound against the size of the summed terms
        d = np.abs(x[idx] - m_old)
        scale = (d.T * w) @ d / (sigma * sigma)
        assert np.all(np.abs(C - Co) <= 1e-12 * scale + 1e-300), (D, n, np.abs(C - Co).max())
        assert np.all(np.abs(C - C.T) <= 1e-12 * scale + 1e-300)  # (w d_a) d_b vs (w d_b) d_a: symmetric up to rounding, like Eigen's
    # xnes form: rows = z, no centre, signed utilities
    z = rng.normal(size=(64, D))
    u = rng.normal(size=64)
    G, _ = ctx.weighted_gram(z, u, idx=rng.permutation(64))
    Go, _ = orc.weighted_gram(z, u, idx=None)
    assert G.shape == (D, D)


@pytest.mark.parametrize("D", (1, 5, 10, 50, 100, 128, 130, 444))
def test_sampling(ctx, orc, D):
    rng = np.random.default_rng(100 + D)
    q, _ = np.linalg.qr(rng.normal(size=(D, D)))
    bd = q * rng.uniform(0.5, 20, D)  # B * D
    mean = rng.normal(size=D) * 10
    for lam in (1, 9, 1000, 65536 if D <= 50 else (5000 if D <= 128 else 700)):
        x, z = ctx.cmaes_sample(mean, bd, 0.5, lam, seed=9, generation=3)
        xo, zo = orc.cmaes_sample(mean, bd, 0.5, lam, 9, 3)
        assert np.allclose(z, zo, rtol=1e-13, atol=1e-15)  # log / cos / sqrt differ by ulps between libdevice and glibc
        scale = np.abs(mean) + 0.5 * np.abs(zo) @ np.abs(bd).T
        assert np.all(np.abs(x - xo) <= 1e-12 * scale)
    assert abs(z.mean()) < 0.05 and abs(z.std() - 1) < 0.05


def test_unsupported_dimension_fails_loudly(ctx):
    from pagmo2_b200 import capi
    with pytest.raises(capi.PgcError):  # a 32-row chunk of full rows no longer fits in shared memory
        ctx.weighted_gram(np.zeros((4, 2000)), np.ones(4))


@pytest.mark.parametrize("fam,dim,lam", [("rosenbrock", 8, 20), ("rastrigin", 10, 32), ("ackley", 30, 64)])
def test_cmaes_evolve_follows_the_restated_loop(ctx, orc, fam, dim, lam):
    """cmaes::evolve (cmaes.cpp:111-407): the device loop (sampling, evaluation, recombination, rank-mu on the GPU; paths and
    Jacobi eigendecomposition on the host) against the triple-loop restatement consuming the same Philox normals.  The two
    Gram matrices differ by DMMA summation order (<= 1e-12 of the summed terms), which the adaptation feeds back every
    generation: the trajectories are compared over the first generations, the convergence over a long run."""
    from pagmo2_b200 import capi
    rng = np.random.default_rng(dim)
    prob = capi.Problem(ctx, fam, dim=dim)
    op = orc.problem(fam, dim=dim)
    lb, ub = prob.bounds()
    x = rng.uniform(lb, ub, (lam, dim))
    f = orc.simple(fam, x)
    kw = dict(sigma0=0.5, ftol=0.0, xtol=0.0, seed=11, first_generation=1)
    for gens, tol in ((1, 1e-11), (3, 1e-9), (8, 1e-6)):
        xg, fg, dg, sg = prob.cmaes_evolve(x, f, gens=gens, **kw)
        xo, fo, do, so = orc.cmaes_evolve(op, lb, ub, x, f, gens=gens, **kw)
        assert dg == do == gens
        scale = np.abs(xo).max()
        assert np.abs(xg - xo).max() <= tol * scale, (gens, np.abs(xg - xo).max())
        assert abs(sg - so) <= tol * so
        assert np.allclose(prob.eval_host(xg)[:, 0], fg, rtol=1e-12)
    # force_bounds keeps every sample inside the box (cmaes.cpp:301-315)
    xb, fb, _, _ = prob.cmaes_evolve(x, f, gens=5, force_bounds=True, **kw)
    assert (xb >= lb).all() and (xb <= ub).all()
    # a long run converges like the restated one does
    xg, fg, dg, sg = prob.cmaes_evolve(x, f, gens=400, sigma0=0.5, ftol=1e-10, xtol=1e-10, seed=11)
    xo, fo, do, so = orc.cmaes_evolve(op, lb, ub, x, f, gens=400, sigma0=0.5, ftol=1e-10, xtol=1e-10, seed=11)
    # (rastrigin and ackley are multimodal: both runs settle in a local optimum, rosenbrock's valley leads to the global one)
    assert fg.min() < 0.5 * f.min() and fo.min() < 0.5 * f.min()
    assert fam != "rosenbrock" or (fg.min() < 1e-4 and fo.min() < 1e-4)
    prob.close()


def test_cmaes_argument_checks(ctx):
    from pagmo2_b200 import capi
    prob = capi.Problem(ctx, "rastrigin", dim=5)
    x, f = np.zeros((8, 5)), np.zeros(8)
    for bad in (dict(cc=1.5), dict(cs=-0.5), dict(c1=2.0), dict(cmu=-2.0)):   # cmaes.cpp:64-88
        with pytest.raises(capi.PgcError):
            prob.cmaes_evolve(x, f, gens=1, **bad)
    with pytest.raises(capi.PgcError):                                         # :140-143: at least 5 individuals
        prob.cmaes_evolve(x[:4], f[:4], gens=1)
    mo = capi.Problem(ctx, "zdt", prob_id=1, dim=5)
    with pytest.raises(capi.PgcError):
        mo.cmaes_evolve(np.zeros((8, 5)), np.zeros(8), gens=1)


@pytest.mark.parametrize("fam,dim,lam", [("rosenbrock", 8, 20), ("rastrigin", 10, 32), ("ackley", 30, 64)])
def test_xnes_evolve_follows_the_restated_loop(ctx, orc, fam, dim, lam):
    """xnes::evolve (xnes.cpp:96-303): sampling, evaluation and the natural-gradient contractions on the GPU, the D x D updates
    (A <- A exp(d_A) through a Jacobi eigendecomposition) on the host, against the restatement consuming the same Philox normals."""
    from pagmo2_b200 import capi
    rng = np.random.default_rng(dim + 1)
    prob = capi.Problem(ctx, fam, dim=dim)
    op = orc.problem(fam, dim=dim)
    lb, ub = prob.bounds()
    x = rng.uniform(lb, ub, (lam, dim))
    f = orc.simple(fam, x)
    kw = dict(ftol=0.0, xtol=0.0, seed=13, first_generation=1)
    for gens, tol in ((1, 1e-11), (3, 1e-9), (8, 1e-6)):
        for extra in (dict(), dict(eta_mu=0.8, eta_sigma=0.3, eta_b=0.2, sigma0=0.3)):
            xg, fg, dg, sg = prob.xnes_evolve(x, f, gens=gens, **kw, **extra)
            xo, fo, do, so = orc.xnes_evolve(op, lb, ub, x, f, gens=gens, **kw, **extra)
            assert dg == do == gens
            scale = np.abs(xo).max()
            assert np.abs(xg - xo).max() <= tol * scale, (gens, np.abs(xg - xo).max())
            assert abs(sg - so) <= tol * so
            assert np.allclose(prob.eval_host(xg)[:, 0], fg, rtol=1e-12)
    xb, _, _, _ = prob.xnes_evolve(x, f, gens=5, force_bounds=True, **kw)
    assert (xb >= lb).all() and (xb <= ub).all()
    # through the descriptor (pgc_algo_evolve_device) and with the log (xnes.cpp:238-256): line `gen` = the population that generation made
    d = capi.algo_desc("xnes", gens=6, seed=13, ftol=0.0, xtol=0.0)
    xd, fd, done, log = prob.evolve_logged(d, x, f, 2)
    x6, f6, _, s6 = prob.xnes_evolve(x, f, gens=6, **kw)
    assert done == 6 and np.array_equal(xd, x6) and log[:, 0].tolist() == [1, 3, 5] and log[:, 1].tolist() == [lam, 3 * lam, 5 * lam]
    for row in log:
        _, fg, _, sg = prob.xnes_evolve(x, f, gens=int(row[0]), **kw)
        _, _, _, sprev = prob.xnes_evolve(x, f, gens=int(row[0]) - 1, **kw) if row[0] > 1 else (None, None, None, 0.5)
        assert row[2] == fg.min() and np.isclose(row[4], fg.max() - fg.min(), rtol=1e-13) and row[5] == sprev and row[3] > 0
    # a long run converges like the restated one does
    xg, fg, dg, sg = prob.xnes_evolve(x, f, gens=600, ftol=1e-10, xtol=1e-10, seed=13)
    xo, fo, do, so = orc.xnes_evolve(op, lb, ub, x, f, gens=600, ftol=1e-10, xtol=1e-10, seed=13)
    assert fg.min() < 0.5 * f.min() and fo.min() < 0.5 * f.min()
    prob.close()


def test_xnes_argument_checks(ctx):
    from pagmo2_b200 import capi
    prob = capi.Problem(ctx, "rastrigin", dim=5)
    x, f = np.zeros((8, 5)), np.zeros(8)
    for bad in (dict(eta_mu=1.5), dict(eta_sigma=0.0), dict(eta_b=-0.5), dict(sigma0=2.0)):   # xnes.cpp:55-78
        with pytest.raises(capi.PgcError):
            prob.xnes_evolve(x, f, gens=1, **bad)
    with pytest.raises(capi.PgcError):                                                         # :120-123
        prob.xnes_evolve(x[:3], f[:3], gens=1)
    mo = capi.Problem(ctx, "zdt", prob_id=1, dim=5)
    with pytest.raises(capi.PgcError):
        mo.xnes_evolve(np.zeros((8, 5)), np.zeros(8), gens=1)
