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
