"""The Bringmann-Friedrich approximations on the device (hv_approx.cu): bf_fpras (hv_bf_fpras.cpp:91-146) and bf_approx
(hv_bf_approx.cpp:337-470).  They are Monte-Carlo estimators - the reference draws from one mt19937, the device from Philox
substreams - so parity is what the algorithms promise: the (eps, delta) bound against the exact hypervolume for fpras, and for approx
the exact extreme contributor up to the factor 1 + eps (exact values from the device's own sweeps / WFG, which are pinned elsewhere)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _front(rng, n, m):
    f = rng.uniform(0.05, 1, (n, m))
    return f / np.linalg.norm(f, axis=1, keepdims=True)


@pytest.mark.parametrize("m", (2, 3, 5, 8))
def test_fpras_within_eps_of_the_exact_hypervolume(ctx, m):
    rng = np.random.default_rng(m)
    for n, kind in ((1, "front"), (40, "front"), (200, "random")):
        f = _front(rng, n, m) if kind == "front" else rng.uniform(0, 1, (n, m))
        r = np.full(m, 1.2)
        exact = ctx.hv_compute(f, r)
        for seed in (1, 2, 3):
            got = ctx.hv_fpras(f, r, eps=0.02, delta=0.01, seed=seed)
            assert abs(got - exact) <= 0.02 * exact, (m, n, kind, seed, got, exact)
    # different seeds give different estimates, the same seed the same one
    a, b, c = ctx.hv_fpras(f, r, 0.05, 0.05, 7), ctx.hv_fpras(f, r, 0.05, 0.05, 7), ctx.hv_fpras(f, r, 0.05, 0.05, 8)
    assert a == b and a != c


@pytest.mark.parametrize("m", (2, 3, 4, 5))
@pytest.mark.parametrize("use_exact", (True, False))
def test_approx_extreme_contributors(ctx, m, use_exact):
    rng = np.random.default_rng(10 + m)
    eps = 0.05
    for n in (2, 12, 60 if (use_exact or m < 5) else 20):  # pure sampling in 5 objectives: the greatest contributor of 60 points takes ~25 s
        f = _front(rng, n, m)
        r = np.full(m, 1.2)
        c = ctx.hv_contributions(f, r)
        lo = ctx.hv_approx_extreme(f, r, greatest=False, use_exact=use_exact, eps=eps, delta=1e-4, seed=n)
        hi = ctx.hv_approx_extreme(f, r, greatest=True, use_exact=use_exact, eps=eps, delta=1e-4, seed=n)
        assert c[lo] <= (1 + eps) * c.min() + 1e-15, (m, n, lo, int(np.argmin(c)))
        assert c[hi] * (1 + eps) >= c.max(), (m, n, hi, int(np.argmax(c)))
    # a dominated or duplicated point contributes nothing: the least contributor is found before any sampling (:383-390)
    f = _front(rng, 20, m)
    f[13] = f[4] + 0.01
    assert ctx.hv_approx_extreme(f, np.full(m, 1.3), greatest=False, use_exact=use_exact, seed=1) in (13,)
    f[13] = f[4]
    assert ctx.hv_approx_extreme(f, np.full(m, 1.3), greatest=False, use_exact=use_exact, seed=1) in (4, 13)


def test_approx_argument_checks(ctx):
    from pagmo2_b200 import capi
    f = np.array([[0.2, 0.8], [0.8, 0.2]])
    with pytest.raises(capi.PgcError):
        ctx.hv_fpras(f, [1.0, 1.0], eps=0.0)
    with pytest.raises(capi.PgcError):
        ctx.hv_fpras(f, [1.0, 1.0], delta=1.5)
    with pytest.raises(capi.PgcError):
        ctx.hv_fpras(f, [0.5, 1.0])  # a point outside the reference point
    with pytest.raises(capi.PgcError):
        ctx.hv_approx_extreme(f, [1.0, 1.0], eps=-0.1)
    assert ctx.hv_approx_extreme(f[:1], [1.0, 1.0]) == 0
