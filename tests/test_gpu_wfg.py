"""GPU parity of the WFG1..9 evaluator against the restated oracle, the reference's golden vectors and its own known answers."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
REL_TOL = 1e-12


@pytest.fixture(scope="module")
def capi():
    from pagmo2_b200 import capi as m
    return m


def close(got, want):
    # objectives are sums of O(1) terms (f_i = t_M + 2i * shape_i): bound relative to max(|f|, 1)
    return np.all((np.abs(got - want) <= REL_TOL * np.maximum(np.abs(want), 1.0)) | (np.isnan(got) & np.isnan(want)))


@pytest.mark.parametrize("pid", range(1, 10))
def test_wfg_parity_vs_oracle(capi, ctx, orc, pid):
    rng = np.random.default_rng(900 + pid)
    for n, m, k in ((9, 5, 8), (10, 5, 8), (5, 3, 4), (12, 3, 4), (30, 4, 6), (4, 2, 2), (24, 2, 4), (8, 3, 2), (40, 3, 10)):
        if pid in (2, 3) and (n - k) % 2:
            continue
        prob = capi.Problem(ctx, "wfg", prob_id=pid, dim=n, nobj=m, param=k)
        ub = 2.0 * (np.arange(n) + 1)
        assert prob.name == f"WFG{pid}" and prob.nobj == m and np.array_equal(prob.bounds()[1], ub)
        xs = np.vstack([rng.uniform(0, 1, (1001, n)) * ub, np.zeros((1, n)), ub[None, :], 0.35 * ub[None, :]])
        got = prob.eval_host(xs)
        assert got.shape == (1004, m) and close(got, orc.wfg(pid, xs, m, k)), (pid, n, m, k)
        prob.close()


def test_wfg_golden_and_known_answers(capi, ctx):
    g = np.load(GOLD / "wfg_ref.npz")
    for kx in (k for k in g.files if k.startswith("x_")):
        pid, n, m, k = (int(v) for v in kx[5:].split("_"))
        prob = capi.Problem(ctx, "wfg", prob_id=pid, dim=n, nobj=m, param=k)
        assert close(prob.eval_host(g[kx]), g["f" + kx[1:]]), kx
        prob.close()
    # reference tests/wfg.cpp:75-84
    prob = capi.Problem(ctx, "wfg", prob_id=1, dim=9, nobj=5, param=8)
    assert np.allclose(prob.eval_host(np.full((1, 9), 2.0))[0], [2.67637472191165, 1.00059019674296, 1.00158344827345, 0.999721693168825,
                                                                 0.994938703521363], rtol=1e-8)
    prob.close()
    for bad in ((0, 5, 3, 4), (10, 5, 3, 4), (1, 5, 1, 4), (1, 5, 3, 5), (1, 5, 3, 3), (2, 9, 3, 4), (3, 9, 3, 4)):  # tests/wfg.cpp:52-61
        with pytest.raises(capi.PgcError):
            capi.Problem(ctx, "wfg", prob_id=bad[0], dim=bad[1], nobj=bad[2], param=bad[3])
    prob = capi.Problem(ctx, "wfg", prob_id=4, dim=12, nobj=3, param=4)
    assert prob.eval_host(np.zeros((0, 12))).shape == (0, 3)
    prob.close()
