"""GPU parity of the CEC2013 evaluator (through the C ABI) against the restated oracle and the reference's golden vectors.

Tolerance: 1e-12 relative (north_star) - EXCEPT where the reference function itself is ill-conditioned at the test point.
asyfunc (cec2013.cpp:1053-1059) raises coordinates to powers up to ~x^8, after which schaffer_F7 / ackley / escaffer6 take
sin/cos of arguments whose last-bit spacing exceeds the period: the reference's own value then moves by far more than 1e-12
when an input moves in its last bits.  For such points the bound is 32x the oracle's own response to input perturbations of at
most 4 ulp(100) per coordinate (measured here, per point, 8 random draws); the share of points that needed it is written to the parity report."""
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
GOLD = Path(__file__).resolve().parent / "golden"
OUT = ROOT / "gpurun_out"
REL_TOL = 1e-12
DIMS = (2, 5, 10, 20, 30, 40, 50, 60, 70, 80, 90, 100)


@pytest.fixture(scope="module")
def capi():
    from pagmo2_b200 import capi as m
    return m


def make13(capi, ctx, orc, func, dim):
    mr, os_ = orc.cec2013_tables(dim)
    return capi.Problem(ctx, "cec2013", prob_id=func, dim=dim, rotation=mr, shift=os_)


def noise_floor(orc, func, xs, want, rng, reps=8):
    """max |f(x~) - f(x)| of the ORACLE over `reps` inputs x~ = x + k * ulp(100), k uniform in -4..4 per coordinate: a few units of
    the absolute resolution of the [-100, 100] box (an ulp of x itself says nothing at x = 0, where x - Os absorbs it)."""
    k = np.zeros_like(want)
    for _ in range(reps):
        xp = xs + rng.integers(-4, 5, xs.shape) * np.spacing(100.0)
        k = np.maximum(k, np.abs(orc.cec2013(func, xp) - want))
    return k


def check(orc, func, xs, got, want, rng):
    err = np.abs(got - want)
    # schwefel_func returns 418.98...*nx + (sum of terms) (cec2013.cpp:699): near its optimum the value is the difference of two numbers
    # of that size, so the relative bound applies to that size (the same holds for the reference's own rounding)
    scale = 4.189828872724338e+002 * xs.shape[1] if func in (14, 15, 22, 23, 24, 25, 26, 27, 28) else 0.0
    strict = err <= REL_TOL * np.maximum(np.abs(want), scale)
    if strict.all():
        return 0.0, float((err / np.abs(want)).max())
    loose = err <= 32.0 * noise_floor(orc, func, xs, want, rng)
    loose &= err <= 1e-6 * np.abs(want) if func not in (8, 20) else True  # hard cap; ackley / escaffer6 after asy are chaotic
    assert (strict | loose).all(), (func, xs.shape[1], float(err[~(strict | loose)].max()), want[~(strict | loose)][:3])
    return float((~strict).mean()), float((err[strict] / np.abs(want[strict])).max()) if strict.any() else 0.0


_report = {}


@pytest.mark.parametrize("dim", DIMS)
def test_cec2013_parity_vs_oracle(capi, ctx, orc, dim):
    rng = np.random.default_rng(1300 + dim)
    _, os_ = orc.cec2013_tables(dim)
    n = 203 if dim > 50 else 403  # ragged: not a multiple of the 8-individual warp tile
    rep = {}
    for func in range(1, 29):
        prob = make13(capi, ctx, orc, func, dim)
        xs = np.vstack([rng.uniform(-100, 100, (n - 103, dim)), os_[:dim] + rng.normal(0, 1.0, (100, dim)), os_[None, :dim],
                        np.zeros((1, dim)), rng.uniform(-100, 100, (1, dim))])
        xs[-1, 0] = os_[0]
        got = prob.eval_host(xs)[:, 0]
        want = orc.cec2013(func, xs)
        rep[func] = check(orc, func, xs, got, want, rng)
        prob.close()
    _report[f"cec2013_d{dim}"] = {f: {"ill_conditioned_share": a, "max_rel_err_elsewhere": b} for f, (a, b) in rep.items()}
    try:
        OUT.mkdir(exist_ok=True)
        (OUT / "parity_report_cec2013.json").write_text(json.dumps(_report, indent=1, sort_keys=True))
    except OSError:
        pass
    # well-conditioned families must not need the relaxation at all
    for func in (1, 2, 4, 5, 6, 10, 11, 14, 15, 17, 19):
        assert rep[func][0] == 0.0, (func, rep[func])


def test_cec2013_vs_reference_golden(capi, ctx, orc):
    g = np.load(GOLD / "cec2013_ref.npz")
    rng = np.random.default_rng(5)
    for dim in (10, 30, 50):
        for func in range(1, 29):
            prob = make13(capi, ctx, orc, func, dim)
            xs, want = g[f"x_f{func}_d{dim}"], g[f"f_f{func}_d{dim}"]
            check(orc, func, xs, prob.eval_host(xs)[:, 0], want, rng)
            prob.close()


@pytest.mark.parametrize("n", (0, 1, 7, 8, 9, 1185))
def test_cec2013_ragged_and_empty_batches(capi, ctx, orc, n):
    rng = np.random.default_rng(n)
    for func in (3, 12, 18, 27):
        prob = make13(capi, ctx, orc, func, 30)
        _, os_ = orc.cec2013_tables(30)
        xs = os_[:30] + rng.normal(0, 2.0, (n, 30))
        got = prob.eval_host(xs)
        assert got.shape == (n, 1)
        if n:
            check(orc, func, xs, got[:, 0], orc.cec2013(func, xs), rng)
        prob.close()


def test_cec2013_large_batch_is_consistent(capi, ctx, orc):
    """cfg5-sized and larger batches: the same rows give the same values wherever they sit in the batch (tiles are independent)."""
    rng = np.random.default_rng(77)
    for func, dim in ((12, 50), (28, 50), (9, 100)):
        prob = make13(capi, ctx, orc, func, dim)
        base = rng.uniform(-100, 100, (1024, dim))
        big = np.tile(base, (64, 1))
        got = prob.eval_host(big)[:, 0].reshape(64, 1024)
        assert np.array_equal(got, np.tile(got[0], (64, 1)))
        assert np.array_equal(got[0], prob.eval_host(base)[:, 0])
        prob.close()


def test_cec2013_metadata_and_bad_arguments(capi, ctx, orc):
    mr, os_ = orc.cec2013_tables(10)
    p = capi.Problem(ctx, "cec2013", prob_id=9, dim=10, rotation=mr, shift=os_)
    assert p.name == "CEC2013 - f9(weierstrass_func)" and p.nx == 10 and p.nobj == 1
    lb, ub = p.bounds()
    assert (lb == -100).all() and (ub == 100).all()
    p.close()
    for func, dim in ((0, 10), (29, 10), (1, 3), (1, 101)):  # reference cec2013.cpp:53-63
        with pytest.raises(capi.PgcError):
            capi.Problem(ctx, "cec2013", prob_id=func, dim=dim, rotation=mr, shift=os_)
    with pytest.raises(capi.PgcError):
        capi.Problem(ctx, "cec2013", prob_id=3, dim=10, rotation=mr[:100], shift=os_)  # needs 2 matrices
    with pytest.raises(capi.PgcError):
        capi.Problem(ctx, "cec2013", prob_id=21, dim=10, rotation=mr, shift=os_[:10])  # needs 5 shifts


# strict mode, measured on B200 over all 28 functions x D in {10, 30, 50, 100} (scripts/cec2013_strict_report.py, profiles/r2f_*):
# every function meets 1e-12 on every point except the three below, where what is left is libdevice-vs-glibc pow (2 ulp vs < 1 ulp)
# feeding a trigonometric term whose argument is huge after asyfunc: f7/f28 take sin(50 z^0.2) of z up to 1e12 (an ulp of z^0.2 moves
# the argument by 1e-12), f8 takes cos(2 pi z) of z up to 1e18 - there ulp(z) = 512 and the term is a function of pow's last bit.
STRICT_CAP = {7: 2e-11, 28: 2e-11}


@pytest.mark.parametrize("dim", (10, 30, 50, 100))
def test_cec2013_strict_mode_meets_the_tolerance_on_ill_conditioned_functions(capi, ctx, orc, dim):
    """pgc_problem_set_strict: the rotations accumulate in the reference's own order (rotatefunc, cec2013.cpp:1046-1051), so the
    rotated vectors are bit-identical and only libdevice-vs-glibc ulps remain.  Every function then meets REL_TOL on every point,
    ill-conditioned points included, except f7 / f28 (hard cap 2e-11, five times tighter than the default mode's worst) and f8
    (chaotic after asyfunc: held to the oracle's own noise floor, and to fewer exceedances than the default mode)."""
    rng = np.random.default_rng(1400 + dim)
    _, os_ = orc.cec2013_tables(dim)
    n = 403
    xs = np.vstack([rng.uniform(-100, 100, (n - 103, dim)), os_[:dim] + rng.normal(0, 1.0, (100, dim)), os_[None, :dim], np.zeros((1, dim)),
                    rng.uniform(-100, 100, (1, dim))])
    worst = {}
    for func in range(1, 29):   # strict mode is a mode of every function
        prob = make13(capi, ctx, orc, func, dim)
        want = orc.cec2013(func, xs)
        loose = prob.eval_host(xs)[:, 0]
        prob.set_strict(True)
        got = prob.eval_host(xs)[:, 0]
        prob.set_strict(False)
        assert np.array_equal(prob.eval_host(xs)[:, 0], loose)   # the switch is a switch
        scale = 4.189828872724338e+002 * dim if func in (14, 15, 22, 23, 24, 25, 26, 27, 28) else 0.0
        den = np.maximum(np.abs(want), scale)
        rel, rel_loose = np.abs(got - want) / den, np.abs(loose - want) / den
        worst[func] = float(rel.max())
        if func == 8:
            over = rel > REL_TOL
            assert over.sum() <= (rel_loose > REL_TOL).sum()
            assert (np.abs(got - want)[over] <= 32.0 * noise_floor(orc, func, xs, want, rng)[over]).all()
        else:
            assert rel.max() <= STRICT_CAP.get(func, REL_TOL), (func, dim, float(rel.max()))
        prob.close()
    _report[f"cec2013_strict_d{dim}"] = worst
    try:
        OUT.mkdir(exist_ok=True)
        (OUT / "parity_report_cec2013.json").write_text(json.dumps(_report, indent=1, sort_keys=True))
    except OSError:
        pass
    other = capi.Problem(ctx, "rastrigin", dim=5)
    with pytest.raises(capi.PgcError):
        other.set_strict(True)


@pytest.mark.parametrize("dim", (10, 50))
def test_cec2013_value_does_not_depend_on_the_batch_size(capi, ctx, orc, dim):
    """island-sized batches run with 1, 2 or 4 individuals per warp tile instead of 8 (launch13): an individual's fitness is the
    same bits whatever batch it arrives in."""
    rng = np.random.default_rng(77 + dim)
    xs = rng.uniform(-100, 100, (5000, dim))
    for func in (1, 7, 12, 15, 19, 23, 28):
        prob = make13(capi, ctx, orc, func, dim)
        full = prob.eval_host(xs)[:, 0]                      # 8 per tile
        for n in (3000, 2000, 500, 7):                       # 4, 2, 1, 1 per tile
            assert np.array_equal(prob.eval_host(xs[:n])[:, 0], full[:n]), (func, n)
        prob.close()
