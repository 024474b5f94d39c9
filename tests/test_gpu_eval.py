"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the
CPU oracle on identical seeded inputs, against the committed golden vectors of the unmodified reference, and -
at BASELINE.json's full size - through size-independent properties.

Tolerance (north_star): FP64 benchmark functions agree within 1e-12 RELATIVE.  The differences that remain are
(i) fused multiply-add in the rotation (one rounding instead of two), (ii) libdevice-vs-glibc ulps of
cos/sin/exp/pow, (iii) pairwise instead of sequential summation order across the two lanes of an individual.
"""
import json
import os
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
GOLD = Path(__file__).resolve().parent / "golden"
OUT = ROOT / "gpurun_out"
REL_TOL = 1e-12
CEC_DIMS = (2, 10, 20, 30, 50, 100)


def cec_defined(func, dim):
    return not (dim == 2 and (17 <= func <= 22 or func >= 29))


def rel_err(got, want):
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)))


def make_cec(capi, ctx, orc, func, dim):
    mr, os_c, s = orc.cec2014_problem_tables(func, dim)
    return capi.Problem(ctx, "cec2014", prob_id=func, dim=dim, rotation=mr, shift=os_c, shuffle=s)


@pytest.fixture(scope="module")
def capi():
    from pagmo2_b200 import capi as m
    return m


_report = {}


def _dump_report():
    try:
        OUT.mkdir(exist_ok=True)
        (OUT / "parity_report.json").write_text(json.dumps(_report, indent=1, sort_keys=True))
    except OSError:
        pass


@pytest.mark.parametrize("dim", CEC_DIMS)
def test_cec2014_parity_vs_oracle(capi, ctx, orc, dim):
    rng = np.random.default_rng(1000 + dim)
    n = 1003  # ragged: not a multiple of the 16-individual warp tile
    worst = {}
    for func in range(1, 31):
        if not cec_defined(func, dim):
            continue
        prob = make_cec(capi, ctx, orc, func, dim)
        xs = np.vstack([rng.uniform(-100, 100, (n - 3, dim)), rng.normal(0, 1e-2, (2, dim)), np.zeros((1, dim))])
        got = prob.eval_host(xs)[:, 0]
        want = orc.cec2014(func, xs, nthreads=8)
        worst[func] = rel_err(got, want)
        prob.close()
    _report[f"cec2014_d{dim}_max_rel_err"] = worst
    _dump_report()
    bad = {f: e for f, e in worst.items() if not e <= REL_TOL}
    assert not bad, f"D={dim}: functions over {REL_TOL}: {bad}"


def test_cec2014_vs_reference_golden(capi, ctx, orc):
    g = np.load(GOLD / "cec2014_ref.npz")
    worst = {}
    for dim in (10, 30, 100):
        for func in range(1, 31):
            prob = make_cec(capi, ctx, orc, func, dim)
            got = prob.eval_host(g[f"x_f{func}_d{dim}"])[:, 0]
            want = g[f"f_f{func}_d{dim}"]
            # rows 6 and 7 are x = shift (f == bias up to the reference's own residue) and x = 0
            worst[f"f{func}_d{dim}"] = rel_err(got, want)
            prob.close()
    _report["cec2014_golden_max_rel_err"] = worst
    _dump_report()
    bad = {k: e for k, e in worst.items() if not e <= REL_TOL}
    assert not bad, bad


def test_cec2014_minimum_is_bias(capi, ctx, orc):
    # reference tests/cec2014.cpp:111-121 at D=10: f(origin shift) == 100*func exactly
    for func in range(1, 31):
        prob = make_cec(capi, ctx, orc, func, 10)
        _, os_c, _ = orc.cec2014_problem_tables(func, 10)
        f = prob.eval_host(os_c[:10][None, :])[0, 0]
        assert f == 100.0 * func, (func, f)
        prob.close()


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 31, 257])
def test_cec2014_ragged_and_empty_batches(capi, ctx, orc, n):
    rng = np.random.default_rng(n)
    for func, dim in ((1, 100), (22, 50), (27, 20)):
        prob = make_cec(capi, ctx, orc, func, dim)
        xs = rng.uniform(-100, 100, (n, dim))
        got = prob.eval_host(xs)
        assert got.shape == (n, 1)
        if n:
            assert rel_err(got[:, 0], orc.cec2014(func, xs)) <= REL_TOL
        prob.close()


def test_cec2014_device_path_unaligned_and_offset(capi, ctx, orc):
    """pgc_eval_device on a device pointer that is only 8-byte aligned (a shard that starts mid-allocation)."""
    rng = np.random.default_rng(5)
    dim, n = 30, 500
    prob = make_cec(capi, ctx, orc, 9, dim)
    xs = rng.uniform(-100, 100, (n, dim))
    flat = np.concatenate([[0.0], xs.ravel()])  # shift by one double: 8-byte aligned only
    d_in = ctx.to_device(flat)
    d_out = ctx.malloc(8 * n)
    prob.eval_device(d_in + 8, n, d_out)
    ctx.synchronize()
    got = ctx.from_device(d_out, (n,))
    assert rel_err(got, orc.cec2014(9, xs)) <= REL_TOL
    # and the aligned path must give bit-identical results
    d_in2 = ctx.to_device(xs)
    prob.eval_device(d_in2, n, d_out)
    ctx.synchronize()
    assert np.array_equal(got, ctx.from_device(d_out, (n,)))
    for p in (d_in, d_in2, d_out):
        ctx.free(p)
    prob.close()


def test_cec2014_rejects_bad_arguments(capi, ctx, orc):
    mr, os_c, s = orc.cec2014_problem_tables(1, 10)
    for func, dim in ((0, 2), (29, 2), (10, 3), (31, 10)):  # reference tests/cec2014.cpp:70-72
        with pytest.raises(capi.PgcError) as ei:
            capi.Problem(ctx, "cec2014", prob_id=func, dim=dim, rotation=mr, shift=os_c, shuffle=s)
        assert ei.value.status == capi.PGC_ERR_INVALID_ARGUMENT
    with pytest.raises(capi.PgcError):  # tables too short
        capi.Problem(ctx, "cec2014", prob_id=23, dim=10, rotation=mr[:100], shift=os_c, shuffle=s)
    capi.FAMILY["golomb_ruler"] = 12  # a UDP family without a device evaluator: error, not a CPU fallback
    try:
        with pytest.raises(capi.PgcError) as ei:
            capi.Problem(ctx, "golomb_ruler", dim=5)
        assert ei.value.status == capi.PGC_ERR_UNSUPPORTED
    finally:
        del capi.FAMILY["golomb_ruler"]
    prob = make_cec(capi, ctx, orc, 5, 10)
    assert prob.name == "CEC2014 - f5(ackley_func)"
    lb, ub = prob.bounds()
    assert (lb == -100).all() and (ub == 100).all() and prob.nx == 10 and prob.nf == 1
    with pytest.raises(ValueError):
        prob.eval_host(np.zeros(15))
    prob.close()


@pytest.mark.parametrize("fam", ["rastrigin", "ackley", "griewank", "schwefel", "rosenbrock"])
def test_simple_udps_parity(capi, ctx, orc, fam):
    rng = np.random.default_rng(11)
    g = np.load(GOLD / "simple_ref.npz")
    worst = {}
    for dim in (1, 2, 5, 10, 31, 32, 33, 64, 100, 333):
        if fam == "rosenbrock" and dim < 2:
            continue
        prob = capi.Problem(ctx, fam, dim=dim)
        lb, ub = prob.bounds()
        xs = rng.uniform(lb, ub, (777, dim))
        got = prob.eval_host(xs)[:, 0]
        want = orc.simple(fam, xs)
        denom = np.maximum(np.abs(want), 1e-9)  # rosenbrock/ackley can be ~0 near the optimum
        worst[dim] = float(np.max(np.abs(got - want) / denom))
        key = f"x_{fam}_d{dim}"
        if key in g.files:
            gg = prob.eval_host(g[key])[:, 0]
            assert np.all(np.abs(gg - g[f"f_{fam}_d{dim}"]) <= REL_TOL * np.maximum(np.abs(g[f"f_{fam}_d{dim}"]), 1e-3)), (fam, dim)
        prob.close()
    _report[f"simple_{fam}_max_rel_err"] = worst
    _dump_report()
    assert all(e <= REL_TOL for e in worst.values()), worst
    # reference KATs: tests/rastrigin.cpp:56-57, rosenbrock.cpp:60-61 (exact)
    if fam == "rastrigin":
        assert capi.Problem(ctx, fam, dim=5).eval_host(np.ones((1, 5)))[0, 0] == 5.0
    if fam == "rosenbrock":
        assert capi.Problem(ctx, fam, dim=5).eval_host(np.ones((1, 5)))[0, 0] == 0.0
    with pytest.raises(capi.PgcError):
        capi.Problem(ctx, fam, dim=0)


def test_full_size_properties(capi, ctx, orc):
    """BASELINE size (1 Mi x D=100): results do not depend on where in the batch (tile, warp, CTA) an individual
    sits, the device and host paths agree bit for bit, and a seeded sample agrees with the oracle."""
    n, dim = 1 << 20, 100
    rng = np.random.default_rng(99)
    xs = rng.uniform(-100, 100, (n, dim))
    perm = rng.permutation(n)
    sample = rng.choice(n, 4096, replace=False)
    d_x = ctx.to_device(xs)
    d_f = ctx.malloc(8 * n)
    for func in (5, 11, 20, 27, 30):
        prob = make_cec(capi, ctx, orc, func, dim)
        prob.eval_device(d_x, n, d_f)
        ctx.synchronize()
        f_dev = ctx.from_device(d_f, (n,))
        f_host = prob.eval_host(xs)[:, 0]
        assert np.array_equal(f_dev, f_host), func
        f_perm = prob.eval_host(xs[perm])[:, 0]
        assert np.array_equal(f_perm, f_dev[perm]), func
        want = orc.cec2014(func, xs[sample], nthreads=8)
        assert rel_err(f_dev[sample], want) <= REL_TOL, func
        assert np.isfinite(f_dev).all() and (f_dev >= 100.0 * func).all()  # bias is the global minimum
        prob.close()
    ctx.free(d_x)
    ctx.free(d_f)


def test_zdt_dtlz_parity(capi, ctx, orc):
    rng = np.random.default_rng(21)
    g = np.load(GOLD / "mo_ref.npz")
    worst = {}
    for pid in range(1, 7):
        for param in (2, 11, 30, 100):
            prob = capi.Problem(ctx, "zdt", prob_id=pid, dim=param)
            lb, ub = prob.bounds()
            xs = rng.uniform(lb, ub, (1001, prob.nx))
            got, want = prob.eval_host(xs), orc.zdt(pid, xs)
            assert got.shape == (1001, 2)
            worst[f"zdt{pid}_p{param}"] = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-9)))
            key = f"zdt{pid}_p{param}"
            if "x_" + key in g.files:
                gg = prob.eval_host(g["x_" + key])
                assert np.all(np.abs(gg - g["f_" + key]) <= REL_TOL * np.maximum(np.abs(g["f_" + key]), 1e-9)), key
            prob.close()
    for pid in range(1, 8):
        for dim, fdim, alpha in ((5, 3, 100), (12, 3, 100), (7, 2, 3), (30, 5, 100), (9, 8, 10)):
            prob = capi.Problem(ctx, "dtlz", prob_id=pid, dim=dim, nobj=fdim, param=alpha)
            xs = rng.uniform(0, 1, (1001, dim))
            got, want = prob.eval_host(xs), orc.dtlz(pid, xs, fdim, alpha)
            assert got.shape == (1001, fdim)
            worst[f"dtlz{pid}_d{dim}_m{fdim}"] = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-9)))
            prob.close()
    _report["mo_max_rel_err"] = worst
    _dump_report()
    bad = {k: e for k, e in worst.items() if not e <= REL_TOL}
    assert not bad, bad
    # reference KATs, tests/zdt.cpp:67-80
    f = capi.Problem(ctx, "zdt", prob_id=1, dim=30).eval_host(np.full((1, 30), 0.25))[0]
    assert f[0] == 0.25 and abs(f[1] - 2.3486121811340026) <= 1e-15 * 2.35
    for bad_args in (dict(prob_id=0, dim=30), dict(prob_id=7, dim=30), dict(prob_id=1, dim=1)):
        with pytest.raises(capi.PgcError):
            capi.Problem(ctx, "zdt", **bad_args)
    for bad_args in (dict(prob_id=0, dim=5, nobj=3), dict(prob_id=8, dim=5, nobj=3), dict(prob_id=1, dim=3, nobj=3),
                     dict(prob_id=1, dim=5, nobj=1)):
        with pytest.raises(capi.PgcError):
            capi.Problem(ctx, "dtlz", **bad_args)
