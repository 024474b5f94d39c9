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


@pytest.mark.parametrize("family,algo,kw", [("rastrigin", "de1220", dict(variant_adptv=1)), ("rastrigin", "de", dict(variant=7)),
                                            ("cec2013", "sade", dict(variant=2, variant_adptv=1))])
def test_generation_graph_replay_equals_plain_launches(capi, ctx, orc, monkeypatch, family, algo, kw):
    """From the second evolve() on a population the generation loop replays an instantiated CUDA graph (de.cu): three calls with
    graphs == the same three calls with plain launches (PGC_GRAPHS=0) == the restated loop, the exit conditions included."""
    import ctypes as C
    rng = np.random.default_rng(4242)
    NP, dim = 256, 10
    if family == "cec2013":
        mr, os_ = orc.cec2013_tables(dim)
        make = lambda: capi.Problem(ctx, "cec2013", prob_id=12, dim=dim, rotation=mr, shift=os_)  # noqa: E731
        op = orc.problem("cec2013", prob_id=12, dim=dim, tables=(mr, os_))
    else:
        make = lambda: capi.Problem(ctx, family, dim=dim)  # noqa: E731
        op = orc.problem(family, dim=dim)
    code = {"de": 0, "sade": 1, "de1220": 2}[algo]
    al = np.array([2, 3, 7, 10, 13, 14, 15, 16], dtype=np.uint32)

    monkeypatch.setenv("PGC_DE_RESIDENT", "0")  # the simple UDPs would otherwise take the resident loop (tested below)

    def three_calls(graphs, ftol=0.0):
        monkeypatch.setenv("PGC_GRAPHS", "1" if graphs else "0")
        prob = make()
        lb, ub = prob.bounds()
        x = np.random.default_rng(7).uniform(lb, ub, (NP, dim))
        f = prob.eval_host(x)[:, 0]
        dx, df = ctx.to_device(x), ctx.to_device(f)
        l0, gens_done = ctx.launches, []
        for call in range(3):
            done = C.c_uint()
            capi.check(capi.lib().pgc_de_evolve_device(prob._h, dx, df, NP, 12, code, kw.get("variant", 2), kw.get("variant_adptv", 1), 0.8, 0.9,
                                                       al.ctypes.data_as(C.c_void_p), al.size, ftol, 0.0, None, None, None, 31, 1 + 12 * call,
                                                       C.byref(done), None))
            gens_done.append(done.value)
        out = ctx.from_device(dx, x.shape), ctx.from_device(df, f.shape), gens_done, ctx.launches - l0
        ctx.free(dx)
        ctx.free(df)
        prob.close()
        return (x, f, lb, ub) + out

    x, f, lb, ub, xg, fg, gg, lg = three_calls(True)
    _, _, _, _, xp, fp, gp, lp = three_calls(False)
    assert gg == gp == [12, 12, 12] and lg == lp  # the replayed launches are counted like the plain ones
    assert np.array_equal(xg, xp) and np.array_equal(fg, fp)
    xo, fo = x, f
    for call in range(3):
        xo, fo, go, *_ = orc.de_evolve(op, lb, ub, xo, fo, gens=12, algo=algo, seed=31, first_generation=1 + 12 * call, ftol=0.0, xtol=0.0, **kw)
    assert np.allclose(xg, xo, rtol=1e-9, atol=1e-12) and np.allclose(fg, fo, rtol=1e-9)
    # an exit condition that fires inside a replayed batch: same generation count as the plain path
    big = float(np.abs(fg.max() - fg.min()) * 4)
    *_, g1, _ = three_calls(True, ftol=big)
    *_, g0, _ = three_calls(False, ftol=big)
    assert g1 == g0 and g1[-1] < 12


@pytest.mark.parametrize("family", ("rastrigin", "ackley", "griewank", "schwefel", "rosenbrock"))
def test_resident_loop_equals_one_launch_per_phase(capi, ctx, orc, monkeypatch, family):
    """Simple UDPs at island-sized populations run all generations of an evolve() in one cooperative launch (de_resident_kernel):
    same population, fitness and generation count, bit for bit, as the trial / evaluate / select launches (PGC_DE_RESIDENT=0),
    odd and even generation counts (the result lives in the second copy after an odd number), exit conditions included."""
    NP, dim = 300, 13  # ragged: not a multiple of the 8 warps of a CTA
    prob = capi.Problem(ctx, family, dim=dim)
    op = orc.problem(family, dim=dim)
    lb, ub = prob.bounds()
    x = np.random.default_rng(3).uniform(lb, ub, (NP, dim))
    f = prob.eval_host(x)[:, 0]
    for algo, kw in (("de", dict(variant=2)), ("de", dict(variant=8)), ("sade", dict(variant=12, variant_adptv=1)),
                     ("sade", dict(variant=3, variant_adptv=2)), ("de1220", dict(variant_adptv=1)), ("de1220", dict(variant_adptv=2))):
        for gens, ftol in ((7, 0.0), (10, 0.0), (40, float(np.ptp(f)) * 1.5)):  # the last one: |worst - best| < ftol at once
            args = dict(gens=gens, algo=algo, seed=9, first_generation=5, ftol=ftol, xtol=0.0, **kw)
            monkeypatch.setenv("PGC_DE_RESIDENT", "1")
            l0 = ctx.launches
            xr, fr, gr = prob.de_evolve(x, f, **args)
            resident_launches = ctx.launches - l0
            monkeypatch.setenv("PGC_DE_RESIDENT", "0")
            xp, fp, gp = prob.de_evolve(x, f, **args)
            assert gr == gp and (ftol == 0.0) == (gr == gens), (algo, kw, gens, gr, gp)
            assert np.array_equal(xr, xp) and np.array_equal(fr, fp), (algo, kw, gens)
            assert resident_launches <= 3  # init (global best) + the resident loop
            if ftol == 0.0 and gens == 7:
                xo, fo, go, *_ = orc.de_evolve(op, lb, ub, x, f, **args)
                assert np.allclose(xr, xo, rtol=1e-9, atol=1e-12) and np.allclose(fr, fo, rtol=1e-9)
    prob.close()


def test_more_populations_than_cached_workspaces_on_one_problem_handle(capi, ctx, orc):
    """Six populations evolved through ONE problem handle (four workspaces are cached per handle), from six threads at once and
    then again one after the other: a workspace is never evicted while a call runs on it, and the results do not depend on which
    workspaces survived."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    prob = capi.Problem(ctx, "ackley", dim=12)
    lb, ub = prob.bounds()
    NP, K = 200, 6
    xs = [np.random.default_rng(40 + k).uniform(lb, ub, (NP, 12)) for k in range(K)]
    fs = [prob.eval_host(x)[:, 0] for x in xs]
    al = np.array([2, 3, 7, 10, 13, 14, 15, 16], dtype=np.uint32)

    def run(k, dx, df):
        for call in range(3):
            capi.check(capi.lib().pgc_de_evolve_device(prob._h, dx, df, NP, 9, 2, 2, 1, 0.8, 0.9, al.ctypes.data_as(C.c_void_p), al.size, 0.0, 0.0,
                                                       None, None, None, 100 + k, 1 + 9 * call, None, None))

    def all_populations(parallel):
        bufs = [(ctx.to_device(xs[k]), ctx.to_device(fs[k])) for k in range(K)]
        if parallel:
            with ThreadPoolExecutor(K) as pool:
                list(pool.map(lambda k: run(k, *bufs[k]), range(K)))
        else:
            for k in range(K):
                run(k, *bufs[k])
        ctx.synchronize()
        out = [(ctx.from_device(dx, xs[0].shape), ctx.from_device(df, fs[0].shape)) for dx, df in bufs]
        for dx, df in bufs:
            ctx.free(dx)
            ctx.free(df)
        return out

    seq = all_populations(False)
    par = all_populations(True)
    for (xa, fa), (xb, fb), f0 in zip(seq, par, fs):
        assert np.array_equal(xa, xb) and np.array_equal(fa, fb) and fa.min() < f0.min()
    prob.close()
