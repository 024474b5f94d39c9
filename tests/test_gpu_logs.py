"""The log lines of the reference UDAs (`set_verbosity(v)` / `get_log()`) from `pgc_algo_evolve_logged_device`: every line is
recomputed here, with the reference's formula, from the population an evolve() of exactly that many generations returns (the draws
are keyed by generation, so a shorter run is a prefix of a longer one), and logging must not change the trajectory."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from pagmo2_b200 import capi as m
    return m


def _due(gens, v):
    return [g for g in range(1, gens + 1) if g % v == 1 or v == 1]  # de.cpp:327


def _dx_df(x, f):
    """de.cpp:328-336: best = first minimum, worst = first maximum; dx = sum |x_worst - x_best|, df = |f_worst - f_best|."""
    b, w = int(np.argmin(f)), int(np.argmax(f))
    return f[b], float(np.sum(np.abs(x[w] - x[b]))), abs(f[w] - f[b])


@pytest.mark.parametrize("name,kw,cols", [("de", dict(variant=2), 5), ("sade", dict(variant=2, variant_adptv=1), 7),
                                          ("de1220", dict(variant_adptv=1), 8), ("sade", dict(variant=7, variant_adptv=2), 7)])
@pytest.mark.parametrize("verbosity", (1, 4))
def test_de_family_log(capi, ctx, name, kw, cols, verbosity):
    prob = capi.Problem(ctx, "rastrigin", dim=8)
    lb, ub = prob.bounds()
    n, gens = 40, 10
    x = np.random.default_rng(2).uniform(lb, ub, (n, prob.nx))
    f = prob.eval_host(x)
    algo = capi.algo_desc(name, gens=gens, seed=7, ftol=0.0, xtol=0.0, **kw)
    xl, fl, done, log = prob.evolve_logged(algo, x, f, verbosity)
    x0, f0, _ = prob.evolve(algo, x, f)
    assert done == gens and np.array_equal(xl, x0) and np.array_equal(fl, f0)  # logging does not change the run
    due = _due(gens, verbosity)
    assert log.shape == (len(due), cols) and log[:, 0].tolist() == due and log[:, 1].tolist() == [g * n for g in due]
    for row in log:
        xg, fg, _ = prob.evolve(capi.algo_desc(name, gens=int(row[0]), seed=7, ftol=0.0, xtol=0.0, **kw), x, f)
        best, dx, df = _dx_df(xg, fg[:, 0])
        assert row[2] == best and np.isclose(row[-2], dx, rtol=1e-13) and np.isclose(row[-1], df, rtol=1e-13)
        if name != "de":  # F and CR of the individual that last improved the global best
            assert 0.0 < row[3] <= 1.0 + 1e-9 or kw["variant_adptv"] == 2
            assert np.isfinite(row[4])
        if name == "de1220":
            assert row[5] in (2, 3, 7, 10, 13, 14, 15, 16)
    prob.close()


def test_de_log_stops_where_the_exit_test_fires(capi, ctx):
    """de.cpp:308-321: the generation whose xtol / ftol test fires returns before its log line."""
    prob = capi.Problem(ctx, "rosenbrock", dim=3)
    lb, ub = prob.bounds()
    x = np.random.default_rng(5).uniform(lb, ub, (20, 3))
    f = prob.eval_host(x)
    algo = capi.algo_desc("de", gens=2000, seed=3, ftol=1e-3, xtol=1e-3)
    _, _, done, log = prob.evolve_logged(algo, x, f, 1)
    assert 0 < done < 2000 and log.shape[0] == done - 1 and log[-1, 0] == done - 1
    prob.close()


@pytest.mark.parametrize("variant,ntype", [(5, 2), (1, 1), (6, 3)])
def test_pso_gen_log(capi, ctx, variant, ntype):
    """pso_gen.cpp:464-518, as written: gbest = min lbfit; the running "mean velocity" m <- (m + sum_j |V_ij / width_j|) / dim over the
    particles; mean lbest; mean pairwise distance of the CURRENT positions in units of the box."""
    prob = capi.Problem(ctx, "ackley", dim=6)
    lb, ub = prob.bounds()
    n, gens, v = 30, 7, 3
    x = np.random.default_rng(9).uniform(lb, ub, (n, prob.nx))
    f = prob.eval_host(x)
    algo = capi.algo_desc("pso_gen", gens=gens, seed=11, variant=variant, neighb_type=ntype)
    xl, fl, _, log = prob.evolve_logged(algo, x, f, v)
    x0, f0, _ = prob.evolve(algo, x, f)
    assert np.array_equal(xl, x0) and np.array_equal(fl, f0)
    assert log[:, 0].tolist() == _due(gens, v) and log[:, 1].tolist() == [g * n for g in _due(gens, v)] and log.shape[1] == 6
    for row in log:
        g = int(row[0])
        v0 = np.zeros_like(x)
        # the velocities of a memory-less run: drawn inside; read them back by running with memory from the same start
        xa, fa, _, st = prob.evolve_memory(capi.algo_desc("pso_gen", gens=g, seed=11, variant=variant, neighb_type=ntype), x, f)
        _, _, _, xcur = prob.pso_evolve(x, f[:, 0], gens=g, seed=11, variant=variant, neighb_type=ntype)
        V, lbfit = st["a"], fa[:, 0]
        m = 0.0
        for i in range(n):
            m = (m + np.sum(np.abs(V[i] / (ub - lb)))) / prob.nx
        d = 0.0
        for i in range(n):
            for j in range(i + 1, n):
                d += np.sqrt(np.sum((xcur[i] - xcur[j]) ** 2 / (ub - lb) / (ub - lb)))
        d /= (n - 1) * n / 2
        assert row[2] == lbfit.min() and np.isclose(row[3], m, rtol=1e-12) and np.isclose(row[4], lbfit.sum() / n, rtol=1e-13)
        assert np.isclose(row[5], d, rtol=1e-12)
    prob.close()


@pytest.mark.parametrize("name", ("nsga2", "nspso"))
def test_mo_log_is_the_ideal_point_before_the_generation(capi, ctx, name):
    """nsga2.cpp:144-173 / nspso.cpp:163-192: logged at the top of the loop, so line `gen` describes the population (nspso: the
    archive, which a memory-less run starts as the population) after gen - 1 generations, with (gen - 1) * NP evaluations."""
    prob = capi.Problem(ctx, "dtlz", prob_id=2, dim=7, nobj=3, param=100)
    lb, ub = prob.bounds()
    n, gens, v = 36, 9, 4
    x = np.random.default_rng(4).uniform(lb, ub, (n, prob.nx))
    f = prob.eval_host(x)
    algo = capi.algo_desc(name, gens=gens, seed=13)
    xl, fl, _, log = prob.evolve_logged(algo, x, f, v)
    x0, f0, _ = prob.evolve(algo, x, f)
    assert np.array_equal(xl, x0) and np.array_equal(fl, f0)
    assert log[:, 0].tolist() == _due(gens, v) and log[:, 1].tolist() == [(g - 1) * n for g in _due(gens, v)] and log.shape[1] == 5
    for row in log:
        g = int(row[0]) - 1
        if name == "nsga2":
            fg = f if g == 0 else prob.evolve(capi.algo_desc(name, gens=g, seed=13), x, f)[1]
        else:
            fg = f if g == 0 else prob.evolve_memory(capi.algo_desc(name, gens=g, seed=13), x, f)[3]["c"]
        assert np.array_equal(row[2:], fg.min(axis=0))
    prob.close()


@pytest.mark.parametrize("verbosity", (1, 3))
def test_sga_log(capi, ctx, verbosity):
    """sga.cpp:252-274: (gen, fevals, best of the parents, improvement = that minus the best child); verbosity 1 logs only the
    generations whose children improve on the parents, larger values every `verbosity` generations whatever the improvement."""
    prob = capi.Problem(ctx, "rastrigin", dim=6)
    lb, ub = prob.bounds()
    n, gens = 30, 12
    x = np.random.default_rng(6).uniform(lb, ub, (n, prob.nx))
    f = prob.eval_host(x)
    kw = dict(seed=21, crossover=capi.SGA_CROSSOVER["sbx"], mutation=capi.SGA_MUTATION["gaussian"])
    algo = capi.algo_desc("sga", gens=gens, **kw)
    xl, fl, _, log = prob.evolve_logged(algo, x, f, verbosity)
    x0, f0, _ = prob.evolve(algo, x, f)
    assert np.array_equal(xl, x0) and np.array_equal(fl, f0) and log.shape[1] == 4
    best = [f.min()] + [prob.evolve(capi.algo_desc("sga", gens=g, **kw), x, f)[1].min() for g in range(1, gens + 1)]
    rows = {int(r[0]): r for r in log}
    for g in range(1, gens + 1):
        improved = best[g] < best[g - 1]
        if verbosity == 1:
            assert (g in rows) == improved, g
        else:
            assert (g in rows) == (g % verbosity == 1), g
        if g in rows:
            r = rows[g]
            assert r[1] == g * n and r[2] == best[g - 1]
            assert (r[3] > 0) == improved and (not improved or np.isclose(r[2] - r[3], best[g], rtol=1e-12))
    prob.close()


def test_cmaes_log(capi, ctx):
    """cmaes.cpp:276-296: (gen, fevals, best, dx, df, sigma) logged after the sampling and the exit tests of a generation, i.e. on the
    population and the step size the previous generations left."""
    prob = capi.Problem(ctx, "rosenbrock", dim=5)
    lb, ub = prob.bounds()
    n, gens, v = 12, 9, 4
    x = np.random.default_rng(3).uniform(lb, ub, (n, prob.nx))
    f = prob.eval_host(x)
    algo = capi.algo_desc("cmaes", gens=gens, seed=31, ftol=0.0, xtol=0.0)
    xl, fl, _, log = prob.evolve_logged(algo, x, f, v)
    x0, f0, _ = prob.evolve(algo, x, f)
    assert np.array_equal(xl, x0) and np.array_equal(fl, f0)
    assert log[:, 0].tolist() == _due(gens, v) and log[:, 1].tolist() == [(g - 1) * n for g in _due(gens, v)] and log.shape[1] == 6
    for row in log:
        g = int(row[0]) - 1
        if g == 0:
            fg, sigma = f[:, 0], 0.5
        else:
            _, fg, _, sigma = prob.cmaes_evolve(x, f[:, 0], gens=g, seed=31, ftol=0.0, xtol=0.0)
        assert row[2] == fg.min() and np.isclose(row[4], fg.max() - fg.min(), rtol=1e-13) and row[5] == sigma and row[3] > 0
    prob.close()


def test_gaco_maco_moead_logs(capi, ctx):
    """the algorithms with their own entry points log through pgc_log_capture_begin / _end."""
    # gaco (gaco.cpp:254-287, :405-445): in-loop lines for every due generation but the last, from the archive; a final line from the champion
    prob = capi.Problem(ctx, "rastrigin", dim=6)
    lb, ub = prob.bounds()
    n, gens, ker = 30, 7, 8
    x = np.random.default_rng(1).uniform(lb, ub, (n, 6))
    f = prob.eval_host(x)[:, 0]
    with capi.log_capture(ctx, 3, 8, 7) as cap:
        xl, fl, st, _ = prob.gaco_evolve(x, f, gens=gens, ker=ker, oracle=1e9, seed=5)
    x0, f0, _, _ = prob.gaco_evolve(x, f, gens=gens, ker=ker, oracle=1e9, seed=5)
    assert np.array_equal(xl, x0) and np.array_equal(fl, f0)
    log = cap.rows
    assert log[:, 0].tolist() == [1, 4, 7] and log[:, 1].tolist() == [0, 3 * n, 7 * n] and (log[:, 3] == ker).all()
    assert log[0, 2] == f.min() and log[0, 4] == 1e9            # generation 1: the archive is the best of the start, the oracle untouched
    assert log[2, 2] == min(f.min(), log[2, 2]) and log[2, 4] == st.oracle and (log[:, 5] > 0).all() and (log[:, 6] >= 0).all()
    xa, fa, _, _ = prob.gaco_evolve(x, f, gens=3, ker=ker, oracle=1e9, seed=5)   # the archive after 3 generations = rows 0..ker of the result
    assert log[1, 2] <= fa[:ker, 0].min() + 1e-12
    prob.close()
    # maco (maco.cpp:415-463) and moead_gen (moead_gen.cpp:180-211)
    mo = capi.Problem(ctx, "zdt", prob_id=1, dim=8)
    lb, ub = mo.bounds()
    n = 24
    x = np.random.default_rng(2).uniform(lb, ub, (n, 8))
    f = mo.eval_host(x)
    with capi.log_capture(ctx, 2, 8, 4) as cap:
        xl, fl, _, _ = mo.maco_evolve(x, f, gens=5, ker=10, seed=3)
    assert np.array_equal(xl, mo.maco_evolve(x, f, gens=5, ker=10, seed=3)[0])
    log = cap.rows
    assert log[:, 0].tolist() == [1, 3, 5] and log[:, 1].tolist() == [0, 2 * n, 4 * n]
    assert np.array_equal(log[0, 2:], f.min(axis=0))  # the first archive holds the whole first front: its ideal point is the population's
    assert (log[1:, 2:] <= log[0, 2:] + 1e-15).all()
    w = np.array([[i / (n - 1), 1 - i / (n - 1)] for i in range(n)])
    nb = np.argsort(((w[:, None, :] - w[None, :, :]) ** 2).sum(axis=2), axis=1)[:, 1:6].astype(np.uint32)
    with capi.log_capture(ctx, 2, 8, 5) as cap:
        xl, fl = mo.moead_gen_evolve(x, f, w, nb, gens=4, seed=3)
    assert np.array_equal(xl, mo.moead_gen_evolve(x, f, w, nb, gens=4, seed=3)[0])
    log = cap.rows
    assert log[:, 0].tolist() == [1, 3] and log[:, 1].tolist() == [0, 2 * n]
    ideal = f.min(axis=0)
    adf = sum(max(w[i, k] * abs(f[i, k] - ideal[k]) if w[i, k] != 0 else 1e-4 * abs(f[i, k] - ideal[k]) for k in range(2)) for i in range(n))
    assert np.array_equal(log[0, 3:], ideal) and np.isclose(log[0, 2], adf, rtol=1e-12)  # tchebycheff, multi_objective.cpp:601-611
    assert (log[1, 3:] <= ideal).all()
    mo.close()
    with pytest.raises(capi.PgcError):  # one capture per thread
        with capi.log_capture(ctx, 1, 4, 7):
            with capi.log_capture(ctx, 1, 4, 7):
                pass


def test_log_argument_checks(capi, ctx):
    prob = capi.Problem(ctx, "rastrigin", dim=4)
    x = np.random.default_rng(1).uniform(-5, 5, (16, 4))
    f = prob.eval_host(x)
    with pytest.raises(capi.PgcError):  # verbosity > 0 needs somewhere to put the lines
        capi.check(capi.lib().pgc_algo_evolve_logged_device(prob._h, capi.C.byref(capi.algo_desc("de", gens=2, seed=1)), None, None, 0, 1, None, None,
                                                            1, None, 0, None, None))
    # verbosity 0: a plain evolve, no lines
    xl, fl, _, log = prob.evolve_logged(capi.algo_desc("de", gens=3, seed=1), x, f, 0)
    assert log.shape[0] == 0 and np.array_equal(xl, prob.evolve(capi.algo_desc("de", gens=3, seed=1), x, f)[0])
    prob.close()
