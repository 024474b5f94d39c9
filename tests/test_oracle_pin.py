"""Pins the restated random-draw operators and generation loops (oracle/restate_*.c) to the UNMODIFIED compiled reference.

The reference algorithms draw from one sequential std::mt19937 through libstdc++ distributions and order indices with
std::sort.  oracle/mt19937.h and oracle/restate_std_sort.c restate those library algorithms; with the draw source switched to
them (the *_mt entry points) the restated loops must reproduce the compiled reference BIT FOR BIT: sbx_crossover_impl,
polynomial_mutation_impl, mo_tournament_selection_impl (genetic_operators.cpp:71-211), nsga2::evolve (nsga2.cpp:91-307),
pso_gen::evolve (pso_gen.cpp:120-590), de / sade / de1220 ::evolve, sga::evolve and population(prob, n, seed).
The device is compared (tests/test_gpu_*.py) with the SAME restated statements fed from Philox draws.
"""
import numpy as np
import pytest


# ---------------------------------------------------------------- the library layer itself
@pytest.mark.parametrize("seed", [0, 1, 5489, 123456789])
def test_mt19937_and_distributions_match_libstdcxx(orc, ref, seed):
    assert (orc.mt_sequence(seed, "raw", 2000) == ref.std_sequence(seed, "raw", 2000)).all()
    assert (orc.mt_sequence(seed, "u01", 2000) == ref.std_sequence(seed, "u01", 2000)).all()
    assert (orc.mt_sequence(seed, "normal", 2001) == ref.std_sequence(seed, "normal", 2001)).all()
    assert (orc.mt_sequence(seed, "real", 500, 3, 7) == ref.std_sequence(seed, "real", 500, 3, 7)).all()
    for a, b in ((0, 1), (0, 6), (3, 1023), (0, 65535), (0, 99999), (5, 5), (0, 2 ** 31)):
        assert (orc.mt_sequence(seed, "int", 3000, a, b) == ref.std_sequence(seed, "int", 3000, a, b)).all(), (a, b)
    for t, p in ((30, 0.02), (10, 0.5), (100, 0.05), (7, 0.9)):
        assert (orc.mt_binomial(seed, t, p, 400) == ref.std_binomial(seed, t, p, 400)).all(), (t, p)


def test_mt19937_known_answer(orc):
    # the C++ standard's check value: the 10000th output of a default-seeded (5489) mt19937 is 4123659995
    assert int(orc.mt_sequence(5489, "raw", 10000)[-1]) == 4123659995


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 64, 65, 1000, 65536, 65537, 70000, 131072])
def test_shuffle_matches_std_shuffle(orc, ref, n):
    # n <= 65536: two swap positions per draw; above: one (stl_algo.h:3768)
    assert (orc.mt_shuffles(7, n, 3) == ref.std_shuffles(7, n, 3)).all()


def test_sort_matches_std_sort_tie_order(orc, ref):
    rng = np.random.default_rng(3)
    for n in (1, 2, 15, 16, 17, 18, 33, 100, 1000, 5000, 70000):
        for vals in (3, 10, 1000, 10 ** 9):
            for desc in (False, True):
                k = rng.integers(0, vals, n).astype(float)
                assert (orc.std_argsort(k, desc) == ref.std_argsort(k, desc)).all(), (n, vals, desc)
    for n in (1000, 4096, 20000):
        for k in (np.r_[np.arange(n // 2), np.arange(n // 2)[::-1]].astype(float), (np.arange(n) % 7).astype(float), np.zeros(n)):
            assert (orc.std_argsort(k) == ref.std_argsort(k)).all()

    def killer(n):  # median-of-three killer: drives introsort into its heapsort fallback
        k, a = n // 2, [0] * n
        for i in range(k):
            a[i] = i + 1 if i % 2 == 0 else k + i + (0 if k % 2 else -1) + 1
            a[k + i] = (i + 1) * 2
        return np.array(a, float)
    for n in (64, 1000, 10000, 100000):
        assert (orc.std_argsort(killer(n)) == ref.std_argsort(killer(n))).all()


def test_mo_utilities_with_ties_match_reference_in_libstdcxx_mode(orc, ref):
    """crowding_distance / select_best_N_mo / sort_population_mo on inputs FULL of ties (integer-valued objectives, duplicated
    points): in libstdc++ sort mode the restatement equals the compiled reference exactly (the default stable mode, which the
    device follows, is pinned on tie-free inputs and the reference's own KATs in tests/test_oracle.py)."""
    rng = np.random.default_rng(11)
    try:
        for n, m, vals in ((40, 2, 6), (200, 2, 12), (300, 3, 8), (1000, 2, 40)):
            f = rng.integers(0, vals, (n, m)).astype(float)
            f[n // 2:n // 2 + n // 10] = f[:n // 10]
            orc.set_sort_mode(True)
            assert np.array_equal(orc.crowding_distance(f), ref.crowding_distance(f))
            assert (orc.sort_population_mo(f) == ref.sort_population_mo(f)).all()
            for N in (1, n // 3, n // 2, n - 1):
                assert (orc.select_best_N_mo(f, N) == ref.select_best_N_mo(f, N)).all()
    finally:
        orc.set_sort_mode(False)


# ---------------------------------------------------------------- operators on injected (mt19937) draws
def test_genetic_operators_bit_exact(orc, ref):
    rng = np.random.default_rng(1)
    for t in range(300):
        nx = int(rng.integers(1, 40))
        lb = rng.uniform(-5, 0, nx)
        ub = lb + rng.uniform(0.1, 5, nx)
        p1, p2 = rng.uniform(lb, ub), rng.uniform(lb, ub)
        if t % 3 == 0:
            p2[::2] = p1[::2]          # |p1 - p2| <= 1e-14: the gene gate draw is consumed, nothing else
        if t % 7 == 0:
            p1[0], p2[0] = lb[0], ub[0]  # parents on the bounds
        rank = rng.integers(0, 3, 20)
        cd = rng.choice([0.1, 0.5, np.inf], 20)
        p_cr, eta_c, p_m, eta_m = rng.choice([0.5, 0.95, 1.0]), rng.choice([1., 10., 100.]), rng.choice([0.01, 0.3, 1.0]), rng.choice([1., 50.])
        a = orc.genetic_operators_mt(p1, p2, lb, ub, p_cr, eta_c, p_m, eta_m, rank, cd, t)
        b = ref.genetic_operators(p1, p2, lb, ub, p_cr, eta_c, p_m, eta_m, rank, cd, t)
        for u, v in zip(a, b):
            assert np.array_equal(u, v), t


# ---------------------------------------------------------------- whole evolve() runs
@pytest.mark.parametrize("fam,pid,args,nobj", [("zdt", 1, (1, 30), 2), ("zdt", 4, (4, 10), 2), ("zdt", 6, (6, 10), 2),
                                               ("dtlz", 2, (2, 12, 3, 100), 3), ("dtlz", 1, (1, 7, 3, 100), 3)])
def test_nsga2_evolve_bit_exact(orc, ref, fam, pid, args, nobj):
    rng = np.random.default_rng(pid)
    rp = ref.problem(fam, *args)
    lb, ub = rp.bounds()
    for NP, gens, seed in ((8, 3, 1), (52, 10, 2), (200, 12, 3), (512, 4, 4)):
        x0 = rng.uniform(lb, ub, (NP, lb.size))
        xr, fr = ref.evolve_from(rp, "nsga2", [0.95, 10., 0.01, 50.], x0, gens, seed)
        f0 = np.array([rp.fitness(x) for x in x0])
        xo, fo = orc.nsga2_evolve_mt(fam, pid, nobj, 100, lb, ub, x0, f0, gens, 0.95, 10., 0.01, 50., seed)
        assert np.array_equal(xr, xo) and np.array_equal(fr, fo), (NP, gens)


@pytest.mark.parametrize("fam,dim", [("rastrigin", 10), ("rosenbrock", 7), ("ackley", 5)])
def test_pso_gen_evolve_bit_exact(orc, ref, fam, dim):
    rng = np.random.default_rng(dim)
    rp = ref.problem(fam, dim)
    lb, ub = rp.bounds()
    op = orc.problem(fam, dim=dim)
    for variant in (1, 2, 3, 4, 5, 6):  # 6: the fully informed swarm (one draw per gene and neighbour, :318-326)
        # gbest, lbest rings, von Neumann lattice (24 = 4 x 6; 23 is prime: one row), adaptive random graphs (out-degree 3, 1, 4)
        for nt, npar, n in ((1, 4, 23), (2, 4, 23), (2, 2, 23), (2, 7, 23), (3, 4, 24), (3, 4, 23), (4, 3, 23), (4, 1, 12), (4, 4, 40)):
            gens, seed = 12, variant * 10 + nt
            x0 = rng.uniform(lb, ub, (n, dim))
            xr, fr = ref.evolve_from(rp, "pso_gen", [0.7298, 2.05, 2.05, 0.5, variant, nt, npar], x0, gens, seed)
            f0 = np.array([rp.fitness(x) for x in x0])[:, 0]
            xo, fo = orc.pso_evolve_mt(op, lb, ub, x0, f0, gens=gens, variant=variant, neighb_type=nt, neighb_param=npar, seed=seed)
            assert np.array_equal(xr, xo) and np.array_equal(fr[:, 0], fo), (variant, nt, npar)


@pytest.mark.parametrize("fam,dim", [("rastrigin", 10), ("rosenbrock", 6), ("schwefel", 4)])
def test_de_family_evolve_bit_exact(orc, ref, fam, dim):
    rng = np.random.default_rng(dim)
    rp = ref.problem(fam, dim)
    lb, ub = rp.bounds()
    op = orc.problem(fam, dim=dim)
    NP, gens = 20, 15
    x0 = rng.uniform(lb, ub, (NP, dim))
    f0 = np.array([rp.fitness(x) for x in x0])[:, 0]
    for variant in range(1, 11):          # de.cpp:154-275
        xr, fr = ref.evolve_from(rp, "de", [0.8, 0.9, variant, 1e-6, 1e-6], x0, gens, variant)
        xo, fo, _ = orc.de_evolve_mt(op, lb, ub, x0, f0, gens=gens, algo="de", variant=variant, F=0.8, CR=0.9, seed=variant)
        assert np.array_equal(xr, xo) and np.array_equal(fr[:, 0], fo), ("de", variant)
    for variant in range(1, 19):          # sade.cpp:188-494, jDE (1) and iDE (2) self-adaptation
        for adptv in (1, 2):
            seed = variant * 3 + adptv
            xr, fr = ref.evolve_from(rp, "sade", [variant, adptv, 1e-6, 1e-6], x0, gens, seed)
            xo, fo, _ = orc.de_evolve_mt(op, lb, ub, x0, f0, gens=gens, algo="sade", variant=variant, variant_adptv=adptv, seed=seed)
            assert np.array_equal(xr, xo) and np.array_equal(fr[:, 0], fo), ("sade", variant, adptv)
    for allowed in ((2, 3, 7, 10, 13, 14, 15, 16), tuple(range(1, 19)), (5,)):   # de1220.cpp:80-600
        for adptv in (1, 2):
            seed = len(allowed) + adptv
            xr, fr = ref.evolve_from(rp, "de1220", [adptv, 1e-6, 1e-6, len(allowed), *allowed], x0, gens, seed)
            xo, fo, _ = orc.de_evolve_mt(op, lb, ub, x0, f0, gens=gens, algo="de1220", variant_adptv=adptv, allowed=allowed, seed=seed)
            assert np.array_equal(xr, xo) and np.array_equal(fr[:, 0], fo), ("de1220", allowed, adptv)


def test_de_generational_equals_sequential_where_the_reference_allows(orc):
    """de, and sade / de1220 with jDE adaptation (variant_adptv = 1), read only the previous generation while building trials,
    so the generational loop the device runs and the reference's one-at-a-time loop are the same function of the draws: checked
    here on the SAME (Philox) draws by running the restatement both ways."""
    import ctypes as C
    rng = np.random.default_rng(5)
    op = orc.problem("rastrigin", dim=8)
    lb, ub = np.full(8, -5.12), np.full(8, 5.12)
    x0 = rng.uniform(lb, ub, (16, 8))
    f0 = orc.simple("rastrigin", x0)
    for algo, variant, adptv in (("de", 2, 1), ("de", 7, 1), ("sade", 2, 1), ("sade", 16, 1), ("de1220", 2, 1)):
        a = orc.de_evolve(op, lb, ub, x0, f0, gens=10, algo=algo, variant=variant, variant_adptv=adptv, seed=9)
        b = orc.de_evolve(op, lb, ub, x0, f0, gens=10, algo=algo, variant=variant, variant_adptv=adptv, seed=9, sequential=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (algo, variant)


@pytest.mark.parametrize("fam,dim", [("rastrigin", 10), ("rosenbrock", 6)])
def test_sga_evolve_bit_exact(orc, ref, fam, dim):
    rng = np.random.default_rng(dim)
    rp = ref.problem(fam, dim)
    lb, ub = rp.bounds()
    op = orc.problem(fam, dim=dim)
    NP, gens = 24, 8
    x0 = rng.uniform(lb, ub, (NP, dim))
    f0 = np.array([rp.fitness(x) for x in x0])[:, 0]
    k = 0
    for xo_ in ("exponential", "binomial", "single", "sbx"):
        for mu in ("gaussian", "uniform", "polynomial"):
            for se, ps in (("tournament", 2), ("tournament", 5), ("truncated", 3)):
                k += 1
                m, pm = 0.1, (0.05 if mu == "gaussian" else 1.0)
                xr, fr = ref.evolve_from(rp, "sga", [0.9, 1.0, m, pm, ps], x0, gens, k, strategies=f"{xo_},{mu},{se}")
                xo, fo = orc.sga_evolve_mt(op, lb, ub, x0, f0, gens=gens, cr=0.9, eta_c=1.0, m=m, param_m=pm, param_s=ps, crossover=xo_,
                                           mutation=mu, selection=se, seed=k)
                assert np.array_equal(xr, xo) and np.array_equal(fr[:, 0], fo), (xo_, mu, se, ps)


def test_population_constructor_bit_exact(orc, ref):
    for fam, args in (("rastrigin", (10,)), ("zdt", (1, 30)), ("lennard_jones", (5,)), ("cec2014", (3, 10))):
        rp = ref.problem(fam, *args)
        lb, ub = rp.bounds()
        for n, seed in ((1, 0), (7, 3), (100, 42)):
            xr, ir = ref.population_init(rp, n, seed)
            xo, io = orc.population_init_mt(lb, ub, n, seed)
            assert np.array_equal(xr, xo) and np.array_equal(ir, io), (fam, n)


@pytest.mark.parametrize("fam,args,n", [("zdt", (1, 8), 24), ("zdt", (3, 6), 15), ("dtlz", (2, 7, 3, 100), 20), ("dtlz", (1, 6, 3, 100), 12)])
def test_nspso_evolve_bit_exact(orc, ref, fam, args, n):
    """nspso is generational in the reference itself (nspso.cpp:293-395): the restatement on the mt19937 stream must reproduce
    nspso::evolve bit for bit - leaders (FNDS order, sort_population_mo, niche counts, max-min keys and the tie order std::sort
    leaves), the rejection loop of the leader draw, the move, the archive of the best N of 2N."""
    rp = ref.problem(fam, *args)
    lb, ub = rp.bounds()
    if fam == "zdt":
        op = orc.problem("zdt", prob_id=args[0], dim=args[1])
    else:
        op = orc.problem("dtlz", prob_id=args[0], dim=args[1], nobj=args[2], param=args[3])
    rng = np.random.default_rng(n)
    x0 = rng.uniform(lb, ub, (n, len(lb)))
    x0[3] = x0[1]  # duplicates: ties in every sort
    f0 = np.array([rp.fitness(x) for x in x0])
    for diversity in ("crowding distance", "niche count", "max min"):
        for lsr, gens in ((60, 9), (5, 4), (100, 4)):
            seed = lsr + gens
            xr, fr = ref.evolve_from(rp, "nspso", [0.6, 2.0, 2.0, 1.0, 0.5, lsr], x0, gens, seed, strategies=diversity)
            xo, fo, *_ = orc.nspso_evolve(op, lb, ub, x0, f0, gens=gens, leader_selection_range=lsr, diversity=diversity, seed=seed, mt=True)
            assert np.array_equal(xr, xo) and np.array_equal(fr, fo), (diversity, lsr, gens)


@pytest.mark.parametrize("fam,dim", [("rastrigin", 8), ("rosenbrock", 5), ("ackley", 6), ("griewank", 4)])
def test_gaco_evolve_bit_exact(orc, ref, fam, dim):
    """gaco::evolve (gaco.cpp:104-445) restated on the mt19937 stream: penalties against the oracle parameter, the archive update with
    its accuracy filter, the kernel weights and their threshold switch, sigma from the archive's spread and the generation mark,
    the ants (kernel choice, normal deviates with the ten redraws and the clamp), the evaluation / improvement counters and the oracle
    update must reproduce the compiled reference bit for bit."""
    rp = ref.problem(fam, dim)
    lb, ub = rp.bounds()
    op = orc.problem(fam, dim=dim)
    rng = np.random.default_rng(dim)
    # ker, q, oracle, acc, threshold, n_gen_mark, impstop, evalstop, focus
    cases = [(20, (13, 1.0, 0.0, 0.01, 1, 7, 100000, 100000, 0.0), 12), (30, (30, 1.0, 1e9, 0.0, 5, 3, 100000, 100000, 0.0), 15),
             (16, (5, 0.5, 0.0, 0.5, 3, 7, 100000, 100000, 4.0), 10), (25, (8, 1.0, 50.0, 0.01, 2, 2, 3, 100000, 0.0), 20),
             (25, (8, 1.0, 0.0, 0.01, 1, 7, 100000, 4, 0.0), 30), (12, (2, 2.0, -5.0, 0.01, 4, 1, 100000, 100000, 100.0), 9)]
    for n, par, gens in cases:
        seed = n + gens
        x0 = rng.uniform(lb, ub, (n, dim))
        x0[3] = x0[1]  # duplicates: equal penalties in both sorts
        f0 = np.array([rp.fitness(x) for x in x0])[:, 0]
        xr, fr = ref.evolve_from(rp, "gaco", list(par), x0, gens, seed)
        ker, q, oracle, acc, threshold, n_gen_mark, impstop, evalstop, focus = par
        xo, fo, _, _ = orc.gaco_evolve(op, lb, ub, x0, f0, gens=gens, ker=ker, q=q, oracle=oracle, acc=acc, threshold=threshold,
                                       n_gen_mark=n_gen_mark, impstop=impstop, evalstop=evalstop, focus=focus, seed=seed, mt=True)
        assert np.array_equal(xr, xo) and np.array_equal(fr[:, 0], fo), (n, par, gens)
    # memory = true: ONE algorithm object evolved several times (gaco.cpp:106-108, :223-250, :732-752, :778-784): the archive and the call
    # counter carry over, the weights follow the counter, the archive is never written back into the population
    for n, par, gens, calls in ((20, (8, 1.0, 1e9, 0.01, 3, 7, 100000, 100000, 0.0), 1, 6), (24, (6, 1.0, 0.0, 0.01, 2, 3, 100000, 100000, 5.0), 3, 4)):
        seed = n + calls
        x0 = rng.uniform(lb, ub, (n, dim))
        f0 = np.array([rp.fitness(x) for x in x0])[:, 0]
        xr, fr = ref.evolve_from(rp, "gaco", list(par) + [1, calls], x0, gens, seed)
        ker, q, oracle, acc, threshold, n_gen_mark, impstop, evalstop, focus = par
        xo, fo, _, _ = orc.gaco_evolve(op, lb, ub, x0, f0, gens=gens, ker=ker, q=q, oracle=oracle, acc=acc, threshold=threshold,
                                       n_gen_mark=n_gen_mark, impstop=impstop, evalstop=evalstop, focus=focus, seed=seed, mt=True, memory=True,
                                       calls=calls)
        assert np.array_equal(xr, xo) and np.array_equal(fr[:, 0], fo), ("memory", n, par, gens, calls)


@pytest.mark.parametrize("fam,args", [("zdt", (1, 8)), ("zdt", (2, 6)), ("zdt", (3, 7)), ("dtlz", (2, 7, 3, 100)), ("dtlz", (1, 6, 3, 100))])
def test_maco_evolve_bit_exact(orc, ref, fam, args):
    """maco::evolve (maco.cpp:88-533) restated on the mt19937 stream: the archive rebuilt from the fronts of (archive + population) in
    order of decreasing hypervolume contribution, the forced extremities of an overflowing first front, the ideal-point counters, gaco's
    pheromone values and ants.  The restated contributions agree with hv2d / HyCon3D to ~1e-16 of the hypervolume; it is the ORDER they
    induce that the algorithm consumes, so the runs must coincide bit for bit."""
    rp = ref.problem(fam, *args)
    lb, ub = rp.bounds()
    if fam == "zdt":
        op = orc.problem("zdt", prob_id=args[0], dim=args[1])
    else:
        op = orc.problem("dtlz", prob_id=args[0], dim=args[1], nobj=args[2], param=args[3])
    rng = np.random.default_rng(len(lb))
    # n, (ker, q, threshold, n_gen_mark, evalstop, focus), gens
    for n, par, gens in ((24, (10, 1.0, 1, 7, 100000, 0.0), 8), (30, (30, 1.0, 3, 3, 100000, 0.0), 7), (20, (4, 0.5, 2, 7, 100000, 5.0), 9),
                         (28, (12, 1.0, 1, 2, 3, 0.0), 15)):
        seed = n + gens
        x0 = rng.uniform(lb, ub, (n, len(lb)))
        f0 = np.array([rp.fitness(x) for x in x0])
        xr, fr = ref.evolve_from(rp, "maco", list(par), x0, gens, seed)
        ker, q, threshold, n_gen_mark, evalstop, focus = par
        xo, fo, _, _ = orc.maco_evolve(op, lb, ub, x0, f0, gens=gens, ker=ker, q=q, threshold=threshold, n_gen_mark=n_gen_mark,
                                       evalstop=evalstop, focus=focus, seed=seed, mt=True)
        assert np.array_equal(xr, xo) and np.array_equal(fr, fo), (n, par, gens)


@pytest.mark.parametrize("fam,args,NP,wgen", [("zdt", (1, 8), 24, "grid"), ("zdt", (2, 6), 30, "low discrepancy"), ("zdt", (3, 7), 20, "random"),
                                              ("dtlz", (2, 7, 3, 100), 21, "grid"), ("dtlz", (1, 6, 3, 100), 28, "low discrepancy")])
def test_moead_gen_evolve_bit_exact(orc, ref, fam, args, NP, wgen):
    """moead_gen::evolve (the reference's own generational MOEA/D) restated on the mt19937 stream: candidate construction (diversity
    draw, parent rejection loop, DE operator with bound repair, polynomial mutation), batch evaluation, the sequential insertion
    with its per-individual std::shuffle and the `limit` cut - bit for bit, for the three decompositions, with and without
    diversity preservation.  Weights and neighbourhoods come from the compiled reference's own utilities."""
    rp = ref.problem(fam, *args)
    lb, ub = rp.bounds()
    op = orc.problem("zdt", prob_id=args[0], dim=args[1]) if fam == "zdt" else orc.problem("dtlz", prob_id=args[0], dim=args[1], nobj=args[2],
                                                                                             param=args[3])
    m = rp.nf
    x0 = np.random.default_rng(NP).uniform(lb, ub, (NP, len(lb)))
    f0 = np.array([rp.fitness(x) for x in x0])
    for decomposition in ("tchebycheff", "weighted", "bi"):
        for T, CR, F, realb, limit, preserve, gens in ((5, 1.0, 0.5, 0.9, 2, True, 6), (8, 0.6, 0.8, 0.5, 1, True, 4), (4, 0.9, 0.5, 0.9, 2, False, 4)):
            seed = T * 7 + gens
            w = ref.decomposition_weights(m, NP, wgen, seed)
            nb = ref.knn(w, T)
            xr, fr = ref.evolve_from(rp, "moead_gen", [T, CR, F, 20.0, realb, limit, 1.0 if preserve else 0.0], x0, gens, seed,
                                     strategies=f"{wgen},{decomposition}")
            burn = (NP - m) * (m - 1) if wgen == "random" else 0
            xo, fo = orc.moead_gen_evolve(op, lb, ub, x0, f0, w, nb, gens=gens, decomposition=decomposition, CR=CR, F=F, eta_m=20.0, realb=realb,
                                          limit=limit, preserve_diversity=preserve, seed=seed, mt=True, burn_draws=burn)
            assert np.array_equal(xr, xo) and np.array_equal(fr, fo), (decomposition, T, CR, realb, limit, preserve)


@pytest.mark.parametrize("param,NP", [(3, 24), (11, 32), (5, 20)])
def test_nsga2_on_zdt5_integer_alleles_bit_exact(orc, ref, param, NP):
    """ZDT5 is all-integer (nix == nx): sbx_crossover_impl's two-point crossover of the integer part and polynomial_mutation_impl's
    uniform integer redraw (genetic_operators.cpp:125-137, :187-195), inside whole nsga2::evolve runs on the mt19937 stream."""
    rp = ref.problem("zdt", 5, param)
    lb, ub = rp.bounds()
    nx = len(lb)
    x0 = np.floor(np.random.default_rng(param).uniform(lb, ub + 1, (NP, nx))).clip(lb, ub)
    f0 = np.array([rp.fitness(x) for x in x0])
    orc.set_nix(nx)
    try:
        for cr, m, gens in ((0.95, 0.01, 8), (0.5, 0.2, 5), (0.9, 1.0 / nx, 6)):
            seed = gens + param
            xr, fr = ref.evolve_from(rp, "nsga2", [cr, 10.0, m, 50.0], x0, gens, seed)
            xo, fo = orc.nsga2_evolve_mt("zdt", 5, 2, 0, lb, ub, x0, f0, gens, cr, 10.0, m, 50.0, seed)
            assert np.array_equal(xr, xo) and np.array_equal(fr, fo), (cr, m, gens)
            assert np.array_equal(xo, np.round(xo))
    finally:
        orc.set_nix(0)
