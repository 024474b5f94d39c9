"""GPU parity: Lennard-Jones batch energy and the generational PSO (pso_gen) against the oracle with injected Philox draws."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REL_TOL = 1e-12


@pytest.fixture(scope="module")
def capi():
    from pagmo2_b200 import capi as m
    return m


@pytest.mark.parametrize("atoms", [3, 4, 7, 33, 38, 150, 256, 300])
def test_lennard_jones_parity(capi, ctx, orc, atoms):
    rng = np.random.default_rng(atoms)
    prob = capi.Problem(ctx, "lennard_jones", dim=atoms)
    assert prob.nx == 3 * atoms - 6 and prob.name == f"Lennard Jones Cluster ({atoms} atoms)"
    lb, ub = prob.bounds()
    xs = rng.uniform(lb, ub, (515, prob.nx))
    got, want = prob.eval_host(xs)[:, 0], orc.lennard_jones(atoms, xs)
    assert np.max(np.abs(got - want) / np.abs(want)) <= REL_TOL
    if atoms == 3:  # reference tests/lennard_jones.cpp:59-60
        f = prob.eval_host(np.array([[1.12, -0.33, 2.34], [1.23, -1.23, 0.33]]))[:, 0]
        assert np.allclose(f, [-1.7633355813175688, -1.833100934753864], rtol=1e-13)
        lb3, ub3 = prob.bounds()
        assert list(lb3) == [-3, -3, -3] and list(ub3) == [3, 3, 3]
    if atoms >= 4:  # coincident atoms: the reference assigns DBL_MAX, keeps summing, multiplies by 4 -> +inf (:81-91)
        x = xs[0].copy()
        x[3:6] = [0.0, x[1], x[2]]  # atom 3 on top of atom 2
        assert prob.eval_host(x[None, :])[0, 0] == orc.lennard_jones(atoms, x[None, :])[0] == np.inf
    with pytest.raises(capi.PgcError):
        capi.Problem(ctx, "lennard_jones", dim=2)
    prob.close()


@pytest.mark.parametrize("variant,ntype", [(5, 2), (5, 1), (1, 2), (2, 1), (3, 2), (4, 1), (5, 3), (1, 3), (5, 4), (3, 4), (2, 4), (6, 1), (6, 2), (6, 3), (6, 4)])
def test_pso_matches_oracle(capi, ctx, orc, variant, ntype):
    """Same Philox draws => same trajectories.  Positions are compared with a tolerance (fitness values differ by ulps,
    which can only matter through the <= comparisons of the memory update)."""
    rng = np.random.default_rng(variant * 10 + ntype)
    n, dim = (60 if ntype == 3 else 64), 10   # 60 = a 6 x 10 von Neumann lattice (rows = largest divisor <= sqrt(n))
    prob = capi.Problem(ctx, "rastrigin", dim=dim)
    op = orc.problem("rastrigin", dim=dim)
    lb, ub = prob.bounds()
    x = rng.uniform(lb, ub, (n, dim))
    f = orc.simple("rastrigin", x)
    # topology 4 (adaptive random): 12 generations so that the graph is both kept (best improved) and re-drawn (it did not)
    kw = dict(gens=12 if ntype == 4 else 8, variant=variant, neighb_type=ntype, seed=42, first_generation=1)
    xo, fo, _, co = orc.pso_evolve(op, lb, ub, x, f, **kw)
    xg, fg, _, cg = prob.pso_evolve(x, f, **kw)
    assert np.allclose(cg, co, rtol=1e-9, atol=1e-12) and np.allclose(xg, xo, rtol=1e-9, atol=1e-12)
    assert np.allclose(fg, fo, rtol=1e-9)
    assert (fg <= f + 1e-12).all() and fg.min() < f.min()  # memories never get worse; the swarm improves
    assert (xg >= lb).all() and (xg <= ub).all()
    prob.close()


def test_pso_on_lennard_jones_and_argument_checks(capi, ctx, orc):
    rng = np.random.default_rng(3)
    atoms, n = 13, 256
    prob = capi.Problem(ctx, "lennard_jones", dim=atoms)
    lb, ub = prob.bounds()
    x = rng.uniform(lb, ub, (n, prob.nx))
    f = prob.eval_host(x)[:, 0]
    xg, fg, _, _ = prob.pso_evolve(x, f, gens=40, seed=1)
    assert np.isfinite(fg).all() and fg.min() < f.min() and np.allclose(prob.eval_host(xg)[:, 0], fg, rtol=1e-12)
    for bad in (dict(omega=1.5), dict(eta1=5.0), dict(max_vel=0.0), dict(variant=7), dict(neighb_type=5), dict(variant=0), dict(neighb_param=0)):
        with pytest.raises(capi.PgcError):
            prob.pso_evolve(x, f, gens=1, **bad)
    mo = capi.Problem(ctx, "zdt", prob_id=1, dim=5)
    with pytest.raises(capi.PgcError):
        mo.pso_evolve(np.zeros((8, 5)), np.zeros(8), gens=1)


@pytest.mark.parametrize("variant", (1, 2, 3, 4, 5))
def test_sharded_swarm_one_shard_equals_evolve(capi, ctx, variant):
    """pagmo2_b200.swarm with a single shard (its halos wrap onto itself) must move exactly like pgc_pso_evolve_device; with two
    shards in one process (halos swapped by hand) too - the multi-process version only replaces the swap by an all_gather."""
    from pagmo2_b200.swarm import DeviceShard, ShardedSwarm, halo_rows
    rng = np.random.default_rng(variant)
    prob = capi.Problem(ctx, "rastrigin", dim=7)
    lb, ub = prob.bounds()
    n = 96
    x = rng.uniform(lb, ub, (n, 7))
    f = prob.eval_host(x)[:, 0]
    want_x, want_f, _, _ = prob.pso_evolve(x, f, gens=5, variant=variant, seed=3, first_generation=1)
    one = ShardedSwarm(DeviceShard(ctx, prob, x, f, 0, 2), variant=variant, seed=3, first_generation=1)
    one.evolve(5)
    bx, bf = one.shard.best()
    assert np.array_equal(bx, want_x) and np.array_equal(bf, want_f)
    # two shards of 48 particles, exchanged by hand
    shards = [DeviceShard(ctx, prob, x[i * 48:(i + 1) * 48], f[i * 48:(i + 1) * 48], i * 48, 2) for i in range(2)]
    p = dict(omega=0.7298, eta1=2.05, eta2=2.05, max_vel=0.5, variant=variant, seed=3)
    for s in shards:
        s.step(p, 1, init_velocity=True)
    for g in range(1, 6):
        blocks = np.stack([s.boundary() for s in shards])
        for r, s in enumerate(shards):
            s.set_halos(*halo_rows(blocks, r, 2, 2))
        for s in shards:
            s.step(p, g)
    got = np.vstack([s.best()[0] for s in shards])
    assert np.array_equal(got, want_x)
    prob.close()


@pytest.mark.parametrize("variant", (1, 3, 5))
def test_sharded_gbest_swarm_equals_evolve(capi, ctx, variant):
    """gbest topology (pso_gen.cpp:644-677, tracking :452-457) sharded: one shard, and three shards in one process whose candidates
    are reduced by hand with the same functions GbestSwarm uses after its all_gather, must move exactly like pgc_pso_evolve_device."""
    from pagmo2_b200.swarm import GbestShard, GbestSwarm, pick_initial_best, pick_next_best
    rng = np.random.default_rng(10 + variant)
    prob = capi.Problem(ctx, "rastrigin", dim=6)
    lb, ub = prob.bounds()
    n, gens = 90, 8
    x = rng.uniform(lb, ub, (n, 6))
    f = prob.eval_host(x)[:, 0]
    want_x, want_f, _, _ = prob.pso_evolve(x, f, gens=gens, variant=variant, neighb_type=1, seed=5, first_generation=1)
    one = GbestSwarm(GbestShard(ctx, prob, x, f, 0), variant=variant, seed=5, first_generation=1)
    one.evolve(gens)
    bx, bf = one.shard.best()
    assert np.array_equal(bx, want_x) and np.array_equal(bf, want_f)
    shards = [GbestShard(ctx, prob, x[i * 30:(i + 1) * 30], f[i * 30:(i + 1) * 30], i * 30) for i in range(3)]
    p = dict(omega=0.7298, eta1=2.05, eta2=2.05, max_vel=0.5, variant=variant, seed=5)
    for s in shards:
        s.step(p, 1, init_velocity=True)
    current = pick_initial_best(np.stack([s.initial_candidate() for s in shards]))
    for s in shards:
        s.set_best(current)
    for g in range(1, gens + 1):
        for s in shards:
            s.step(p, g)
        winner = pick_next_best(np.stack([s.candidate() for s in shards]), current[0])
        if winner is not None:
            current = winner
            for s in shards:
                s.set_best(current)
    assert np.array_equal(np.vstack([s.best()[0] for s in shards]), want_x)
    assert np.array_equal(np.concatenate([s.best()[1] for s in shards]), want_f)
    assert current[0] == want_f.min()
    prob.close()
