"""CPU tests of the migration path: restated policies against the compiled reference, topologies against the reference's own test
(tests/ring.cpp:47-74, tests/fully_connected.cpp), the archipelago's host logic, and the N>1 exchange with world_size 2 over gloo."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))


def test_policies_restatement_vs_reference(orc, ref):
    rng = np.random.default_rng(3)
    for nobj in (1, 2, 3):
        for n, nm, rate in ((20, 5, 1), (20, 5, 3), (20, 0, 2), (7, 9, 0.5), (16, 16, 1.0), (5, 3, 0), (64, 64, 0.1)):
            ids = rng.integers(0, 2**63, n, dtype=np.uint64)
            x, f = rng.normal(size=(n, 4)), rng.normal(size=(n, nobj))
            mids = rng.integers(0, 2**63, nm, dtype=np.uint64)
            mx, mf = rng.normal(size=(nm, 4)), rng.normal(size=(nm, nobj))
            # select_best_N_mo cuts the last front by crowding distance, where the boundary points tie at +inf: the reference's
            # std::sort leaves their order unspecified (and introsort reorders them beyond 16 elements), so multi-objective groups
            # of that size are compared as sets of individuals
            as_set = nobj > 1 and n + nm > 16

            def same(a, b):
                if as_set:
                    oa, ob = np.argsort(a[0]), np.argsort(b[0])
                    return all(np.array_equal(u[oa], v[ob]) for u, v in zip(a, b))
                return all(np.array_equal(u, v) for u, v in zip(a, b))

            assert same(orc.fair_replace(ids, x, f, rate, mids, mx, mf), ref.fair_replace(ids, x, f, rate, mids, mx, mf)), (nobj, n, nm, rate)
            assert same(orc.select_best(ids, x, f, rate), ref.select_best(ids, x, f, rate)), (nobj, n, nm, rate)
    # NaN fitness sorts last (detail::less_than_f); absolute rate above the population size throws (fair_replace.cpp:92-99)
    f = np.array([[3.0], [np.nan], [1.0], [2.0]])
    ids = np.arange(4, dtype=np.uint64)
    for impl in (orc, ref):
        assert list(impl.select_best(ids, np.zeros((4, 2)), f, 4)[0]) == [2, 3, 0, 1]
        with pytest.raises(Exception):
            impl.fair_replace(ids, np.zeros((4, 2)), f, 5, ids, np.zeros((4, 2)), f)
        with pytest.raises(Exception):
            impl.select_best(ids, np.zeros((4, 2)), f, 5)


def _constrained_group(rng, n, nec, nic, mixed):
    """rows [f | nec equality | nic inequality constraints]: about a third feasible, the others violating a few constraints; unless
    `mixed`, an individual violates constraints of ONE kind only (there compare_fc's two norms coincide, constrained.cpp:96-106)."""
    f = np.empty((n, 1 + nec + nic))
    f[:, 0] = rng.normal(size=n)
    eq = rng.normal(scale=1e-3, size=(n, nec))          # within the tolerance of 1e-2
    ineq = -np.abs(rng.normal(size=(n, nic)))           # satisfied
    for i in range(n):
        kind = rng.integers(0, 3)
        if kind == 1 or (mixed and kind == 2):
            if nec:
                cols = rng.choice(nec, size=rng.integers(1, nec + 1), replace=False)
                eq[i, cols] = rng.normal(scale=2.0, size=cols.size) + 0.5
        if (kind == 2 or (mixed and kind == 1)) and nic:
            cols = rng.choice(nic, size=rng.integers(1, nic + 1), replace=False)
            ineq[i, cols] = np.abs(rng.normal(scale=2.0, size=cols.size)) + 0.05
    f[:, 1:1 + nec] = eq
    f[:, 1 + nec:] = ineq
    return f


def test_constrained_policies_restatement_vs_reference(orc, ref):
    """the single-objective constrained branches (select_best.cpp:137-152, fair_replace.cpp:158-188, sort_population_con): the stable
    restatement against the compiled reference on tie-free groups, with equality-only, inequality-only and both kinds of constraints."""
    rng = np.random.default_rng(11)
    for nec, nic in ((2, 0), (0, 3), (2, 2), (1, 4)):
        tol = np.concatenate([np.full(nec, 1e-2), np.full(nic, 1e-2)])
        for n, nm, rate in ((20, 5, 1), (20, 5, 3), (12, 9, 0.5), (16, 16, 1.0), (14, 6, 4)):  # <= 16 + ...: see the note on introsort above
            ids = rng.integers(0, 2**63, n, dtype=np.uint64)
            mids = rng.integers(0, 2**63, nm, dtype=np.uint64)
            x, mx = rng.normal(size=(n, 3)), rng.normal(size=(nm, 3))
            f, mf = _constrained_group(rng, n, nec, nic, False), _constrained_group(rng, nm, nec, nic, False)
            a, b = orc.select_best_con(ids, x, f, rate, nec, nic, tol), ref.select_best_con(ids, x, f, rate, nec, nic, tol)
            assert all(np.array_equal(u, v) for u, v in zip(a, b)), (nec, nic, n, rate)
            a = orc.fair_replace_con(ids, x, f, rate, mids, mx, mf, nec, nic, tol)
            b = ref.fair_replace_con(ids, x, f, rate, mids, mx, mf, nec, nic, tol)
            assert all(np.array_equal(u, v) for u, v in zip(a, b)), (nec, nic, n, nm, rate)
    # NaN constraints: std::max(NaN - tol, 0.) stays NaN (constrained.hpp:55,73), so the constraint is never satisfied and the norm is NaN
    for nec, nic in ((2, 0), (0, 3), (2, 2)):
        tol = np.full(nec + nic, 1e-2)
        for trial in range(10):
            n = 14
            ids = rng.integers(0, 2**63, n, dtype=np.uint64)
            x = rng.normal(size=(n, 3))
            f = _constrained_group(rng, n, nec, nic, False)
            f[3, 1] = np.nan
            if trial % 2:
                f[9, nec + nic] = np.nan
            a, b = orc.select_best_con(ids, x, f, n, nec, nic, tol), ref.select_best_con(ids, x, f, n, nec, nic, tol)
            assert all(np.array_equal(u, v, equal_nan=True) for u, v in zip(a, b)), (nec, nic, trial)
    # the order itself: feasible first by objective, then by the number of violated constraints, then by the violation norm
    f = np.array([[5.0, 0.0, -1.0], [1.0, 0.5, -1.0], [9.0, 0.0, 2.0], [2.0, 0.0, -1.0], [0.0, 3.0, 1.0], [7.0, 0.2, -1.0]])
    assert list(orc.sort_population_con(f, 1, 1, [1e-2, 1e-2])) == [3, 0, 5, 1, 2, 4]


@pytest.mark.parametrize("kind", ("ring", "fully_connected"))
def test_topologies(orc, kind):
    from pagmo2_b200 import capi
    for n in range(1, 12):
        for i in range(n):
            src, w = capi.topology_connections(kind, n, i, 0.5)
            assert list(src) == list(orc.connections(kind, n, i)) and (w == 0.5).all()
            if kind == "ring":  # verify_ring_topology, reference tests/ring.cpp:47-74
                assert len(src) == (0 if n < 2 else 1 if n == 2 else 2)
                if n >= 2:
                    assert (i + 1) % n in src and (i - 1) % n in src
            else:
                assert list(src) == [j for j in range(n) if j != i]
    with pytest.raises(capi.PgcError):
        capi.topology_connections(kind, 3, 3)
    with pytest.raises(capi.PgcError):
        capi.topology_connections(kind, 3, 0, 1.5)
    assert len(capi.topology_connections("unconnected", 4, 2)[0]) == 0


def _archi(orc, n_islands, topology, mtype, handling, **kw):
    from oracle_island import OracleIsland
    from pagmo2_b200.archipelago import Archipelago
    return Archipelago(n_islands, lambda g: OracleIsland(orc, "rastrigin", 6, 16, seed=100 + g, algo="sade", gens=2, algo_seed=7 + g, s_rate=2,
                                                         r_rate=2, ftol=0.0, xtol=0.0),
                       topology=topology, weight=0.75, migration_type=mtype, migrant_handling=handling, seed=5, **kw)


def test_archipelago_host_logic(orc):
    # unconnected islands never exchange anything and evolve exactly like stand-alone islands
    a = _archi(orc, 3, "unconnected", "p2p", "preserve")
    a.evolve(3)
    from oracle_island import OracleIsland
    solo = OracleIsland(orc, "rastrigin", 6, 16, seed=101, algo="sade", gens=2, algo_seed=8, ftol=0.0, xtol=0.0)
    for _ in range(3):
        solo.evolve()
    assert np.array_equal(a.islands[1].population().x, solo.x) and not a.log
    # ring: migrants enter, the log only holds IDs that are inside the destination afterwards, the database holds s_rate rows
    for mtype in ("p2p", "broadcast"):
        for handling in ("preserve", "evict"):
            a = _archi(orc, 4, "ring", mtype, handling)
            before = [isl.population().f.min() for isl in a.islands]
            a.evolve(4)
            assert a.log, (mtype, handling)
            assert all(len(g) in (0, 2) for g in a.db) and (handling == "evict" or all(len(g) == 2 for g in a.db))
            for e in a.log:
                assert e.src in a.conn[e.dst][0] and e.src != e.dst
            assert all(isl.population().f.min() <= b for isl, b in zip(a.islands, before))
            for isl in a.islands:  # fair_replace leaves the population sorted only right after a migration; sizes never change
                assert isl.population().x.shape == (16, 6)
    with pytest.raises(ValueError):
        _archi(orc, 4, "ring", "p2p", "keep")


@pytest.mark.parametrize("mtype,handling", (("p2p", "preserve"), ("broadcast", "evict")))
def test_world_size_2_gloo_matches_single_process(orc, tmp_path, mtype, handling):
    """4 islands on 2 processes (2 each, gloo all_gather of the migrants database) == the same archipelago in one process."""
    out = tmp_path / "archi"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    port = 29500 + (os.getpid() + (0 if mtype == "p2p" else 1)) % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           str(port), str(ROOT / "tests" / "dist_archi_worker.py"), str(out), "4", "ring", mtype, handling, "3"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    single = _archi(orc, 4, "ring", mtype, handling)
    single.evolve(3)
    log = []
    for rank in range(2):
        z = np.load(f"{out}.rank{rank}.npz")
        assert int(z["first"]) == 2 * rank
        for i in range(2):
            p = single.islands[2 * rank + i].population()
            assert np.array_equal(z["x"][i], p.x) and np.array_equal(z["f"][i], p.f) and np.array_equal(z["ids"][i], p.ids)
        log += [tuple(row) for row in z["log"]]
    assert sorted(log) == sorted((e.round, e.id % (1 << 62), e.src, e.dst) for e in single.log) and log


def test_swarm_halo_wiring():
    """ring wiring of the sharded PSO swarm's halo exchange (pagmo2_b200/swarm.py): shard r gets the LAST rows of shard r-1 and the
    FIRST rows of shard r+1, with wrap-around; one shard wraps onto itself."""
    from pagmo2_b200.swarm import halo_rows
    radius, world = 2, 3
    blocks = np.stack([np.array([[10 * r + 0], [10 * r + 1], [10 * r + 8], [10 * r + 9]], dtype=float) for r in range(world)])
    for r in range(world):
        left, right = halo_rows(blocks, r, world, radius)
        assert left[:, 0].tolist() == [10 * ((r - 1) % world) + 8, 10 * ((r - 1) % world) + 9]
        assert right[:, 0].tolist() == [10 * ((r + 1) % world) + 0, 10 * ((r + 1) % world) + 1]
    left, right = halo_rows(blocks[:1], 0, 1, radius)
    assert left[:, 0].tolist() == [8, 9] and right[:, 0].tolist() == [0, 1]


def test_swarm_exchange_world_size_2_gloo(tmp_path):
    """the all_gather of boundary rows over gloo: two fake shards (no device) must end up with each other's rows as halos."""
    worker = tmp_path / "w.py"
    worker.write_text('''
import sys, numpy as np, torch.distributed as dist
sys.path.insert(0, %r)
from pagmo2_b200.swarm import ShardedSwarm
class Fake:
    radius = 2
    def __init__(self, r): self.r, self.halos = r, None
    def boundary(self): return np.array([[self.r, 0.], [self.r, 1.], [self.r, 8.], [self.r, 9.]])
    def set_halos(self, left, right): self.halos = (left.copy(), right.copy())
    def step(self, p, generation, init_velocity=False): pass
dist.init_process_group("gloo")
s = ShardedSwarm(Fake(dist.get_rank()))
s.evolve(2)
other = 1 - dist.get_rank()
left, right = s.shard.halos
assert left.tolist() == [[other, 8.], [other, 9.]] and right.tolist() == [[other, 0.], [other, 1.]], (left, right)
assert s.generation == 3
dist.barrier(); dist.destroy_process_group()
''' % str(ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           str(31500 + os.getpid() % 2000), str(worker)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


def test_gbest_reduction_rules():
    """pick_initial_best = pop.best_idx() (first minimum); pick_next_best = the sequential scan of pso_gen.cpp:452-457 over the
    gathered candidates: smallest fitness, LAST global index on ties, accepted if <= the current best."""
    from pagmo2_b200.swarm import pick_initial_best, pick_next_best
    c = np.array([[3.0, 40, 1.0], [1.0, 70, 2.0], [1.0, 10, 3.0]])
    assert pick_initial_best(c).tolist() == [1.0, 10, 3.0]
    assert pick_next_best(c, 5.0).tolist() == [1.0, 70, 2.0]
    assert pick_next_best(c, 1.0).tolist() == [1.0, 70, 2.0]  # equal fitness replaces (<=)
    assert pick_next_best(c, 0.5) is None
    none = np.array([[np.inf, -1, 0.0], [np.inf, -1, 0.0]])
    assert pick_next_best(none, 5.0) is None
    mixed = np.array([[np.inf, -1, 0.0], [2.0, 5, 9.0]])
    assert pick_next_best(mixed, 2.5).tolist() == [2.0, 5, 9.0]


def test_gbest_swarm_world_size_2_gloo(tmp_path):
    """GbestSwarm over gloo with fake shards: both ranks must agree on the swarm's best after every generation."""
    worker = tmp_path / "w.py"
    worker.write_text('''
import sys, numpy as np, torch.distributed as dist
sys.path.insert(0, %r)
from pagmo2_b200.swarm import GbestSwarm
class Fake:
    def __init__(self, r): self.r, self.seen, self.g = r, [], 0
    def initial_candidate(self): return np.array([10.0 - self.r, 100 * self.r + 3, float(self.r)])
    def candidate(self):
        # generation 1: only rank 0 improves (to 5); generation 2: both reach 4 -> the larger global index (rank 1) wins; 3: nobody
        if self.g == 1: return np.array([5.0, 7, 0.5]) if self.r == 0 else np.array([np.inf, -1, 0.0])
        if self.g == 2: return np.array([4.0, 100 * self.r + 1, 10.0 + self.r])
        return np.array([np.inf, -1, 0.0])
    def set_best(self, cand): self.seen.append(cand.tolist())
    def step(self, p, generation, init_velocity=False):
        if not init_velocity: self.g += 1
dist.init_process_group("gloo")
s = GbestSwarm(Fake(dist.get_rank()))
s.evolve(3)
assert s.shard.seen == [[9.0, 103, 1.0], [5.0, 7, 0.5], [4.0, 101, 11.0]], s.shard.seen
assert s.generation == 4
dist.barrier(); dist.destroy_process_group()
''' % str(ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           str(33500 + os.getpid() % 2000), str(worker)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
