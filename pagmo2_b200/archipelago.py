"""Archipelago of device-resident islands with migration: host-side mirror of pagmo::archipelago / pagmo::island for the path
`island::evolve` (reference src/island.cpp:428-652): pull migrants -> replacement policy -> algorithm evolve -> selection policy
-> publish, per island and per round.

What runs where
  * populations live in HBM for the whole run; the algorithm (`pgc_algo_evolve_device`), select_best and fair_replace
    (`pgc_select_best_device`, `pgc_fair_replace_device`: migration.cu) are device code behind the C ABI;
  * the decisions of island.cpp:461-620 (which neighbour, Bernoulli(weight) per edge, p2p vs broadcast, preserve vs evict) are
    taken on the host exactly as in the reference, from Philox draws (seed, tag 9, round, island, slot) instead of a
    random_device-seeded mt19937 (island.cpp:467-469) so that a run is reproducible on any number of processes;
  * the migrants database (archipelago::set_migrants / get_migrants / extract_migrants) is replicated on every process: after
    each round the processes exchange what their islands published (k rows of ids | x | f per island - a few hundred bytes) with
    ONE all_gather over torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests).  This is the only collective of the
    path; evaluation and variation never communicate.

Differences from the reference, all deliberate: islands advance in lock step (round r pulls what round r-1 published), where the
reference's island threads race (island.cpp:461 reads whatever the neighbours have published so far); with `evict` the islands
pull in index order.  Both make the run deterministic.

This module holds no numerical code and has no CPU fallback: `DeviceIsland` calls libpgc.so.  The `backend` argument exists so
that the host logic and the exchange can be exercised on a machine without a GPU (tests/ pass oracle-backed islands).
"""
from __future__ import annotations

import ctypes as C
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field

import time

import numpy as np

from . import capi

TAG_MIGRATE = 9


@dataclass
class Group:
    """individuals_group_t (types.hpp:57): ids [k], x [k x nx], f [k x nf]."""
    ids: np.ndarray
    x: np.ndarray
    f: np.ndarray

    @staticmethod
    def empty(nx: int, nf: int) -> "Group":
        return Group(np.empty(0, dtype=np.uint64), np.empty((0, nx)), np.empty((0, nf)))

    def __len__(self) -> int:
        return int(self.ids.shape[0])

    @staticmethod
    def concat(groups, nx: int, nf: int) -> "Group":
        groups = [g for g in groups if len(g)]
        if not groups:
            return Group.empty(nx, nf)
        return Group(np.concatenate([g.ids for g in groups]), np.vstack([g.x for g in groups]), np.vstack([g.f for g in groups]))


class DeviceIsland:
    """One island: a problem, an algorithm and a population resident on one GPU (own context = own stream, so the islands of a
    process overlap on the device like the reference's island threads, thread_island.cpp:79-159)."""

    def __init__(self, device: int, family: str, algo: capi.AlgoDesc, pop_size: int, seed: int, r_rate=1, s_rate=1, **problem_kw):
        self.ctx = capi.Context(device)
        self.prob = capi.Problem(self.ctx, family, **problem_kw)
        self.algo, self.n, self.nx, self.nf = algo, pop_size, self.prob.nx, self.prob.nf
        self.r_rate, self.s_rate = r_rate, s_rate
        self.generation = 1
        self.d_x = self.ctx.malloc(8 * pop_size * self.nx)
        self.d_f = self.ctx.malloc(8 * pop_size * self.nf)
        self.d_ids = self.ctx.malloc(8 * pop_size)
        self._m = [self.ctx.malloc(8 * pop_size * w) for w in (1, self.nx, self.nf)]  # staging for groups in flight
        capi.check(capi.lib().pgc_population_init_device(self.prob._h, pop_size, seed, self.d_x, self.d_f, self.d_ids, None))

    @staticmethod
    def _rate(rate):
        return (1, float(rate)) if isinstance(rate, float) else (0, float(rate))

    def evolve(self):
        done = C.c_uint()
        capi.check(capi.lib().pgc_algo_evolve_device(self.prob._h, C.byref(self.algo), self.d_x, self.d_f, self.n, self.generation,
                                                     C.byref(done), None))
        self.generation += max(int(self.algo.gens), 1)

    def select(self) -> Group:
        frac, rate = self._rate(self.s_rate)
        k = C.c_size_t()
        capi.check(capi.lib().pgc_select_best_device(self.ctx._h, self.d_ids, self.d_x, self.d_f, self.n, self.nx, self.nf, frac, rate,
                                                     self._m[0], self._m[1], self._m[2], C.byref(k), None))
        k = k.value
        return Group(self.ctx.from_device(self._m[0], (k,), np.uint64), self.ctx.from_device(self._m[1], (k, self.nx)),
                     self.ctx.from_device(self._m[2], (k, self.nf)))

    def replace(self, mig: Group):
        frac, rate = self._rate(self.r_rate)
        bufs = [0, 0, 0]
        if len(mig):
            bufs = [self.ctx.to_device(a) for a in (mig.ids.astype(np.uint64), mig.x, mig.f)]
        try:
            capi.check(capi.lib().pgc_fair_replace_device(self.ctx._h, self.d_ids, self.d_x, self.d_f, self.n, self.nx, self.nf, frac, rate,
                                                          bufs[0] or None, bufs[1] or None, bufs[2] or None, len(mig), None))
            self.ctx.synchronize()
        finally:
            for b in bufs:
                if b:
                    self.ctx.free(b)

    def population(self) -> Group:
        self.ctx.synchronize()
        return Group(self.ctx.from_device(self.d_ids, (self.n,), np.uint64), self.ctx.from_device(self.d_x, (self.n, self.nx)),
                     self.ctx.from_device(self.d_f, (self.n, self.nf)))

    def ids(self) -> np.ndarray:
        return self.ctx.from_device(self.d_ids, (self.n,), np.uint64)


@dataclass
class MigrationEntry:
    """one row of archipelago::get_migration_log() (archipelago.hpp:103): a migrant that entered `dst`'s population."""
    round: int
    id: int
    src: int
    dst: int


@dataclass
class Archipelago:
    """n_islands islands; this process owns islands [rank*L, (rank+1)*L), L = n_islands / world_size.

    make_island(g) builds global island g (an object with evolve / select / replace / population / ids and attributes nx, nf).
    topology: 'unconnected' | 'ring' | 'fully_connected' with edge `weight` (topologies/ring.hpp, fully_connected.hpp);
    migration_type 'p2p' | 'broadcast', migrant_handling 'preserve' | 'evict' (archipelago.hpp:71-88; defaults as the reference)."""
    n_islands: int
    make_island: callable
    topology: str = "unconnected"
    weight: float = 1.0
    migration_type: str = "p2p"
    migrant_handling: str = "preserve"
    seed: int = 0
    group: object = None  # torch.distributed process group (None: default group when initialised, else single process)
    distributed: bool = True  # False: ignore an initialised torch.distributed and own every island in this process
    log: list = field(default_factory=list)

    def __post_init__(self):
        self.rank, self.world = 0, 1
        self._dist = None
        try:
            import torch.distributed as dist
            if self.distributed and dist.is_available() and dist.is_initialized():
                self._dist = dist
                self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        except ImportError:
            pass
        if self.n_islands % self.world:
            raise ValueError(f"{self.n_islands} islands cannot be split evenly over {self.world} processes")
        if self.migration_type not in ("p2p", "broadcast") or self.migrant_handling not in ("preserve", "evict"):
            raise ValueError("migration_type must be 'p2p' or 'broadcast', migrant_handling 'preserve' or 'evict'")
        self.local = self.n_islands // self.world
        self.first = self.rank * self.local
        self.islands = [self.make_island(self.first + i) for i in range(self.local)]
        self.nx, self.nf = self.islands[0].nx, self.islands[0].nf
        self.conn = [capi.topology_connections(self.topology, self.n_islands, g, self.weight) for g in range(self.n_islands)]
        self.db = [Group.empty(self.nx, self.nf) for _ in range(self.n_islands)]  # archipelago migrants database
        self.round = 0
        self._pool = ThreadPoolExecutor(max_workers=max(self.local, 1))

    # island.cpp:461-620 for destination island g: the migrants it pulls this round, or None when no replacement takes place
    def _pull(self, g: int):
        src, w = self.conn[g]
        if len(src) == 0:
            return None
        take = (lambda s: self.db[s]) if self.migrant_handling == "preserve" else self._extract
        u = lambda slot: capi.philox_u01(self.seed, TAG_MIGRATE, self.round, g, slot)
        if self.migration_type == "p2p":
            j = min(int(u(0) * len(src)), len(src) - 1)  # uniform_int_distribution(0, size-1), :497-499
            if not u(1) < w[j]:                           # :502
                return None
            s = int(src[j])
            return [(s, take(s))]
        return [(int(s), take(int(s))) for j, s in enumerate(src) if u(j) < w[j]]  # :551-575

    def _extract(self, s: int) -> Group:  # archipelago::extract_migrants: the entry is emptied
        g, self.db[s] = self.db[s], Group.empty(self.nx, self.nf)
        return g

    def _step_island(self, isl, g: int, pulled):
        if pulled is not None:
            mig = Group.concat([m for _, m in pulled], self.nx, self.nf)
            isl.replace(mig)  # r_pol.replace + set_individuals, :511-517 / :578-585
            if len(mig):      # migration log: migrants that made it into the population, :525-536 / :594-607
                inside = set(int(i) for i in isl.ids())
                entries = [MigrationEntry(self.round, int(i), s, g) for s, m in pulled for i in m.ids if int(i) in inside]
            else:
                entries = []
        else:
            entries = []
        isl.evolve()           # isl_ptr->run_evolve, :623
        return isl.select(), entries  # s_pol.select + set_migrants, :629-640

    def evolve(self, rounds: int = 1):
        """archipelago::evolve(n) + wait_check(): every island runs `rounds` x (migrate in, evolve, publish)."""
        for _ in range(rounds):
            pulls = [self._pull(g) for g in range(self.n_islands)]  # replicated on every process, in island order
            futs = [self._pool.submit(self._step_island, isl, self.first + i, pulls[self.first + i]) for i, isl in enumerate(self.islands)]
            results = [f.result() for f in futs]
            for _, entries in results:
                self.log.extend(entries)
            self._publish([r[0] for r in results])
            self.round += 1

    def _publish(self, local_groups):
        """set_migrants for the local islands, then one all_gather so that every process holds the whole database."""
        if self.world == 1:
            for i, gr in enumerate(local_groups):
                self.db[self.first + i] = gr
            return
        import torch
        dist = self._dist
        backend = dist.get_backend(self.group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        kmax = max((len(g) for g in local_groups), default=0)
        kt = torch.tensor([kmax], dtype=torch.int64, device=dev)
        dist.all_reduce(kt, op=dist.ReduceOp.MAX, group=self.group)
        kmax = int(kt.item())
        width = 1 + self.nx + self.nf  # ids (bit pattern in a double slot) | x | f
        pack = np.zeros((self.local, 1 + kmax * width))
        for i, gr in enumerate(local_groups):
            k = len(gr)
            pack[i, 0] = k
            if k:
                rows = np.concatenate([gr.ids.astype(np.uint64).view(np.float64)[:, None], gr.x, gr.f], axis=1)
                pack[i, 1:1 + k * width] = rows.reshape(-1)
        mine = torch.from_numpy(pack).to(dev)
        everything = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(everything, mine, group=self.group)
        for r, t in enumerate(everything):
            a = t.cpu().numpy()
            for i in range(self.local):
                k = int(a[i, 0])
                rows = a[i, 1:1 + k * width].reshape(k, width)
                self.db[r * self.local + i] = Group(rows[:, 0].copy().view(np.uint64), rows[:, 1:1 + self.nx].copy(),
                                                    rows[:, 1 + self.nx:].copy())

    def champions_f(self) -> np.ndarray:
        """archipelago::get_champions_f() of the LOCAL islands (single objective): best fitness per island."""
        return np.array([isl.population().f[:, 0].min() for isl in self.islands])


class ResidentArchipelago:
    """The harness-side twin of pagmo_cuda::cuda_archipelago (include/pagmo_cuda/cuda_island.hpp): islands are pgc_island objects
    (population, outbox, inbox in HBM) and the migrants travel device to device through pgc_migrate - ncclSend / ncclRecv along the
    topology's edges between GPUs, a device copy inside one GPU.  No torch.distributed collective, no host copy of any migrant.

    islands: list of dicts {device, family, problem_kw, algo (capi.AlgoDesc), pop_size, seed, r_rate, s_rate, owner} in GLOBAL
    island order; every process passes the same list and materialises the islands whose owner rank is one of its own.
    comm: capi.Comm (None when every island sits on one GPU); my_ranks: the communicator ranks this process drives."""

    def __init__(self, islands, topology="ring", weight=1.0, migration_type="p2p", migrant_handling="preserve", seed=0, comm=None,
                 my_ranks=(0,), log=True):
        if migration_type not in ("p2p", "broadcast") or migrant_handling not in ("preserve", "evict"):
            raise ValueError("migration_type must be 'p2p' or 'broadcast', migrant_handling 'preserve' or 'evict'")
        self.spec, self.comm, self.seed, self.round = islands, comm, seed, 0
        self.migration_type, self.migrant_handling, self.log_on, self.log = migration_type, migrant_handling, log, []
        G = len(islands)
        self.conn = [capi.topology_connections(topology, G, g, weight) for g in range(G)]
        self.k_out = [min(int(s["s_rate"] * s["pop_size"]) if isinstance(s["s_rate"], float) else int(s["s_rate"]), s["pop_size"]) for s in islands]
        cap, max_in = max(self.k_out + [1]), max([len(c[0]) for c in self.conn] + [1])
        self.owner = [int(s.get("owner", 0)) for s in islands]
        self.isl, self.ctx, self.prob = [None] * G, [None] * G, [None] * G
        for g, s in enumerate(islands):
            if self.owner[g] not in my_ranks:
                continue
            self.ctx[g] = capi.Context(s["device"])  # own context = own stream: islands sharing a GPU overlap on it
            self.prob[g] = capi.Problem(self.ctx[g], s["family"], **s.get("problem_kw", {}))
            self.isl[g] = capi.Island(self.prob[g], s["pop_size"], cap, max_in)
            self.isl[g].init(s["seed"])
        self.published = [False] * G
        self._pending = {}  # island -> (round, sources) of a replace whose migration-log rows have not been read back yet
        self.phase_seconds = {}  # host wall time per phase of evolve(), summed over rounds
        self.local = [g for g in range(G) if self.isl[g] is not None]
        for g in self.local:  # islands that share a GPU fill it together: each keeps fuller tiles (pgc_ctx_set_sharers)
            capi.check(capi.lib().pgc_ctx_set_sharers(self.ctx[g]._h, sum(1 for h in self.local if islands[h]["device"] == islands[g]["device"])))
        self._pool = ThreadPoolExecutor(max_workers=max(len(self.local), 1))

    def _u(self, g, slot):
        return capi.philox_u01(self.seed, TAG_MIGRATE, self.round, g, slot)

    def _collect(self, g):
        """migration-log rows of island g's last replace (enqueued one round earlier): waits for that replace's log copies only"""
        pend = self._pending.pop(g, None)
        if pend is None:
            return []
        rnd, sources = pend
        return [MigrationEntry(rnd, a_id, sources[slot][0], g) for a_id, slot in self.isl[g].replace_collect()]

    def _step(self, g, replace, sources):
        # nothing here blocks on the device but _collect, whose wait ends when the PREVIOUS round's replace has run: replace, evolve and
        # select of this round are enqueued back to back, so the GPU goes from one round's generations into the next one's without
        # waiting for the host (the slot counts are the senders' policy counts, known to every process)
        isl, s = self.isl[g], self.spec[g]
        entries = self._collect(g)
        if replace:
            counts = [self.k_out[src] if had else 0 for src, had in sources]
            isl.replace_enqueue(s["r_rate"], counts, log=self.log_on)
            if self.log_on:
                self._pending[g] = (self.round, sources)
        isl.evolve(s["algo"])
        isl.select(s["s_rate"])
        return entries

    def evolve(self, rounds: int = 1):
        G = len(self.spec)
        for _ in range(rounds):
            pulls, edges = [(False, []) for _ in range(G)], []
            for g in range(G):  # island.cpp:461-575, replicated on every process in island order
                src, w = self.conn[g]
                if len(src) == 0:
                    continue
                sources = []

                def take(s):
                    had = self.published[s]
                    if self.migrant_handling == "evict":
                        self.published[s] = False
                    sources.append((int(s), had))
                if self.migration_type == "p2p":
                    j = min(int(self._u(g, 0) * len(src)), len(src) - 1)
                    if not self._u(g, 1) < w[j]:
                        continue
                    take(int(src[j]))
                else:
                    for j, s in enumerate(src):
                        if self._u(g, j) < w[j]:
                            take(int(s))
                pulls[g] = (True, sources)
                edges += [(s, g, q) for q, (s, had) in enumerate(sources) if had]
            t0 = time.perf_counter()
            for g in self.local:
                for q, (s, had) in enumerate(pulls[g][1]):
                    if not had:
                        self.isl[g].inbox_upload(q, np.empty(0, np.uint64), np.empty((0, self.isl[g].nx)), np.empty((0, self.isl[g].nf)))
            t1 = time.perf_counter()
            capi.migrate(self.comm, self.isl, self.owner, edges)
            t2 = time.perf_counter()
            futs = [self._pool.submit(self._step, g, *pulls[g]) for g in self.local]
            for f in futs:
                self.log.extend(f.result())
            t3 = time.perf_counter()
            for k, dt in (("empty_inbox_upload", t1 - t0), ("migrate_enqueue", t2 - t1), ("replace_evolve_select", t3 - t2)):
                self.phase_seconds[k] = self.phase_seconds.get(k, 0.0) + dt
            self.published = [k > 0 for k in self.k_out]
            self.round += 1
        for g in self.local:  # the last round's log rows
            self.log.extend(self._collect(g))

    def synchronize(self):
        for g in self.local:
            self.ctx[g].synchronize()

    def champions_f(self) -> np.ndarray:
        return np.array([self.isl[g].champion()[1] for g in self.local])

    def populations(self):
        return {g: self.isl[g].download() for g in self.local}
