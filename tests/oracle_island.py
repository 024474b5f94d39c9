"""Test helper: an island whose numerics come from the CPU oracle (restated population init, DE family, select_best, fair_replace).
It implements the interface `pagmo2_b200.archipelago.Archipelago` expects, so the archipelago's host logic and its exchange can run
without a GPU, and so a device archipelago can be checked round by round.  TEST INFRASTRUCTURE ONLY."""
import numpy as np

from pagmo2_b200.archipelago import Group


class OracleIsland:
    def __init__(self, orc, family, dim, pop_size, seed, algo="sade", gens=3, r_rate=1, s_rate=1, algo_seed=0, func=0, **algo_kw):
        self.orc, self.n = orc, pop_size
        if family == "cec2013":
            self.prob = orc.problem("cec2013", prob_id=func, dim=dim, tables=orc.cec2013_tables(dim))
            self.lb, self.ub = np.full(dim, -100.0), np.full(dim, 100.0)
            self._eval = lambda x: orc.cec2013(func, x)
        else:
            self.prob = orc.problem(family, dim=dim)
            bounds = {"rastrigin": 5.12, "ackley": 15.0, "griewank": 600.0, "schwefel": 500.0, "rosenbrock": None}[family]
            self.lb, self.ub = (np.full(dim, -5.0), np.full(dim, 10.0)) if bounds is None else (np.full(dim, -bounds), np.full(dim, bounds))
            if family == "ackley":
                self.lb, self.ub = np.full(dim, -15.0), np.full(dim, 30.0)
            self._eval = lambda x: orc.simple(family, x)
        self.nx, self.nf = dim, 1
        self.algo, self.gens, self.algo_seed, self.algo_kw = algo, gens, algo_seed, algo_kw
        self.r_rate, self.s_rate = r_rate, s_rate
        self.generation = 1
        self.x, self.ids_ = orc.population_init(self.lb, self.ub, pop_size, seed)
        self.f = self._eval(self.x).reshape(pop_size, 1)

    def evolve(self):
        kw = dict(variant=2, variant_adptv=1, ftol=1e-6, xtol=1e-6)
        kw.update(self.algo_kw)
        x, f, *_ = self.orc.de_evolve(self.prob, self.lb, self.ub, self.x, self.f[:, 0], gens=self.gens, algo=self.algo, seed=self.algo_seed,
                                      first_generation=self.generation, **kw)
        self.x, self.f = x, f.reshape(-1, 1)
        self.generation += max(self.gens, 1)

    def select(self) -> Group:
        return Group(*self.orc.select_best(self.ids_, self.x, self.f, self.s_rate))

    def replace(self, mig: Group):
        self.ids_, self.x, self.f = (a.copy() for a in self.orc.fair_replace(self.ids_, self.x, self.f, self.r_rate, mig.ids, mig.x, mig.f))

    def population(self) -> Group:
        return Group(self.ids_.copy(), self.x.copy(), self.f.copy())

    def ids(self):
        return self.ids_
