"""One pso_gen swarm (lbest ring topology, the reference's default: pso_gen.hpp:127, pso_gen.cpp:679-698) sharded over several GPUs.

Particles are split into contiguous blocks, one per process / GPU.  The ring couples a particle with `radius = neighb_param / 2`
neighbours on each side, so before every generation a shard needs the best positions / fitness of `radius` particles beyond each of
its ends: the HALO.  All shards publish their first and last `radius` rows with ONE all_gather per generation (a few rows of nx + 1
doubles: 14 KB at 150 atoms) and pick their two neighbours' rows - NCCL over NVLink on GPUs, gloo in the CPU test.  Everything else
(velocity update, move, batch evaluation, memory update) is the device step `pgc_pso_shard_step_device`; draws are addressed by the
global particle index, so the sharded swarm moves exactly like `pgc_pso_evolve_device` on one GPU (checked bit for bit by
scripts/dist_swarm_gpu.py).  No CPU fallback: `DeviceShard` calls libpgc.so; the `shard` argument exists so that the exchange can be
exercised without a GPU."""
from __future__ import annotations

import numpy as np

from . import capi


def halo_rows(blocks: np.ndarray, rank: int, world: int, radius: int):
    """blocks[r] = rows published by shard r: its first `radius` rows followed by its last `radius` rows.  Returns (left halo, right
    halo) of shard `rank` on the ring: the last rows of its left neighbour, the first rows of its right neighbour."""
    left, right = (rank - 1) % world, (rank + 1) % world
    return blocks[left][radius:2 * radius], blocks[right][:radius]


class DeviceShard:
    """n_loc particles of the swarm on one GPU."""

    def __init__(self, ctx: capi.Context, prob: capi.Problem, x: np.ndarray, f: np.ndarray, index_offset: int, radius: int):
        self.ctx, self.prob, self.n, self.nx, self.radius, self.offset = ctx, prob, x.shape[0], x.shape[1], radius, index_offset
        ext = np.zeros((self.n + 2 * radius, self.nx))
        ext[radius:radius + self.n] = x
        fext = np.zeros(self.n + 2 * radius)
        fext[radius:radius + self.n] = np.asarray(f).reshape(-1)
        self.d_X, self.d_V = ctx.to_device(np.ascontiguousarray(x, dtype=np.float64)), ctx.malloc(8 * self.n * self.nx)
        self.d_lbX, self.d_lbf = ctx.to_device(ext), ctx.to_device(fext)

    def _rows(self, ptr: int, first_row: int, rows: int, width: int) -> np.ndarray:
        return self.ctx.from_device(ptr + 8 * first_row * width, (rows, width))

    def boundary(self) -> np.ndarray:
        """[2 radius x (nx + 1)]: best position | best fitness of the first and of the last `radius` particles."""
        r, n = self.radius, self.n
        x = np.vstack([self._rows(self.d_lbX, r, r, self.nx), self._rows(self.d_lbX, n, r, self.nx)])
        f = np.concatenate([self._rows(self.d_lbf, r, r, 1), self._rows(self.d_lbf, n, r, 1)])
        return np.hstack([x, f])

    def set_halos(self, left: np.ndarray, right: np.ndarray):
        r, n, nx = self.radius, self.n, self.nx
        lib = capi.lib()
        for rows, at in ((left, 0), (right, r + n)):
            x, f = np.ascontiguousarray(rows[:, :nx]), np.ascontiguousarray(rows[:, nx])
            capi.check(lib.pgc_memcpy_h2d(self.ctx._h, self.d_lbX + 8 * at * nx, x.ctypes.data, x.nbytes))
            capi.check(lib.pgc_memcpy_h2d(self.ctx._h, self.d_lbf + 8 * at, f.ctypes.data, f.nbytes))

    def step(self, p: dict, generation: int, init_velocity: bool = False):
        capi.check(capi.lib().pgc_pso_shard_step_device(self.prob._h, self.d_X, self.d_V, self.d_lbX, self.d_lbf, self.n, self.radius, self.offset,
                                                        p["omega"], p["eta1"], p["eta2"], p["max_vel"], p["variant"], p["seed"], generation,
                                                        int(init_velocity), None))

    def best(self):
        """(lbX, lbfit) of the shard's own particles: what pso_gen::evolve writes back into the population (pso_gen.cpp:524-527)."""
        return self._rows(self.d_lbX, self.radius, self.n, self.nx), self._rows(self.d_lbf, self.radius, self.n, 1)[:, 0]


class ShardedSwarm:
    """The shard of this process plus the per-generation halo exchange.  `shard` needs boundary() / set_halos() / step() / radius."""

    def __init__(self, shard, omega=0.7298, eta1=2.05, eta2=2.05, max_vel=0.5, variant=5, seed=0, first_generation=1, group=None):
        self.shard, self.group, self.generation = shard, group, first_generation
        self.params = dict(omega=omega, eta1=eta1, eta2=eta2, max_vel=max_vel, variant=variant, seed=seed)
        self.rank, self.world, self._dist = 0, 1, None
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self._dist, self.rank, self.world = dist, dist.get_rank(group), dist.get_world_size(group)
        except ImportError:
            pass
        shard.step(self.params, first_generation, init_velocity=True)  # velocities, pso_gen.cpp:187-196

    def exchange(self):
        mine = self.shard.boundary()
        if self.world == 1:
            blocks = mine[None]
        else:
            import torch
            dist = self._dist
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
            t = torch.from_numpy(mine).to(dev)
            parts = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(parts, t, group=self.group)
            blocks = np.stack([q.cpu().numpy() for q in parts])
        left, right = halo_rows(blocks, self.rank, self.world, self.shard.radius)
        self.shard.set_halos(left, right)

    def evolve(self, gens: int):
        for _ in range(gens):
            self.exchange()
            self.shard.step(self.params, self.generation)
            self.generation += 1


# ---- gbest topology (pso_gen.cpp:644-677; tracking rule :452-457) -----------------------------------------------------------------
# Every particle's best neighbour is the swarm's best particle.  A shard keeps that particle's best position in row 0 of its extended
# arrays; after every generation each shard publishes its candidate - the particle that improved to the smallest fitness, last index
# on ties - and all shards apply the reference's rule to the gathered candidates: smallest fitness, largest global index on ties,
# accepted if <= the current best.  One all_gather of nx + 2 doubles per shard and generation.

def pick_initial_best(cands: np.ndarray):
    """cands[r] = [fitness, global index, row...] of shard r's first minimum.  pop.best_idx(): the FIRST minimum of the swarm."""
    order = np.lexsort((cands[:, 1], cands[:, 0]))
    return cands[order[0]]


def pick_next_best(cands: np.ndarray, current_fit: float):
    """cands[r] = [fitness, global index or -1, row...].  Returns the winning candidate row or None (the best stays)."""
    live = cands[cands[:, 1] >= 0]
    if live.shape[0] == 0:
        return None
    order = np.lexsort((-live[:, 1], live[:, 0]))  # smallest fitness, LARGEST index on ties (the sequential scan keeps the last)
    best = live[order[0]]
    return best if best[0] <= current_fit else None


class GbestShard:
    """n_loc particles of a gbest swarm on one GPU: extended row 0 = the swarm's best, rows 1..n_loc = own particles."""
    radius = 1

    def __init__(self, ctx: capi.Context, prob: capi.Problem, x: np.ndarray, f: np.ndarray, index_offset: int):
        self.ctx, self.prob, self.n, self.nx, self.offset = ctx, prob, x.shape[0], x.shape[1], index_offset
        f = np.asarray(f, dtype=np.float64).reshape(-1)
        ext = np.zeros((self.n + 2, self.nx))
        ext[1:1 + self.n] = x
        fext = np.zeros(self.n + 2)
        fext[1:1 + self.n] = f
        self.d_X, self.d_V = ctx.to_device(np.ascontiguousarray(x, dtype=np.float64)), ctx.malloc(8 * self.n * self.nx)
        self.d_lbX, self.d_lbf, self.d_cand = ctx.to_device(ext), ctx.to_device(fext), ctx.malloc(16)
        i = int(np.argmin(f))  # first minimum
        self._initial = np.concatenate([[f[i], index_offset + i], x[i]])

    def initial_candidate(self) -> np.ndarray:
        return self._initial

    def candidate(self) -> np.ndarray:
        """[fitness, global index or -1, best position of that particle] after a step."""
        fit, idx = self.ctx.from_device(self.d_cand, (2,))
        if idx < 0:
            return np.concatenate([[np.inf, -1.0], np.zeros(self.nx)])
        row = self.ctx.from_device(self.d_lbX + 8 * (1 + int(idx)) * self.nx, (self.nx,))
        return np.concatenate([[fit, self.offset + idx], row])

    def set_best(self, cand: np.ndarray):
        x, f = np.ascontiguousarray(cand[2:]), np.ascontiguousarray(cand[:1])
        lib = capi.lib()
        capi.check(lib.pgc_memcpy_h2d(self.ctx._h, self.d_lbX, x.ctypes.data, x.nbytes))
        capi.check(lib.pgc_memcpy_h2d(self.ctx._h, self.d_lbf, f.ctypes.data, f.nbytes))

    def step(self, p: dict, generation: int, init_velocity: bool = False):
        capi.check(capi.lib().pgc_pso_shard_step_gbest_device(self.prob._h, self.d_X, self.d_V, self.d_lbX, self.d_lbf, self.n, self.offset,
                                                              p["omega"], p["eta1"], p["eta2"], p["max_vel"], p["variant"], p["seed"],
                                                              generation, int(init_velocity), self.d_cand, None))

    def best(self):
        rows = self.ctx.from_device(self.d_lbX + 8 * self.nx, (self.n, self.nx))
        return rows, self.ctx.from_device(self.d_lbf + 8, (self.n,))


class GbestSwarm:
    """The shard of this process plus the per-generation reduction of the shards' candidates.  `shard` needs initial_candidate() /
    candidate() / set_best() / step()."""

    def __init__(self, shard, omega=0.7298, eta1=2.05, eta2=2.05, max_vel=0.5, variant=5, seed=0, first_generation=1, group=None):
        self.shard, self.group, self.generation = shard, group, first_generation
        self.params = dict(omega=omega, eta1=eta1, eta2=eta2, max_vel=max_vel, variant=variant, seed=seed)
        self.rank, self.world, self._dist = 0, 1, None
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self._dist, self.rank, self.world = dist, dist.get_rank(group), dist.get_world_size(group)
        except ImportError:
            pass
        shard.step(self.params, first_generation, init_velocity=True)  # velocities, pso_gen.cpp:187-196
        self.current = pick_initial_best(self._gather(shard.initial_candidate()))  # best_fit / best_neighb, :215-224
        shard.set_best(self.current)

    def _gather(self, mine: np.ndarray) -> np.ndarray:
        if self.world == 1:
            return mine[None]
        import torch
        dist = self._dist
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
        t = torch.from_numpy(np.ascontiguousarray(mine)).to(dev)
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t, group=self.group)
        return np.stack([q.cpu().numpy() for q in parts])

    def evolve(self, gens: int):
        for _ in range(gens):
            self.shard.step(self.params, self.generation)
            winner = pick_next_best(self._gather(self.shard.candidate()), self.current[0])
            if winner is not None:
                self.current = winner
                self.shard.set_best(winner)
            self.generation += 1
