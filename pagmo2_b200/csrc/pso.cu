// pso.cu - generational particle swarm optimisation on a device-resident swarm.
//
// Reference: src/algorithms/pso_gen.cpp:120-530 (the synchronous variant; SURVEY.md F3: the asynchronous `pso` updates the
// best inside the particle loop and cannot be batched).  Per generation: best neighbour of every particle
// (particle__get_best_neighbor :593-623, lbest ring :679-698, gbest :644-664) -> velocity update (:231-327, variants 1-5)
// -> clamp, move, box correction (:329-363) -> batch evaluation (:417-440) -> memory update (:445-459).
// The arithmetic is the reference's expression by expression.  Draws: particle p owns the Philox substream
// (seed, kTagPso, generation, p); slot 2d / 2d+1 are r1 / r2 of coordinate d (variants 1, 5), slot d is r1 (variant 2),
// slots 0 / 1 are the per-particle r1 / r2 (variants 3, 4); initial velocities use (seed, kTagInit, generation, p, d).
// Topologies: 1 gbest, 2 lbest ring, 3 von Neumann lattice (:719-744), 4 adaptive random graph (:772-796: particle p informs itself and
// neighb_param - 1 particles drawn from (seed, kTagPsoTopology, generation, p, j); re-drawn after every generation that did not
// improve the swarm's best, :462).  FIPS (variant 6) is not on the device.
#include <cfloat>
#include <cmath>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "pgc_internal.cuh"
#include "philox.cuh"

namespace pgc
{

namespace
{

__device__ __forceinline__ bool less_f(double a, double b) { return !isnan(a) && (isnan(b) || a < b); } // custom_comparisons.hpp:54-72
__device__ __forceinline__ bool equal_f(double a, double b) { return (isnan(a) && isnan(b)) || a == b; } // :91-98
__device__ __forceinline__ bool leq_f(double a, double b) { return less_f(a, b) || equal_f(a, b); }

// lbest ring, pso_gen.cpp:679-698 + best neighbour :608-621 (ties: the LATER neighbour in the list wins)
__global__ void pso_lbest_kernel(const double *lbfit, unsigned n, unsigned radius, unsigned *bn)
{
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    unsigned best = 0;
    bool first = true;
    for (unsigned j = radius; j > 0u; --j) {
        const unsigned q = (p < j) ? p - j + n : p - j;
        if (first || leq_f(lbfit[q], lbfit[best])) best = q;
        first = false;
    }
    for (unsigned j = 1u; j <= radius; ++j) {
        const unsigned q = (p + j >= n) ? p + j - n : p + j;
        if (first || leq_f(lbfit[q], lbfit[best])) best = q;
        first = false;
    }
    bn[p] = best;
}

// von Neumann lattice, pso_gen.cpp:719-744: rows = the largest divisor of n not above sqrt(n); neighbours W, E, N, S with wrap-around,
// best neighbour :608-621 (a later neighbour wins ties)
__global__ void pso_von_kernel(const double *lbfit, unsigned n, int rows, int cols, unsigned *bn)
{
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int p_x = static_cast<int>(p) % cols, p_y = static_cast<int>(p) / cols;
    const int dx[4] = {-1, 1, 0, 0}, dy[4] = {0, 0, -1, 1};
    unsigned best = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int n_x = (p_x + dx[k]) % cols, n_y = (p_y + dy[k]) % rows;
        if (n_x < 0) n_x = cols + n_x;
        if (n_y < 0) n_y = rows + n_y;
        const unsigned q = static_cast<unsigned>(n_y * cols + n_x);
        if (k == 0 || leq_f(lbfit[q], lbfit[best])) best = q;
    }
    bn[p] = best;
}

// ---- adaptive random topology (:772-796) ------------------------------------------------------------------------------------
// targets[p * K + j], j >= 1: the particles p informs (j = 0 is p itself, implicit).  The reference stores, per particle q, the list
// of its informants in ascending order of the informant's index and picks the best with "a later entry wins ties" (:608-621), i.e.
// the informant of smallest fitness and, among equals, of LARGEST index - an order-free rule, evaluated here with two atomic passes.
__device__ __forceinline__ unsigned long long fit_key(double v) // order-preserving, every NaN last (less_than_f)
{
    unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
    b = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
    return (v != v) ? 0xffffffffffffffffull : b;
}

// re-draws the informants unless *keep != 0 (the swarm's best improved in the generation that just ended, :462)
__global__ void pso_rewire_kernel(unsigned *targets, unsigned n, unsigned K, unsigned long long seed, unsigned gen_key, const int *keep)
{
    if (keep && *keep) return;
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(n) * K) return;
    const unsigned p = static_cast<unsigned>(e / K), j = static_cast<unsigned>(e % K);
    unsigned t = p;
    if (j) {
        t = static_cast<unsigned>(philox_u01(seed, kTagPsoTopology, gen_key, p, j) * static_cast<double>(n));
        if (t >= n) t = n - 1u;
    }
    targets[e] = t;
}

__global__ void pso_ar_min_kernel(const double *lbfit, const unsigned *targets, unsigned n, unsigned K, unsigned long long *key)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(n) * K) return;
    atomicMin(&key[targets[e]], fit_key(lbfit[e / K]));
}

__global__ void pso_ar_arg_kernel(const double *lbfit, const unsigned *targets, unsigned n, unsigned K, const unsigned long long *key,
                                  unsigned *bn)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(n) * K) return;
    const unsigned p = static_cast<unsigned>(e / K), q = targets[e];
    if (fit_key(lbfit[p]) == key[q]) atomicMax(&bn[q], p);
}

struct MoveParams {
    double *X, *V;
    const double *lbX;
    const unsigned *bn;  // best neighbour per particle (lbest) or nullptr (gbest: *gbest)
    const unsigned *gbest;
    const double *lb, *ub;
    unsigned n, dim;
    double omega, eta1, eta2, max_vel;
    unsigned variant;
    unsigned long long seed;
    unsigned generation;
    // the fully informed swarm (variant 6) reads every neighbour: the topology itself instead of the best neighbour
    unsigned neighb_type, radius;
    int von_rows, von_cols;
    const unsigned *ar_off, *ar_list; // adaptive random: informants of q = ar_list[ar_off[q] .. ar_off[q + 1])
};

// k-th neighbour of particle p, in the order the reference's neighb[p] lists them (pso_gen.cpp:656-663 gbest + FIPS: everybody;
// :679-698 ring; :719-744 lattice; :772-796 random graph)
__device__ __forceinline__ unsigned pso_neighbour(const MoveParams &P, unsigned p, unsigned k)
{
    switch (P.neighb_type) {
        case 1: return k;
        case 2: {
            if (k < P.radius) {
                const unsigned j = P.radius - k;
                return (p < j) ? p - j + P.n : p - j;
            }
            const unsigned j = k - P.radius + 1u;
            return (p + j >= P.n) ? p + j - P.n : p + j;
        }
        case 3: {
            const int p_x = static_cast<int>(p) % P.von_cols, p_y = static_cast<int>(p) / P.von_cols;
            const int ddx = k == 0u ? -1 : (k == 1u ? 1 : 0), ddy = k == 2u ? -1 : (k == 3u ? 1 : 0);
            int n_x = (p_x + ddx) % P.von_cols, n_y = (p_y + ddy) % P.von_rows;
            if (n_x < 0) n_x = P.von_cols + n_x;
            if (n_y < 0) n_y = P.von_rows + n_y;
            return static_cast<unsigned>(n_y * P.von_cols + n_x);
        }
        default: return P.ar_list[P.ar_off[p] + k];
    }
}

// adaptive random graph as per-particle informant lists, in the order the reference appends them (ascending informant, its own
// entry at its turn): keys (q, p * K + j) sorted, then read back
__global__ void pso_ar_keys_kernel(const unsigned *targets, unsigned n, unsigned K, unsigned long long *keys)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(n) * K) return;
    keys[e] = static_cast<unsigned long long>(targets[e]) * (static_cast<unsigned long long>(n) * K) + e;
}

__global__ void pso_ar_csr_kernel(const unsigned long long *sorted, unsigned n, unsigned K, unsigned *off, unsigned *list)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t total = static_cast<size_t>(n) * K;
    if (e >= total) return;
    const unsigned q = static_cast<unsigned>(sorted[e] / total);
    list[e] = static_cast<unsigned>((sorted[e] % total) / K);
    if (e == 0 || static_cast<unsigned>(sorted[e - 1] / total) != q) off[q] = static_cast<unsigned>(e); // every q lists itself: no gaps
    if (e + 1 == total) off[n] = static_cast<unsigned>(total);
}

__global__ void pso_move_kernel(const MoveParams P)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(P.n) * P.dim) return;
    const unsigned p = static_cast<unsigned>(e / P.dim), d = static_cast<unsigned>(e % P.dim);
    const unsigned b = P.variant == 6u ? p : (P.bn ? P.bn[p] : *P.gbest);
    const double x = P.X[e], lbx = P.lbX[e], bnx = P.lbX[static_cast<size_t>(b) * P.dim + d];
    double v = P.V[e], r1, r2;
    switch (P.variant) { // pso_gen.cpp:242-326
        case 6: { // fully informed: one draw per (gene, neighbour), :318-326
            const unsigned K = P.neighb_type == 1u ? P.n : (P.neighb_type == 2u ? 2u * P.radius : (P.neighb_type == 3u ? 4u : P.ar_off[p + 1] - P.ar_off[p]));
            const double acceleration_coefficient = P.eta1 + P.eta2;
            double sum_forces = 0.;
            for (unsigned k = 0; k < K; ++k) {
                const unsigned q = pso_neighbour(P, p, k);
                const double u = philox_u01(P.seed, kTagPso, P.generation, p, d * K + k);
                sum_forces += u * acceleration_coefficient * (P.lbX[static_cast<size_t>(q) * P.dim + d] - x);
            }
            v = P.omega * (v + sum_forces / static_cast<double>(K));
            break;
        }
        case 1:
            r1 = philox_u01(P.seed, kTagPso, P.generation, p, 2 * d);
            r2 = philox_u01(P.seed, kTagPso, P.generation, p, 2 * d + 1);
            v = P.omega * v + P.eta1 * r1 * (lbx - x) + P.eta2 * r2 * (bnx - x);
            break;
        case 2:
            r1 = philox_u01(P.seed, kTagPso, P.generation, p, d);
            v = P.omega * v + P.eta1 * r1 * (lbx - x) + P.eta2 * r1 * (bnx - x);
            break;
        case 3:
            r1 = philox_u01(P.seed, kTagPso, P.generation, p, 0);
            r2 = philox_u01(P.seed, kTagPso, P.generation, p, 1);
            v = P.omega * v + P.eta1 * r1 * (lbx - x) + P.eta2 * r2 * (bnx - x);
            break;
        case 4:
            r1 = philox_u01(P.seed, kTagPso, P.generation, p, 0);
            v = P.omega * v + P.eta1 * r1 * (lbx - x) + P.eta2 * r1 * (bnx - x);
            break;
        default: // 5
            r1 = philox_u01(P.seed, kTagPso, P.generation, p, 2 * d);
            r2 = philox_u01(P.seed, kTagPso, P.generation, p, 2 * d + 1);
            v = P.omega * (v + P.eta1 * r1 * (lbx - x) + P.eta2 * r2 * (bnx - x));
    }
    // :329-363
    const double vwidth = (P.ub[d] - P.lb[d]) * P.max_vel, minv = -1. * vwidth, maxv = vwidth;
    if (v > maxv) v = maxv;
    else if (v < minv) v = minv;
    double new_x = x + v;
    if (new_x < P.lb[d]) {
        new_x = P.lb[d];
        v = 0.;
    } else if (new_x > P.ub[d]) {
        new_x = P.ub[d];
        v = 0.;
    }
    P.X[e] = new_x;
    P.V[e] = v;
}

__global__ void pso_init_velocity_kernel(double *V, const double *lb, const double *ub, unsigned n, unsigned dim, double max_vel,
                                         unsigned long long seed, unsigned generation)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(n) * dim) return;
    const unsigned p = static_cast<unsigned>(e / dim), d = static_cast<unsigned>(e % dim);
    const double vwidth = (ub[d] - lb[d]) * max_vel, minv = -1. * vwidth, maxv = vwidth; // :179-183
    const double u = philox_u01(seed, kTagInit, generation, p, d);
    V[e] = (minv == maxv) ? minv : (maxv - minv) * u + minv; // uniform_real_from_range, :190-195
}

// memory update, :445-459: improved[p] = fit <= lbfit
__global__ void pso_memory_flag_kernel(const double *fit, double *lbfit, unsigned n, unsigned char *improved)
{
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const bool imp = leq_f(fit[p], lbfit[p]);
    improved[p] = imp;
    if (imp) lbfit[p] = fit[p];
}

__global__ void pso_memory_copy_kernel(const double *X, double *lbX, const unsigned char *improved, unsigned n, unsigned dim)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(n) * dim) return;
    if (improved[e / dim]) lbX[e] = X[e];
}

// gbest tracking (:452-457): sequential "if improved and fit <= best: best = p" over ascending p == among the improved
// particles the smallest fitness, the LAST index on ties, accepted if <= the previous best.  Single CTA.
__global__ void pso_gbest_kernel(const double *fit, const unsigned char *improved, unsigned n, unsigned *gbest, double *gbest_fit, int init,
                                 int *best_improved = nullptr)
{
    __shared__ double sf[256];
    __shared__ unsigned si[256];
    double bf = 0.;
    unsigned bi = 0xffffffffu;
    for (unsigned p = threadIdx.x; p < n; p += blockDim.x) {
        if (!init && !improved[p]) continue;
        const double f = fit[p];
        // init: pop.best_idx() = first minimum; update: last index among equal minima
        if (bi == 0xffffffffu || less_f(f, bf) || (!init && equal_f(f, bf))) {
            bf = f;
            bi = p;
        }
    }
    sf[threadIdx.x] = bf;
    si[threadIdx.x] = bi;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (unsigned t = 1; t < blockDim.x; ++t) {
            if (si[t] == 0xffffffffu) continue;
            const bool better = bi == 0xffffffffu || less_f(sf[t], bf) || (equal_f(sf[t], bf) && (init ? si[t] < bi : si[t] > bi));
            if (better) {
                bf = sf[t];
                bi = si[t];
            }
        }
        const bool take = bi != 0xffffffffu && (init || leq_f(bf, *gbest_fit));
        if (take) {
            *gbest = bi;
            *gbest_fit = bf;
        }
        if (best_improved) *best_improved = take ? 1 : 0; // best_fit_improved, :447-459
    }
}

// ---- one swarm sharded over several GPUs (contiguous blocks of particles) ---------------------------------------------------
// The lbest ring couples a particle with `radius` neighbours on each side, so a shard needs the best positions / fitness of
// `radius` particles beyond each of its ends.  They live in HALO rows of extended arrays: lbX_ext [(n_loc + 2 radius) x dim],
// lbfit_ext [n_loc + 2 radius], the shard's own particles at rows [radius, radius + n_loc); the caller fills the halos (from
// the neighbouring shards, or from its own other end when there is one shard) before every step.  Neighbour indices are in
// extended coordinates, no wrap-around; Philox substreams are addressed by the GLOBAL particle index, so a sharded swarm
// moves exactly like the same swarm on one device.
__global__ void pso_lbest_ext_kernel(const double *lbfit_ext, unsigned n_loc, unsigned radius, unsigned *bn)
{
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_loc) return;
    const unsigned c = p + radius; // own row in extended coordinates
    unsigned best = 0;
    bool first = true;
    for (unsigned j = radius; j > 0u; --j) { // same visiting order as pso_lbest_kernel: p-radius .. p-1, p+1 .. p+radius
        const unsigned q = c - j;
        if (first || leq_f(lbfit_ext[q], lbfit_ext[best])) best = q;
        first = false;
    }
    for (unsigned j = 1u; j <= radius; ++j) {
        const unsigned q = c + j;
        if (first || leq_f(lbfit_ext[q], lbfit_ext[best])) best = q;
        first = false;
    }
    bn[p] = best;
}

struct ShardMoveParams {
    double *X, *V;           // [n_loc x dim]
    const double *lbX_ext;   // [(n_loc + 2 radius) x dim]
    const unsigned *bn;      // best neighbour, extended row index
    const double *lb, *ub;
    unsigned n_loc, dim, radius, index_offset;
    double omega, eta1, eta2, max_vel;
    unsigned variant;
    unsigned long long seed;
    unsigned generation;
};

__global__ void pso_move_shard_kernel(const ShardMoveParams P)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(P.n_loc) * P.dim) return;
    const unsigned p = static_cast<unsigned>(e / P.dim), d = static_cast<unsigned>(e % P.dim);
    const unsigned gp = P.index_offset + p; // global particle index: the draw address
    const double x = P.X[e], lbx = P.lbX_ext[static_cast<size_t>(p + P.radius) * P.dim + d], bnx = P.lbX_ext[static_cast<size_t>(P.bn[p]) * P.dim + d];
    double v = P.V[e], r1, r2;
    switch (P.variant) { // pso_gen.cpp:242-306, as pso_move_kernel
        case 1:
            r1 = philox_u01(P.seed, kTagPso, P.generation, gp, 2 * d);
            r2 = philox_u01(P.seed, kTagPso, P.generation, gp, 2 * d + 1);
            v = P.omega * v + P.eta1 * r1 * (lbx - x) + P.eta2 * r2 * (bnx - x);
            break;
        case 2:
            r1 = philox_u01(P.seed, kTagPso, P.generation, gp, d);
            v = P.omega * v + P.eta1 * r1 * (lbx - x) + P.eta2 * r1 * (bnx - x);
            break;
        case 3:
            r1 = philox_u01(P.seed, kTagPso, P.generation, gp, 0);
            r2 = philox_u01(P.seed, kTagPso, P.generation, gp, 1);
            v = P.omega * v + P.eta1 * r1 * (lbx - x) + P.eta2 * r2 * (bnx - x);
            break;
        case 4:
            r1 = philox_u01(P.seed, kTagPso, P.generation, gp, 0);
            v = P.omega * v + P.eta1 * r1 * (lbx - x) + P.eta2 * r1 * (bnx - x);
            break;
        default: // 5
            r1 = philox_u01(P.seed, kTagPso, P.generation, gp, 2 * d);
            r2 = philox_u01(P.seed, kTagPso, P.generation, gp, 2 * d + 1);
            v = P.omega * (v + P.eta1 * r1 * (lbx - x) + P.eta2 * r2 * (bnx - x));
    }
    const double vwidth = (P.ub[d] - P.lb[d]) * P.max_vel, minv = -1. * vwidth, maxv = vwidth; // :329-363
    if (v > maxv) v = maxv;
    else if (v < minv) v = minv;
    double new_x = x + v;
    if (new_x < P.lb[d]) {
        new_x = P.lb[d];
        v = 0.;
    } else if (new_x > P.ub[d]) {
        new_x = P.ub[d];
        v = 0.;
    }
    P.X[e] = new_x;
    P.V[e] = v;
}

__global__ void pso_init_velocity_shard_kernel(double *V, const double *lb, const double *ub, unsigned n_loc, unsigned dim, unsigned index_offset,
                                               double max_vel, unsigned long long seed, unsigned generation)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(n_loc) * dim) return;
    const unsigned p = static_cast<unsigned>(e / dim), d = static_cast<unsigned>(e % dim);
    const double vwidth = (ub[d] - lb[d]) * max_vel, minv = -1. * vwidth, maxv = vwidth; // :179-183
    const double u = philox_u01(seed, kTagInit, generation, index_offset + p, d);
    V[e] = (minv == maxv) ? minv : (maxv - minv) * u + minv;
}

// gbest topology on a shard: every particle's best neighbour is the swarm's best particle, whose row the caller keeps in
// extended row 0
__global__ void pso_fill_u32_kernel(unsigned *p, unsigned n, unsigned v)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// the shard's candidate for the swarm's new best (:452-457 restricted to this shard): among the particles that improved in
// this generation the smallest fitness, the LAST index on ties.  cand[0] = fitness (+inf if none), cand[1] = local index (-1).
__global__ void pso_shard_candidate_kernel(const double *fit, const unsigned char *improved, unsigned n, double *cand)
{
    __shared__ double sf[256];
    __shared__ unsigned si[256];
    double bf = 0.;
    unsigned bi = 0xffffffffu;
    for (unsigned p = threadIdx.x; p < n; p += blockDim.x) {
        if (!improved[p]) continue;
        const double f = fit[p];
        if (bi == 0xffffffffu || less_f(f, bf) || equal_f(f, bf)) {
            bf = f;
            bi = p;
        }
    }
    sf[threadIdx.x] = bf;
    si[threadIdx.x] = bi;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (unsigned t = 1; t < blockDim.x; ++t) {
            if (si[t] == 0xffffffffu) continue;
            if (bi == 0xffffffffu || less_f(sf[t], bf) || (equal_f(sf[t], bf) && si[t] > bi)) {
                bf = sf[t];
                bi = si[t];
            }
        }
        cand[0] = bi == 0xffffffffu ? INFINITY : bf;
        cand[1] = bi == 0xffffffffu ? -1.0 : static_cast<double>(bi);
    }
}

inline unsigned nblk(size_t n, unsigned t) { return static_cast<unsigned>((n + t - 1) / t); }

} // namespace

// One generation of a shard.  d_cand == nullptr: lbest ring; d_cand != nullptr: gbest topology (radius must be 1, extended row 0 =
// the swarm's best row, d_cand receives the shard's candidate for the next best).  d_V == nullptr on entry is not allowed: call
// with init_velocity = 1 once to draw them.
int pso_shard_step_device(pgc_problem *prob, double *d_X, double *d_V, double *d_lbX_ext, double *d_lbfit_ext, unsigned n_loc, unsigned radius,
                          unsigned index_offset, double omega, double eta1, double eta2, double max_vel, unsigned variant,
                          unsigned long long seed, unsigned generation, int init_velocity,
                          int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st, double *d_cand)
{
    pgc_ctx *ctx = prob->ctx;
    const unsigned dim = static_cast<unsigned>(prob->nx);
    PGC_REQUIRE(omega >= 0. && omega <= 1., "The particles' inertia (or the constriction factor) must be in the [0,1] range, while a value of %g was detected", omega);
    PGC_REQUIRE(eta1 >= 0. && eta2 >= 0. && eta1 <= 4. && eta2 <= 4., "The eta parameters must be in the [0,4] range, while eta1 = %g, eta2 = %g was detected", eta1, eta2);
    PGC_REQUIRE(max_vel > 0. && max_vel <= 1., "The maximum particle velocity (as a fraction of the bounds) should be in the (0,1] range, while a value of %g was detected", max_vel);
    PGC_REQUIRE(variant >= 1u && variant <= 5u, "pso shards implement variants 1-5, while a value of %u was detected", variant);
    PGC_REQUIRE(prob->nobj == 1, "Multiple objectives detected in %s instance. PSO cannot deal with them", prob->name.c_str());
    PGC_REQUIRE(n_loc >= 1 && radius >= 1, "pso shard: empty shard or zero radius");
    PGC_REQUIRE(!d_cand || radius == 1, "pso shard, gbest topology: the extended arrays carry one row at each end (radius 1), got %u", radius);
    const size_t nd = static_cast<size_t>(n_loc) * dim;
    double *lb = nullptr, *fit = nullptr;
    unsigned *bn = nullptr;
    unsigned char *improved = nullptr;
    StreamScratch scratch(st); // released in stream order on every path out of this function
    PGC_CUDA(scratch.get(&lb, 16 * dim));
    PGC_CUDA(scratch.get(&fit, 8 * static_cast<size_t>(n_loc)));
    PGC_CUDA(scratch.get(&bn, 4 * static_cast<size_t>(n_loc)));
    PGC_CUDA(scratch.get(&improved, n_loc));
    double *ub = lb + dim;
    PGC_CUDA(cudaMemcpyAsync(lb, prob->lb.data(), 8 * dim, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(ub, prob->ub.data(), 8 * dim, cudaMemcpyHostToDevice, st));
    int rc = PGC_OK;
    if (init_velocity) {
        pso_init_velocity_shard_kernel<<<nblk(nd, 256), 256, 0, st>>>(d_V, lb, ub, n_loc, dim, index_offset, max_vel, seed, generation);
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
    } else {
        if (d_cand) pso_fill_u32_kernel<<<nblk(n_loc, 256), 256, 0, st>>>(bn, n_loc, 0u);
        else pso_lbest_ext_kernel<<<nblk(n_loc, 256), 256, 0, st>>>(d_lbfit_ext, n_loc, radius, bn);
        ShardMoveParams mp{d_X, d_V, d_lbX_ext, bn, lb, ub, n_loc, dim, radius, index_offset, omega, eta1, eta2, max_vel, variant, seed, generation};
        pso_move_shard_kernel<<<nblk(nd, 256), 256, 0, st>>>(mp);
        rc = eval(prob, d_X, n_loc, fit, st);
        if (rc == PGC_OK) {
            pso_memory_flag_kernel<<<nblk(n_loc, 256), 256, 0, st>>>(fit, d_lbfit_ext + radius, n_loc, improved);
            pso_memory_copy_kernel<<<nblk(nd, 256), 256, 0, st>>>(d_X, d_lbX_ext + static_cast<size_t>(radius) * dim, improved, n_loc, dim);
            if (d_cand) pso_shard_candidate_kernel<<<1, 256, 0, st>>>(fit, improved, n_loc, d_cand);
            ctx->launches.fetch_add(d_cand ? 5 : 4, std::memory_order_relaxed);
        }
    }
    cudaError_t e = cudaStreamSynchronize(st); // lb / ub came from pageable host vectors
    if (rc != PGC_OK) return rc;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "pso_shard_step_device", __FILE__, __LINE__);
    return PGC_OK;
}

// pso_gen::evolve on a device-resident swarm.  In: d_x = positions, d_f = their fitness, d_v = velocities (or nullptr: drawn
// as in :187-196).  Out: d_x / d_f = the particles' best positions lbX / lbfit (what evolve() puts back into the population,
// :524-527); d_v (if given) = the final velocities, d_xcur (if given) = the final current positions.
// the velocities a first evolve() with memory draws (pso_gen.cpp:193-201)
int pso_init_velocity_device(pgc_problem *prob, unsigned n, double max_vel, unsigned long long seed, unsigned generation, double *d_v,
                             cudaStream_t st)
{
    const unsigned dim = static_cast<unsigned>(prob->nx);
    PGC_REQUIRE(d_v, "pso velocities: null array");
    PGC_REQUIRE(max_vel > 0. && max_vel <= 1., "The maximum particle velocity (as a fraction of the bounds) should be in the (0,1] range, while a value of %g was detected", max_vel);
    StreamScratch scratch(st);
    double *lb = nullptr;
    PGC_CUDA(scratch.get(&lb, 16 * static_cast<size_t>(dim)));
    PGC_CUDA(cudaMemcpyAsync(lb, prob->lb.data(), 8 * dim, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(lb + dim, prob->ub.data(), 8 * dim, cudaMemcpyHostToDevice, st));
    pso_init_velocity_kernel<<<nblk(static_cast<size_t>(n) * dim, 256), 256, 0, st>>>(d_v, lb, lb + dim, n, dim, max_vel, seed, generation);
    PGC_CUDA(cudaGetLastError());
    PGC_CUDA(cudaStreamSynchronize(st));
    return PGC_OK;
}

int pso_evolve_device(pgc_problem *prob, double *d_x, double *d_f, double *d_v, double *d_xcur, unsigned n, unsigned gens, double omega,
                      double eta1, double eta2, double max_vel, unsigned variant, unsigned neighb_type, unsigned neighb_param,
                      unsigned long long seed, unsigned first_generation,
                      int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st)
{
    pgc_ctx *ctx = prob->ctx;
    const unsigned dim = static_cast<unsigned>(prob->nx);
    // constructor / evolve checks, pso_gen.cpp:79-109,133-145
    PGC_REQUIRE(omega >= 0. && omega <= 1., "The particles' inertia (or the constriction factor) must be in the [0,1] range, while a value of %g was detected", omega);
    PGC_REQUIRE(eta1 >= 0. && eta2 >= 0. && eta1 <= 4. && eta2 <= 4., "The eta parameters must be in the [0,4] range, while eta1 = %g, eta2 = %g was detected", eta1, eta2);
    PGC_REQUIRE(max_vel > 0. && max_vel <= 1., "The maximum particle velocity (as a fraction of the bounds) should be in the (0,1] range, while a value of %g was detected", max_vel);
    PGC_REQUIRE(variant >= 1u && variant <= 6u, "The PSO variant must be in [1,6], while a value of %u was detected", variant);
    PGC_REQUIRE(neighb_type >= 1u && neighb_type <= 4u, "The swarm topology variant must be in [1,4], while a value of %u was detected", neighb_type);
    PGC_REQUIRE(neighb_param >= 1u, "The neighborhood parameter must be in (0, inf), while a value of %u was detected", neighb_param);
    PGC_REQUIRE(prob->nobj == 1, "Multiple objectives detected in %s instance. PSO cannot deal with them", prob->name.c_str());
    PGC_REQUIRE(n > 0, "PSO does not work on an empty population");
    struct Buf {
        cudaStream_t st;
        std::vector<void *> owned;
        ~Buf()
        {
            for (void *p : owned) cudaFreeAsync(p, st);
        }
        int get(void **out, size_t bytes)
        {
            PGC_CUDA(cudaMallocAsync(out, bytes ? bytes : 1, st));
            owned.push_back(*out);
            return PGC_OK;
        }
    } buf{st, {}};
    double *X, *V, *fit, *lb, *ub, *gfit;
    unsigned *bn, *gbest, *targets = nullptr, *ar_off = nullptr, *ar_list = nullptr;
    unsigned long long *arkey = nullptr, *ar_k0 = nullptr, *ar_k1 = nullptr;
    void *ar_ws = nullptr;
    size_t ar_ws_bytes = 0;
    int *keep = nullptr;
    unsigned char *improved;
    const size_t nd = static_cast<size_t>(n) * dim;
    const unsigned K = neighb_param; // adaptive random: out-degree including the particle itself
    int rc;
    if (neighb_type == 4u
        && ((rc = buf.get(reinterpret_cast<void **>(&targets), 4 * static_cast<size_t>(n) * K)) || (rc = buf.get(reinterpret_cast<void **>(&arkey), 8 * static_cast<size_t>(n)))
            || (rc = buf.get(reinterpret_cast<void **>(&keep), 4))))
        return rc;
    const bool fips_graph = variant == 6u && neighb_type == 4u; // the informant lists themselves are needed, not only the best informant
    if (fips_graph) {
        const size_t total = static_cast<size_t>(n) * K;
        PGC_REQUIRE(total < (1ull << 31), "pso: swarm size x neighb_param too large for the fully informed random graph");
        PGC_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, ar_ws_bytes, ar_k0, ar_k1, static_cast<int>(total), 0, 64, st));
        if ((rc = buf.get(reinterpret_cast<void **>(&ar_off), 4 * (static_cast<size_t>(n) + 1))) || (rc = buf.get(reinterpret_cast<void **>(&ar_list), 4 * total))
            || (rc = buf.get(reinterpret_cast<void **>(&ar_k0), 8 * total)) || (rc = buf.get(reinterpret_cast<void **>(&ar_k1), 8 * total))
            || (rc = buf.get(&ar_ws, ar_ws_bytes)))
            return rc;
    }
    if ((rc = buf.get(reinterpret_cast<void **>(&X), 8 * nd)) || (rc = buf.get(reinterpret_cast<void **>(&V), 8 * nd))
        || (rc = buf.get(reinterpret_cast<void **>(&fit), 8 * n)) || (rc = buf.get(reinterpret_cast<void **>(&lb), 8 * dim))
        || (rc = buf.get(reinterpret_cast<void **>(&ub), 8 * dim)) || (rc = buf.get(reinterpret_cast<void **>(&gfit), 8))
        || (rc = buf.get(reinterpret_cast<void **>(&bn), 4 * n)) || (rc = buf.get(reinterpret_cast<void **>(&gbest), 4))
        || (rc = buf.get(reinterpret_cast<void **>(&improved), n)))
        return rc;
    PGC_CUDA(cudaMemcpyAsync(lb, prob->lb.data(), 8 * dim, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(ub, prob->ub.data(), 8 * dim, cudaMemcpyHostToDevice, st));
    // X = lbX = pop.x, fit = lbfit = pop.f (:186-192): d_x / d_f play the role of lbX / lbfit from here on
    PGC_CUDA(cudaMemcpyAsync(X, d_x, 8 * nd, cudaMemcpyDeviceToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(fit, d_f, 8 * n, cudaMemcpyDeviceToDevice, st));
    if (d_v) PGC_CUDA(cudaMemcpyAsync(V, d_v, 8 * nd, cudaMemcpyDeviceToDevice, st));
    else pso_init_velocity_kernel<<<nblk(nd, 256), 256, 0, st>>>(V, lb, ub, n, dim, max_vel, seed, first_generation);
    if (neighb_type == 1u || neighb_type == 4u) pso_gbest_kernel<<<1, 256, 0, st>>>(d_f, nullptr, n, gbest, gfit, 1);
    int von_rows = 1, von_cols = 1;
    if (neighb_type == 3u) { // :724-727
        von_rows = static_cast<int>(std::sqrt(static_cast<double>(n)));
        while (static_cast<int>(n) % von_rows != 0) von_rows -= 1;
        von_cols = static_cast<int>(n) / von_rows;
    }
    if (neighb_type == 4u) pso_rewire_kernel<<<nblk(static_cast<size_t>(n) * K, 256), 256, 0, st>>>(targets, n, K, seed, first_generation, nullptr);
    const unsigned radius = neighb_param / 2u;
    PGC_REQUIRE(neighb_type != 2u || (radius >= 1u && 2u * radius < n), "lbest topology: neighb_param / 2 = %u must be in [1, (swarm size - 1) / 2]", radius);
    for (unsigned g = 0; g < gens; ++g) {
        const unsigned generation = first_generation + g;
        if (variant == 6u) { // no best neighbour to find (:236-238)
            if (fips_graph) {
                const size_t total = static_cast<size_t>(n) * K;
                pso_ar_keys_kernel<<<nblk(total, 256), 256, 0, st>>>(targets, n, K, ar_k0);
                PGC_CUDA(cub::DeviceRadixSort::SortKeys(ar_ws, ar_ws_bytes, ar_k0, ar_k1, static_cast<int>(total), 0, 64, st));
                pso_ar_csr_kernel<<<nblk(total, 256), 256, 0, st>>>(ar_k1, n, K, ar_off, ar_list);
            }
        } else if (neighb_type == 2u) pso_lbest_kernel<<<nblk(n, 256), 256, 0, st>>>(d_f, n, radius, bn);
        else if (neighb_type == 3u) pso_von_kernel<<<nblk(n, 256), 256, 0, st>>>(d_f, n, von_rows, von_cols, bn);
        else if (neighb_type == 4u) {
            PGC_CUDA(cudaMemsetAsync(arkey, 0xff, 8 * static_cast<size_t>(n), st));
            PGC_CUDA(cudaMemsetAsync(bn, 0, 4 * static_cast<size_t>(n), st));
            pso_ar_min_kernel<<<nblk(static_cast<size_t>(n) * K, 256), 256, 0, st>>>(d_f, targets, n, K, arkey);
            pso_ar_arg_kernel<<<nblk(static_cast<size_t>(n) * K, 256), 256, 0, st>>>(d_f, targets, n, K, arkey, bn);
        }
        MoveParams mp{X, V, d_x, neighb_type != 1u ? bn : nullptr, gbest, lb, ub, n, dim, omega, eta1, eta2, max_vel, variant, seed, generation,
                      neighb_type, radius, von_rows, von_cols, ar_off, ar_list};
        pso_move_kernel<<<nblk(nd, 256), 256, 0, st>>>(mp);
        if ((rc = eval(prob, X, n, fit, st))) return rc;
        pso_memory_flag_kernel<<<nblk(n, 256), 256, 0, st>>>(fit, d_f, n, improved);
        pso_memory_copy_kernel<<<nblk(nd, 256), 256, 0, st>>>(X, d_x, improved, n, dim);
        if (neighb_type == 1u) pso_gbest_kernel<<<1, 256, 0, st>>>(fit, improved, n, gbest, gfit, 0);
        if (neighb_type == 4u) { // the graph is re-drawn when the swarm's best did not improve in this generation (:462)
            pso_gbest_kernel<<<1, 256, 0, st>>>(fit, improved, n, gbest, gfit, 0, keep);
            pso_rewire_kernel<<<nblk(static_cast<size_t>(n) * K, 256), 256, 0, st>>>(targets, n, K, seed, generation + 1u, keep);
        }
        ctx->launches.fetch_add(neighb_type == 4u ? 8 : 4, std::memory_order_relaxed);
        if (log_due(g + 1u) && (rc = log_pso_device(ctx, X, V, d_f, lb, ub, n, dim, g + 1u, static_cast<unsigned long long>(g + 1u) * n, st))) return rc; // pso_gen.cpp:464-518
    }
    if (d_v) PGC_CUDA(cudaMemcpyAsync(d_v, V, 8 * nd, cudaMemcpyDeviceToDevice, st));
    if (d_xcur) PGC_CUDA(cudaMemcpyAsync(d_xcur, X, 8 * nd, cudaMemcpyDeviceToDevice, st));
    PGC_CUDA(cudaGetLastError());
    PGC_CUDA(cudaStreamSynchronize(st));
    return PGC_OK;
}

} // namespace pgc
