// pso.cu - generational particle swarm optimisation on a device-resident swarm.
//
// Reference: src/algorithms/pso_gen.cpp:120-530 (the synchronous variant; SURVEY.md F3: the asynchronous `pso` updates the
// best inside the particle loop and cannot be batched).  Per generation: best neighbour of every particle
// (particle__get_best_neighbor :593-623, lbest ring :679-698, gbest :644-664) -> velocity update (:231-327, variants 1-5)
// -> clamp, move, box correction (:329-363) -> batch evaluation (:417-440) -> memory update (:445-459).
// The arithmetic is the reference's expression by expression.  Draws: particle p owns the Philox substream
// (seed, kTagPso, generation, p); slot 2d / 2d+1 are r1 / r2 of coordinate d (variants 1, 5), slot d is r1 (variant 2),
// slots 0 / 1 are the per-particle r1 / r2 (variants 3, 4); initial velocities use (seed, kTagInit, generation, p, d).
// FIPS (variant 6) and the von-Neumann / adaptive-random topologies (3, 4) are not on the device.
#include <cfloat>
#include <cmath>
#include <vector>

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
};

__global__ void pso_move_kernel(const MoveParams P)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(P.n) * P.dim) return;
    const unsigned p = static_cast<unsigned>(e / P.dim), d = static_cast<unsigned>(e % P.dim);
    const unsigned b = P.bn ? P.bn[p] : *P.gbest;
    const double x = P.X[e], lbx = P.lbX[e], bnx = P.lbX[static_cast<size_t>(b) * P.dim + d];
    double v = P.V[e], r1, r2;
    switch (P.variant) { // pso_gen.cpp:242-306
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
__global__ void pso_gbest_kernel(const double *fit, const unsigned char *improved, unsigned n, unsigned *gbest, double *gbest_fit, int init)
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
        if (bi != 0xffffffffu && (init || leq_f(bf, *gbest_fit))) {
            *gbest = bi;
            *gbest_fit = bf;
        }
    }
}

inline unsigned nblk(size_t n, unsigned t) { return static_cast<unsigned>((n + t - 1) / t); }

} // namespace

// pso_gen::evolve on a device-resident swarm.  In: d_x = positions, d_f = their fitness, d_v = velocities (or nullptr: drawn
// as in :187-196).  Out: d_x / d_f = the particles' best positions lbX / lbfit (what evolve() puts back into the population,
// :524-527); d_v (if given) = the final velocities, d_xcur (if given) = the final current positions.
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
    if (variant == 6u || neighb_type > 2u) {
        set_error("pso on the device implements variants 1-5 and the gbest / lbest topologies (variant %u, topology %u requested)", variant, neighb_type);
        return PGC_ERR_UNSUPPORTED;
    }
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
    unsigned *bn, *gbest;
    unsigned char *improved;
    const size_t nd = static_cast<size_t>(n) * dim;
    int rc;
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
    if (neighb_type == 1u) pso_gbest_kernel<<<1, 256, 0, st>>>(d_f, nullptr, n, gbest, gfit, 1);
    const unsigned radius = neighb_param / 2u;
    PGC_REQUIRE(neighb_type != 2u || (radius >= 1u && 2u * radius < n), "lbest topology: neighb_param / 2 = %u must be in [1, (swarm size - 1) / 2]", radius);
    for (unsigned g = 0; g < gens; ++g) {
        const unsigned generation = first_generation + g;
        if (neighb_type == 2u) pso_lbest_kernel<<<nblk(n, 256), 256, 0, st>>>(d_f, n, radius, bn);
        MoveParams mp{X, V, d_x, neighb_type == 2u ? bn : nullptr, gbest, lb, ub, n, dim, omega, eta1, eta2, max_vel, variant, seed, generation};
        pso_move_kernel<<<nblk(nd, 256), 256, 0, st>>>(mp);
        if ((rc = eval(prob, X, n, fit, st))) return rc;
        pso_memory_flag_kernel<<<nblk(n, 256), 256, 0, st>>>(fit, d_f, n, improved);
        pso_memory_copy_kernel<<<nblk(nd, 256), 256, 0, st>>>(X, d_x, improved, n, dim);
        if (neighb_type == 1u) pso_gbest_kernel<<<1, 256, 0, st>>>(fit, improved, n, gbest, gfit, 0);
        ctx->launches.fetch_add(neighb_type == 1u ? 4 : 4, std::memory_order_relaxed);
    }
    if (d_v) PGC_CUDA(cudaMemcpyAsync(d_v, V, 8 * nd, cudaMemcpyDeviceToDevice, st));
    if (d_xcur) PGC_CUDA(cudaMemcpyAsync(d_xcur, X, 8 * nd, cudaMemcpyDeviceToDevice, st));
    PGC_CUDA(cudaGetLastError());
    PGC_CUDA(cudaStreamSynchronize(st));
    return PGC_OK;
}

} // namespace pgc
