// nsga2.cu - NSGA-II generation operators and the device-resident generation loop.
//
// Reference: src/algorithms/nsga2.cpp:91-307 (evolve), src/utils/genetic_operators.cpp:49-57 (sbx_betaq), :71-144
// (sbx_crossover_impl), :148-197 (polynomial_mutation_impl), :200-211 (mo_tournament_selection_impl).
//
// One generation (nsga2.cpp:176-304, the bfe branch):
//   shuffle1, shuffle2            two permutations of the population (:180-181)
//   FNDS + crowding distance      mo_utils.cu, nsga2's small-front rule (:184-206)
//   per group of 4 individuals    T,T,SBX,PM,PM on shuffle1 then T,T,SBX,PM,PM on shuffle2 -> 4 children (:215-239)
//   batch evaluation of children  the problem's device evaluator (:253)
//   select_best_N_mo on 2N        mo_utils.cu (:300), survivors gathered in that order (:302-304)
// The operators are the reference's, statement by statement; only the source of randomness changes: every group of 4
// owns the Philox substream (seed, kTagNsga2Var, generation, group) and consumes draws in the reference's order (the
// draw is the first operand of every `&&`, genetic_operators.cpp:94,165), and a permutation is the stable argsort of N
// Philox keys instead of std::shuffle on mt19937 (libstdc++-specific, not reproducible elsewhere: SURVEY.md App. C).
// Integer decision variables (nix > 0, the two-point crossover / integer mutation tails) are not handled on the device.
#include <cub/device/device_radix_sort.cuh>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pgc_internal.cuh"
#include "philox.cuh"

namespace pgc
{

namespace
{

__device__ __forceinline__ double sbx_betaq(double beta, double eta_c, double rand01) // genetic_operators.cpp:49-57
{
    const double alpha = 2. - pow(beta, -(eta_c + 1.));
    if (rand01 < 1. / alpha) return pow(rand01 * alpha, 1. / (eta_c + 1.));
    return pow(1. / (2. - rand01 * alpha), 1. / (eta_c + 1.));
}

struct VarParams {
    const double *x;       // [NP x nx] parents
    const unsigned *rank;  // non-domination rank
    const double *cd;      // crowding distance
    const unsigned *sh1, *sh2;
    const double *lb, *ub; // [nx]
    double *children;      // [NP x nx]
    unsigned NP, nx, nix; // nix: integer alleles at the end of the chromosome (problem::get_nix())
    double cr, eta_c, m, eta_m;
    unsigned long long seed;
    unsigned generation;
};

__device__ __forceinline__ unsigned tournament(unsigned i1, unsigned i2, const VarParams &P, PhiloxStream &rs)
{ // genetic_operators.cpp:200-211
    if (P.rank[i1] < P.rank[i2]) return i1;
    if (P.rank[i1] > P.rank[i2]) return i2;
    if (P.cd[i1] > P.cd[i2]) return i1;
    if (P.cd[i1] < P.cd[i2]) return i2;
    return (rs.next() < 0.5) ? i1 : i2;
}

__device__ void sbx_and_mutate(const double *p1, const double *p2, double *c1, double *c2, const VarParams &P, PhiloxStream &rs)
{
    const unsigned nx = P.nx, ncx = nx - P.nix;
    for (unsigned i = 0; i < nx; ++i) { // children start as copies of the parents, :86-87
        c1[i] = p1[i];
        c2[i] = p2[i];
    }
    if (rs.next() < P.cr) { // :91
        for (unsigned i = 0; i < ncx; ++i) {
            const double a = p1[i], b = p2[i], yl = P.lb[i], yu = P.ub[i];
            if ((rs.next() < 0.5) && (fabs(a - b)) > 1e-14 && yl != yu) { // :94
                const double y1 = (a < b) ? a : b, y2 = (a < b) ? b : a;
                const double rand01 = rs.next();
                double beta = 1. + (2. * (y1 - yl) / (y2 - y1));
                double betaq = sbx_betaq(beta, P.eta_c, rand01);
                double v1 = 0.5 * ((y1 + y2) - betaq * (y2 - y1));
                beta = 1. + (2. * (yu - y2) / (y2 - y1));
                betaq = sbx_betaq(beta, P.eta_c, rand01);
                double v2 = 0.5 * ((y1 + y2) + betaq * (y2 - y1));
                if (v1 < yl) v1 = yl;
                if (v2 < yl) v2 = yl;
                if (v1 > yu) v1 = yu;
                if (v2 > yu) v2 = yu;
                if (rs.next() < .5) { // :119
                    c1[i] = v1;
                    c2[i] = v2;
                } else {
                    c1[i] = v2;
                    c2[i] = v1;
                }
            }
        }
        if (P.nix > 0u) { // two-point crossover of the integer part, :125-137: uniform_int(ncx, nx - 1) twice
            unsigned s1 = static_cast<unsigned>(rs.next() * static_cast<double>(P.nix)), s2;
            if (s1 >= P.nix) s1 = P.nix - 1u;
            s2 = static_cast<unsigned>(rs.next() * static_cast<double>(P.nix));
            if (s2 >= P.nix) s2 = P.nix - 1u;
            const unsigned site1 = ncx + min(s1, s2), site2 = ncx + max(s1, s2);
            for (unsigned j = site1; j <= site2; ++j) {
                c1[j] = p2[j];
                c2[j] = p1[j];
            }
        }
    }
    // polynomial mutation of the first child, then of the second (nsga2.cpp:221-222), genetic_operators.cpp:164-187
    for (int k = 0; k < 2; ++k) {
        double *c = k ? c2 : c1;
        for (unsigned j = 0; j < ncx; ++j) {
            const double yl = P.lb[j], yu = P.ub[j];
            if (rs.next() < P.m && yl != yu) {
                double y = c[j];
                const double delta1 = (y - yl) / (yu - yl), delta2 = (yu - y) / (yu - yl);
                const double rnd = rs.next();
                const double mut_pow = 1. / (P.eta_m + 1.);
                double deltaq;
                if (rnd < 0.5) {
                    const double xy = 1. - delta1;
                    const double val = 2. * rnd + (1. - 2. * rnd) * (pow(xy, (P.eta_m + 1.)));
                    deltaq = pow(val, mut_pow) - 1.;
                } else {
                    const double xy = 1. - delta2;
                    const double val = 2. * (1. - rnd) + 2. * (rnd - 0.5) * (pow(xy, (P.eta_m + 1.)));
                    deltaq = 1. - (pow(val, mut_pow));
                }
                y = y + deltaq * (yu - yl);
                if (y < yl) y = yl;
                if (y > yu) y = yu;
                c[j] = y;
            }
        }
        for (unsigned j = ncx; j < nx; ++j) { // integer mutation, :187-195: uniform_integral_from_range(lb, ub)
            if (rs.next() < P.m) {
                const long long l = static_cast<long long>(P.lb[j]), u = static_cast<long long>(P.ub[j]);
                const unsigned long long range = static_cast<unsigned long long>(u - l + 1);
                unsigned long long v = static_cast<unsigned long long>(rs.next() * static_cast<double>(range));
                if (v >= range) v = range - 1ull;
                c[j] = static_cast<double>(l + static_cast<long long>(v));
            }
        }
    }
}

// one thread per group of 4 individuals, nsga2.cpp:215-239
__global__ void nsga2_variation_kernel(const VarParams P)
{
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.NP / 4) return;
    PhiloxStream rs(P.seed, kTagNsga2Var, P.generation, g);
    const unsigned i = 4 * g;
    for (int half = 0; half < 2; ++half) {
        const unsigned *sh = half ? P.sh2 : P.sh1;
        const unsigned a = tournament(sh[i], sh[i + 1], P, rs);
        const unsigned b = tournament(sh[i + 2], sh[i + 3], P, rs);
        double *c1 = P.children + static_cast<size_t>(i + 2 * half) * P.nx;
        sbx_and_mutate(P.x + static_cast<size_t>(a) * P.nx, P.x + static_cast<size_t>(b) * P.nx, c1, c1 + P.nx, P, rs);
    }
}

__global__ void perm_keys_kernel(unsigned long long seed, unsigned tag, unsigned generation, unsigned n, unsigned long long *keys,
                                 unsigned *vals)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = philox_u64(seed, tag, generation, i, 0);
        vals[i] = i;
    }
}

__global__ void gather_rows_u32_kernel(const double *src, const unsigned *idx, unsigned rows, unsigned width, double *dst)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e < static_cast<size_t>(rows) * width) dst[e] = src[static_cast<size_t>(idx[e / width]) * width + e % width];
}

inline unsigned nblk(size_t n, unsigned t) { return static_cast<unsigned>((n + t - 1) / t); }

struct Scratch { // stream-ordered scratch from the (warm) device memory pool
    cudaStream_t st;
    std::vector<void *> owned;
    explicit Scratch(cudaStream_t s) : st(s) {}
    ~Scratch()
    {
        for (void *p : owned) cudaFreeAsync(p, st);
    }
    int alloc_bytes(void **out, size_t bytes)
    {
        void *p = nullptr;
        PGC_CUDA(cudaMallocAsync(&p, bytes ? bytes : 1, st));
        owned.push_back(p);
        *out = p;
        return PGC_OK;
    }
    template <class T> int alloc(T **out, size_t count)
    {
        void *p = nullptr;
        int rc = alloc_bytes(&p, sizeof(T) * (count ? count : 1));
        *out = static_cast<T *>(p);
        return rc;
    }
};

} // namespace

// permutation of 0..n-1 = stable argsort of the Philox keys (seed, tag, generation, i, slot 0)
int philox_permutation_device(pgc_ctx *ctx, unsigned n, unsigned long long seed, unsigned tag, unsigned generation, unsigned *d_perm,
                              cudaStream_t st)
{
    Scratch sc(st);
    unsigned long long *k0, *k1;
    unsigned *v0;
    int rc;
    if ((rc = sc.alloc(&k0, n)) || (rc = sc.alloc(&k1, n)) || (rc = sc.alloc(&v0, n))) return rc;
    perm_keys_kernel<<<nblk(n, 256), 256, 0, st>>>(seed, tag, generation, n, k0, v0);
    void *tmp = nullptr;
    size_t bytes = 0;
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0, k1, v0, d_perm, static_cast<int>(n), 0, 64, st));
    if ((rc = sc.alloc_bytes(&tmp, bytes))) return rc;
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, k0, k1, v0, d_perm, static_cast<int>(n), 0, 64, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    ctx->launches.fetch_add(2, std::memory_order_relaxed);
    return PGC_OK;
}

int nsga2_variation_device(pgc_ctx *ctx, const double *d_x, const unsigned *d_rank, const double *d_cd, unsigned NP, unsigned nx,
                           const double *d_lb, const double *d_ub, const unsigned *d_sh1, const unsigned *d_sh2, double cr,
                           double eta_c, double m, double eta_m, unsigned long long seed, unsigned generation, double *d_children,
                           cudaStream_t st, unsigned nix)
{
    PGC_REQUIRE(nix <= nx, "nsga2 variation: %u integer alleles in a chromosome of %u", nix, nx);
    PGC_REQUIRE(NP >= 5 && NP % 4 == 0,
                "for NSGA-II at least 5 individuals in the population are needed and the population size must be a multiple of "
                "4. Detected input population size is: %u",
                NP); // nsga2.cpp:121-126
    VarParams P{d_x, d_rank, d_cd, d_sh1, d_sh2, d_lb, d_ub, d_children, NP, nx, nix, cr, eta_c, m, eta_m, seed, generation};
    nsga2_variation_kernel<<<nblk(NP / 4, 64), 64, 0, st>>>(P);
    PGC_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

// The generation loop of nsga2::evolve on a device-resident population (x [NP x nx], f [NP x nobj], both updated in
// place).  `eval` is the problem's device evaluator.
int nsga2_evolve_device(pgc_problem *prob, double *d_x, double *d_f, unsigned NP, unsigned gens, double cr, double eta_c, double m,
                        double eta_m, unsigned long long seed, unsigned first_generation,
                        int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st)
{
    pgc_ctx *ctx = prob->ctx;
    const unsigned nx = static_cast<unsigned>(prob->nx), nobj = static_cast<unsigned>(prob->nobj);
    PGC_REQUIRE(nobj >= 2, "This is a multiobjective algorithm, while number of objectives detected in %s is %u", prob->name.c_str(),
                nobj); // nsga2.cpp:117-120
    PGC_REQUIRE(NP >= 5 && NP % 4 == 0,
                "for NSGA-II at least 5 individuals in the population are needed and the population size must be a multiple of "
                "4. Detected input population size is: %u",
                NP);
    PGC_REQUIRE(cr >= 0. && cr < 1., "The crossover probability must be in the [0,1[ range, while a value of %g was detected", cr);
    PGC_REQUIRE(m >= 0. && m <= 1., "The mutation probability must be in the [0,1] range, while a value of %g was detected", m);
    PGC_REQUIRE(eta_c >= 1. && eta_c <= 100., "The distribution index for crossover must be in [1, 100], while a value of %g was detected", eta_c);
    PGC_REQUIRE(eta_m >= 1. && eta_m <= 100., "The distribution index for mutation must be in [1, 100], while a value of %g was detected", eta_m);
    for (unsigned j = 0; j < nx; ++j)
        PGC_REQUIRE(prob->lb[j] != prob->ub[j], "NSGA-II cannot work on problems having a lower bound equal to an upper bound. Check your bounds.");
    Scratch sc(st);
    double *x2, *f2, *cd, *lb, *ub, *xn, *fn;
    unsigned *rank, *order, *foff, *sh1, *sh2, *sel;
    int rc;
    if ((rc = sc.alloc(&x2, static_cast<size_t>(2) * NP * nx)) || (rc = sc.alloc(&f2, static_cast<size_t>(2) * NP * nobj))
        || (rc = sc.alloc(&cd, NP)) || (rc = sc.alloc(&lb, nx)) || (rc = sc.alloc(&ub, nx)) || (rc = sc.alloc(&rank, NP))
        || (rc = sc.alloc(&order, NP)) || (rc = sc.alloc(&foff, NP + 1)) || (rc = sc.alloc(&sh1, NP)) || (rc = sc.alloc(&sh2, NP))
        || (rc = sc.alloc(&sel, 2 * NP)) || (rc = sc.alloc(&xn, static_cast<size_t>(NP) * nx)) || (rc = sc.alloc(&fn, static_cast<size_t>(NP) * nobj)))
        return rc;
    PGC_CUDA(cudaMemcpyAsync(lb, prob->lb.data(), sizeof(double) * nx, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(ub, prob->ub.data(), sizeof(double) * nx, cudaMemcpyHostToDevice, st));
    const bool trace = std::getenv("PGC_TRACE") != nullptr;
    auto now = [&]() {
        if (trace) cudaStreamSynchronize(st);
        return std::chrono::steady_clock::now();
    };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    SelectedRanking carried{rank, order, foff, 0, false};
    unsigned nfronts = 0;
    for (unsigned g = 0; g < gens; ++g) {
        const unsigned generation = first_generation + g;
        if (log_due(g + 1u) && (rc = log_ideal_device(ctx, d_f, NP, nobj, g + 1u, static_cast<unsigned long long>(g) * NP, st))) return rc; // nsga2.cpp:144-173
        const auto t0 = now();
        // parents occupy the first half of the 2N buffers (popnew = pop, nsga2.cpp:177)
        PGC_CUDA(cudaMemcpyAsync(x2, d_x, sizeof(double) * NP * nx, cudaMemcpyDeviceToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(f2, d_f, sizeof(double) * NP * nobj, cudaMemcpyDeviceToDevice, st));
        if ((rc = philox_permutation_device(ctx, NP, seed, kTagShuffle1, generation, sh1, st))) return rc;
        if ((rc = philox_permutation_device(ctx, NP, seed, kTagShuffle2, generation, sh2, st))) return rc;
        const auto t1 = now();
        // ranks and fronts of the current population: sorted from scratch for the first generation only, afterwards they come
        // with the previous generation's select_best_N_mo (see select_best_device)
        if (!carried.valid && (rc = fnds_device(ctx, d_f, NP, nobj, rank, nullptr, order, foff, &nfronts, st))) return rc;
        if (carried.valid) nfronts = carried.nfronts;
        const auto t2 = now();
        if ((rc = crowding_device(ctx, d_f, NP, nobj, order, foff, nfronts, 1, cd, st))) return rc;
        const auto t3 = now();
        if ((rc = nsga2_variation_device(ctx, d_x, rank, cd, NP, nx, lb, ub, sh1, sh2, cr, eta_c, m, eta_m, seed, generation,
                                         x2 + static_cast<size_t>(NP) * nx, st, static_cast<unsigned>(prob->nix))))
            return rc;
        if ((rc = eval(prob, x2 + static_cast<size_t>(NP) * nx, NP, f2 + static_cast<size_t>(NP) * nobj, st))) return rc;
        unsigned nsel = 0;
        const auto t4 = now();
        if ((rc = select_best_device(ctx, f2, 2 * NP, nobj, NP, sel, &nsel, st, &carried))) return rc;
        const auto t5 = now();
        if (trace)
            std::fprintf(stderr, "[pgc nsga2] gen %u: shuffles %.2f ms, fnds(N) %.2f (%u fronts), crowding %.2f, variation+eval %.2f, select(2N) %.2f\n",
                         generation, ms(t0, t1), ms(t1, t2), nfronts, ms(t2, t3), ms(t3, t4), ms(t4, t5));
        gather_rows_u32_kernel<<<nblk(static_cast<size_t>(NP) * nx, 256), 256, 0, st>>>(x2, sel, NP, nx, xn);
        gather_rows_u32_kernel<<<nblk(static_cast<size_t>(NP) * nobj, 256), 256, 0, st>>>(f2, sel, NP, nobj, fn);
        PGC_CUDA(cudaMemcpyAsync(d_x, xn, sizeof(double) * NP * nx, cudaMemcpyDeviceToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(d_f, fn, sizeof(double) * NP * nobj, cudaMemcpyDeviceToDevice, st));
        ctx->launches.fetch_add(2, std::memory_order_relaxed);
    }
    PGC_CUDA(cudaStreamSynchronize(st));
    return PGC_OK;
}

} // namespace pgc
