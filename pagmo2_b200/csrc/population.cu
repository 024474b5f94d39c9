// population.cu - random initial populations on the device.
// Replaces reference batch_random_decision_vector (include/pagmo/utils/generic.hpp:326-389) + the population constructor from a
// bfe (src/population.cpp:82-103): x_ij = lb_j + (ub_j - lb_j) * u (lb_j when the bounds coincide, generic.hpp:378-380), one batch
// evaluation, and a random 64-bit ID per individual (population.cpp:155-160 draws them from the population's engine).
// Draws are Philox (seed, kTagPopulation, 0, i, j) for the genes and (seed, kTagPopulation, 1, i, 0) for the IDs; the last nix genes
// (zdt5) are integers in [lb, ub] = lb + floor(u * (ub - lb + 1)).
#include "pgc_internal.cuh"
#include "philox.cuh"

namespace pgc
{

namespace
{
__global__ void random_dvs_kernel(double *x, unsigned long long *ids, size_t n, unsigned nx, unsigned ncx, const double *__restrict__ lb,
                                  const double *__restrict__ ub, unsigned long long seed)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n * nx) return;
    const unsigned i = static_cast<unsigned>(e / nx), j = static_cast<unsigned>(e - static_cast<size_t>(i) * nx);
    const double l = lb[j], u = ub[j];
    if (j < ncx) {
        x[e] = (l == u) ? l : (u - l) * philox_u01(seed, kTagPopulation, 0, i, j) + l;
    } else { // integer tail: uniform_integral_from_range (generic.hpp:142-164, :289-295), an integer in [lb, ub]
        const double span = u - l + 1.;
        x[e] = l + fmin(floor(philox_u01(seed, kTagPopulation, 0, i, j) * span), u - l);
    }
    if (j == 0 && ids) ids[i] = philox_u64(seed, kTagPopulation, 1, i, 0);
}
} // namespace

int population_init_device(pgc_problem *prob, size_t n, unsigned long long seed, double *d_x, double *d_f, unsigned long long *d_ids,
                           int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st)
{
    if (n == 0) return PGC_OK;
    PGC_REQUIRE(n < 0xffffffffull, "population_init: population too large");
    const size_t nx = prob->nx;
    for (size_t j = 0; j < nx; ++j) // generic.hpp:347-351 / :66-95
        PGC_REQUIRE(std::isfinite(prob->lb[j]) && std::isfinite(prob->ub[j]), "Cannot generate a random real if the bounds are not finite");
    double *d_b = nullptr;
    PGC_CUDA(cudaMallocAsync(&d_b, 2 * nx * sizeof(double), st));
    PGC_CUDA(cudaMemcpyAsync(d_b, prob->lb.data(), nx * sizeof(double), cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(d_b + nx, prob->ub.data(), nx * sizeof(double), cudaMemcpyHostToDevice, st));
    const size_t tot = n * nx;
    random_dvs_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, st>>>(d_x, d_ids, n, static_cast<unsigned>(nx),
                                                                                 static_cast<unsigned>(nx - prob->nix), d_b, d_b + nx, seed);
    PGC_CUDA(cudaGetLastError());
    prob->ctx->launches.fetch_add(1, std::memory_order_relaxed);
    PGC_CUDA(cudaStreamSynchronize(st)); // lb/ub staging came from pageable host vectors
    PGC_CUDA(cudaFreeAsync(d_b, st));
    return d_f ? eval(prob, d_x, n, d_f, st) : PGC_OK;
}

} // namespace pgc
