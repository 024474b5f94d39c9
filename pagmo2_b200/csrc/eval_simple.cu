// eval_simple.cu - batch fitness of pagmo's simple single-objective UDPs on sm_100a:
//   rastrigin (src/problems/rastrigin.cpp:62-72), ackley (ackley.cpp:61-76), griewank (griewank.cpp:60-75),
//   schwefel (schwefel.cpp:60-69), rosenbrock (rosenbrock.cpp:59-66).
//
// These are O(D) per individual: HBM-bound except for the FP64 transcendental per coordinate.  A CTA stages a
// tile of 128 individuals x (up to 32 coordinates) in shared memory with coalesced 16-byte loads, then each
// thread walks ITS individual's coordinates in the reference's order (j ascending), so every sum/product is
// accumulated exactly as the reference does it; the only differences are libdevice-vs-glibc ulps of cos/sin/exp.
// Compiled with -fmad=false (no contraction: the reference is built for baseline x86-64, which has no FMA).
#include <cmath>

#include "pgc_internal.cuh"
#include "simple_device.cuh"

namespace pgc
{

namespace
{

constexpr int kTile = 128;  // individuals per CTA tile == threads per CTA
constexpr int kChunk = 32;  // coordinates staged per pass
constexpr int kStride = kChunk + 1; // odd stride: conflict-free column walks

using simple::Acc;
using simple::acc_final;
using simple::acc_init;
using simple::acc_step;

template <int FAM>
__global__ void __launch_bounds__(kTile) simple_kernel(const double *__restrict__ x, double *__restrict__ f, long long n, int D)
{
    __shared__ double tile[kTile * kStride];
    const int tid = threadIdx.x;
    const long long ntiles = (n + kTile - 1) / kTile;
    for (long long tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
        const long long t0 = tb * kTile;
        const int nt = (n - t0 < kTile) ? static_cast<int>(n - t0) : kTile;
        Acc s;
        acc_init<FAM>(s);
        // rosenbrock needs x[j+1]: stage one extra coordinate per pass (overlapping chunks)
        const int extra = (FAM == PGC_ROSENBROCK) ? 1 : 0;
        for (int j0 = 0; j0 < D; j0 += kChunk) {
            const int nj = (D - j0 < kChunk) ? D - j0 : kChunk;
            const int njl = (j0 + nj + extra <= D) ? nj + extra : nj; // columns to load
            // coalesced stage: element (t, jj) <- x[(t0+t)*D + j0 + jj]
            for (int e = tid; e < nt * njl; e += kTile) {
                const int t = e / njl, jj = e - t * njl;
                tile[t * kStride + jj] = __ldcs(x + (t0 + t) * D + j0 + jj);
            }
            __syncthreads();
            if (tid < nt) {
                const double *row = tile + tid * kStride;
                for (int jj = 0; jj < nj; ++jj) {
                    const bool has_next = (j0 + jj + 1 < D);
                    const double xn = (FAM == PGC_ROSENBROCK && has_next) ? row[jj + 1] : 0.0;
                    acc_step<FAM>(s, row[jj], j0 + jj, has_next, xn);
                }
            }
            __syncthreads();
        }
        if (tid < nt) f[t0 + tid] = acc_final<FAM>(s, D);
    }
}

template <int FAM> int launch(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream)
{
    const long long ntiles = (static_cast<long long>(n) + kTile - 1) / kTile;
    long long blocks = ntiles;
    const long long cap = static_cast<long long>(p->ctx->sm_count) * 12; // 12 CTAs of 128 threads fit per SM
    if (blocks > cap) blocks = cap;
    simple_kernel<FAM><<<static_cast<unsigned>(blocks), kTile, 0, stream>>>(d_dvs, d_fvs, static_cast<long long>(n),
                                                                            static_cast<int>(p->nx));
    PGC_CUDA(cudaGetLastError());
    p->ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

} // namespace

int simple_create(pgc_problem *p)
{
    const unsigned D = p->desc.dim;
    double lo = 0, hi = 0;
    const char *nm = "";
    switch (p->desc.family) {
        case PGC_RASTRIGIN: lo = -5.12; hi = 5.12; nm = "Rastrigin Function"; break;        // rastrigin.cpp:80-85
        case PGC_ACKLEY: lo = -15; hi = 30; nm = "Ackley Function"; break;                  // ackley.cpp:85-90
        case PGC_GRIEWANK: lo = -600; hi = 600; nm = "Griewank Function"; break;            // griewank.cpp:84-89
        case PGC_SCHWEFEL: lo = -500; hi = 500; nm = "Schwefel Function"; break;            // schwefel.cpp:77-82
        case PGC_ROSENBROCK: lo = -5; hi = 10; nm = "Multidimensional Rosenbrock Function"; break; // rosenbrock.cpp:72-75
        default: set_error("simple_create: family %d is not a simple UDP", p->desc.family); return PGC_ERR_INVALID_ARGUMENT;
    }
    if (p->desc.family == PGC_ROSENBROCK) {
        PGC_REQUIRE(D >= 2u, "Rosenbrock Function must have minimum 2 dimensions, %u requested", D); // rosenbrock.cpp:46-50
    } else {
        PGC_REQUIRE(D >= 1u, "%s must have minimum 1 dimension, %u requested", nm, D); // rastrigin.cpp:46-50 etc.
    }
    p->nx = D;
    p->nobj = 1;
    p->lb.assign(D, lo);
    p->ub.assign(D, hi);
    p->name = nm;
    const double d = D;
    switch (p->desc.family) {
        case PGC_RASTRIGIN: p->flops_per_eval = 5 * d + 2; p->transc_per_eval = d; break;
        case PGC_ACKLEY: p->flops_per_eval = 4 * d + 10; p->transc_per_eval = d + 4; break;
        case PGC_GRIEWANK: p->flops_per_eval = 5 * d + 3; p->transc_per_eval = 2 * d; break;
        case PGC_SCHWEFEL: p->flops_per_eval = 2 * d + 2; p->transc_per_eval = 2 * d; break;
        case PGC_ROSENBROCK: p->flops_per_eval = 10 * (d - 1); p->transc_per_eval = 0; break;
    }
    return PGC_OK;
}

int simple_eval(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream)
{
    if (n == 0) return PGC_OK;
    switch (p->desc.family) {
        case PGC_RASTRIGIN: return launch<PGC_RASTRIGIN>(p, d_dvs, n, d_fvs, stream);
        case PGC_ACKLEY: return launch<PGC_ACKLEY>(p, d_dvs, n, d_fvs, stream);
        case PGC_GRIEWANK: return launch<PGC_GRIEWANK>(p, d_dvs, n, d_fvs, stream);
        case PGC_SCHWEFEL: return launch<PGC_SCHWEFEL>(p, d_dvs, n, d_fvs, stream);
        case PGC_ROSENBROCK: return launch<PGC_ROSENBROCK>(p, d_dvs, n, d_fvs, stream);
        default: set_error("simple_eval: bad family"); return PGC_ERR_INVALID_ARGUMENT;
    }
}

} // namespace pgc
