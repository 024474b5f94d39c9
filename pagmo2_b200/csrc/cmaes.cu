// cmaes.cu - the dense FP64 contractions of CMA-ES / xNES on the FP64 tensor path (sm_100a).
//
// Replaces, for a whole generation, reference
//   cmaes::evolve sampling  x_i = mean + sigma * B * D * z_i          src/algorithms/cmaes.cpp:246-253   (lambda x D times D x D)
//   cmaes::evolve rank-mu   C = sum_i w_i (e_i - m)(e_i - m)^T / s^2  src/algorithms/cmaes.cpp:375-380   (D x mu times mu x D)
//   cmaes::evolve mean      m = sum_i w_i e_i                          src/algorithms/cmaes.cpp:362-366
//   xnes::evolve            cov_grad = sum_i u_i (z_i z_i^T - I)       src/algorithms/xnes.cpp:302-305    (same weighted Gram matrix)
// The eigendecomposition of C (cmaes.cpp:386-401, O(D^3) on a D x D matrix every ~lambda/(10 D (c1+cmu)) generations) stays on the
// host, as SURVEY.md 8(a22) says; the caller hands in B*D.
// Both contractions run on mma.sync.m8n8k4.f64 with operands staged in shared memory in bank-conflict-free strides:
//   * weighted Gram: a CTA owns chunks of 32 individuals (centred rows in shared memory, stride == 8 mod 16), warp `a` owns
//     the 8-row block `a` of the output and all its column tiles (<= 16 accumulator tiles); per-CTA partial sums go to scratch and
//     are added in CTA order by a second kernel, so the result does not depend on scheduling (no FP64 atomics).
//   * sampling: B*D resident in shared memory (stride == 4 mod 16), a warp owns tiles of 8 individuals: normals from Philox
//     (seed, kTagCmaes, generation, i, 2j / 2j+1) by Box-Muller, DMMA, + mean, coalesced store.
// D <= 128 takes the resident-matrix kernels above; larger D the *_big_kernel variants (column groups of 16 tiles, B*D streamed in
// blocks).  The Gram kernels keep a 32-row chunk of full rows in shared memory, which bounds D to ~850 (PGC_ERR_UNSUPPORTED beyond).
// FP64-pipe bound: 2*mu*D^2 resp. 2*lambda*D^2 flop against 8*(mu*D + D*D) resp. 8*lambda*D bytes.
#include <cmath>
#include <vector>

#include "cec_device.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

#include "pgc_internal.cuh"
#include "philox.cuh"

namespace pgc
{

namespace
{

using cecdev::dmma;
using cecdev::pad4;
using cecdev::pad8;
constexpr int kMaxD = 128, kMaxNT = kMaxD / 8;
constexpr int kGramWarps = 16, kGramChunk = 32;

__host__ __device__ inline int stride_mod16(int d, int want) // smallest s >= d with s % 16 == want
{
    int s = d;
    while (s % 16 != want) ++s;
    return s;
}

struct GramParams {
    const double *rows;    // [n x D]
    const unsigned *idx;   // optional gather: row i = rows[idx[i]]
    const double *center;  // optional [D]
    const double *w;       // [k]
    unsigned k, D;
    double *partial;       // [gridDim.x][DP x DP]
};

__global__ void __launch_bounds__(kGramWarps * 32, 1) gram_partial_kernel(const GramParams P)
{
    const int D = static_cast<int>(P.D), DP = pad8(D), NT = DP / 8, S = stride_mod16(DP, 8);
    extern __shared__ __align__(16) double smem[];
    double *d = smem;                    // [kGramChunk][S]
    double *sw = d + kGramChunk * S;     // [kGramChunk]
    double *sc = sw + kGramChunk;        // [DP]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, j = lane & 3;
    for (int a = threadIdx.x; a < DP; a += blockDim.x) sc[a] = (P.center && a < D) ? P.center[a] : 0.0;
    double acc[kMaxNT][2];
#pragma unroll
    for (int t = 0; t < kMaxNT; ++t) acc[t][0] = acc[t][1] = 0.0;
    const unsigned nchunks = (P.k + kGramChunk - 1) / kGramChunk;
    for (unsigned c = blockIdx.x; c < nchunks; c += gridDim.x) {
        __syncthreads();
        for (int e = threadIdx.x; e < kGramChunk * DP; e += blockDim.x) {
            const int r = e / DP, a = e - r * DP;
            const unsigned i = c * kGramChunk + r;
            double v = 0.0;
            if (i < P.k && a < D) {
                const size_t row = P.idx ? P.idx[i] : i;
                v = P.rows[row * D + a] - sc[a];
            }
            d[r * S + a] = v;
        }
        for (int r = threadIdx.x; r < kGramChunk; r += blockDim.x) {
            const unsigned i = c * kGramChunk + r;
            sw[r] = i < P.k ? P.w[i] : 0.0;
        }
        __syncthreads();
        if (warp < NT) { // warp = 8-row block of the output; A[row a][k i] = w_i d_ia, B[k i][col b] = d_ib
            const double *pa = d + j * S + warp * 8 + g;
            const double *pb = d + j * S + g;
#pragma unroll 2
            for (int k0 = 0; k0 < kGramChunk; k0 += 4) {
                const double a = pa[k0 * S] * sw[k0 + j];
#pragma unroll
                for (int t = 0; t < kMaxNT; ++t)
                    if (t < NT) dmma(acc[t][0], acc[t][1], a, pb[k0 * S + t * 8]);
            }
        }
    }
    if (warp < NT) {
        double *out = P.partial + static_cast<size_t>(blockIdx.x) * DP * DP;
#pragma unroll
        for (int t = 0; t < kMaxNT; ++t)
            if (t < NT) {
                double *o = out + (warp * 8 + g) * DP + t * 8 + 2 * j;
                o[0] = acc[t][0];
                o[1] = acc[t][1];
            }
    }
}

// D > 128: the output has more than 16 column tiles per 8-row block, so a warp's work item is (row block, group of 16 column
// tiles) and the CTA walks over the items in passes of kGramWarps, re-reading its chunks of individuals in every pass.
__global__ void __launch_bounds__(kGramWarps * 32, 1) gram_partial_big_kernel(const GramParams P)
{
    const int D = static_cast<int>(P.D), DP = pad8(D), NT = DP / 8, S = stride_mod16(DP, 8), NCG = (NT + kMaxNT - 1) / kMaxNT;
    extern __shared__ __align__(16) double smem[];
    double *d = smem, *sw = d + kGramChunk * S, *sc = sw + kGramChunk;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, j = lane & 3;
    for (int a = threadIdx.x; a < DP; a += blockDim.x) sc[a] = (P.center && a < D) ? P.center[a] : 0.0;
    const unsigned nchunks = (P.k + kGramChunk - 1) / kGramChunk;
    const int items = NT * NCG;
    double *out = P.partial + static_cast<size_t>(blockIdx.x) * DP * DP;
    for (int first = 0; first < items; first += kGramWarps) {
        const int item = first + warp, ra = item / NCG, cg = item - ra * NCG;
        const bool valid = item < items;
        double acc[kMaxNT][2];
#pragma unroll
        for (int t = 0; t < kMaxNT; ++t) acc[t][0] = acc[t][1] = 0.0;
        for (unsigned c = blockIdx.x; c < nchunks; c += gridDim.x) {
            __syncthreads();
            for (int e = threadIdx.x; e < kGramChunk * DP; e += blockDim.x) {
                const int r = e / DP, a = e - r * DP;
                const unsigned i = c * kGramChunk + r;
                double v = 0.0;
                if (i < P.k && a < D) {
                    const size_t row = P.idx ? P.idx[i] : i;
                    v = P.rows[row * D + a] - sc[a];
                }
                d[r * S + a] = v;
            }
            for (int r = threadIdx.x; r < kGramChunk; r += blockDim.x) {
                const unsigned i = c * kGramChunk + r;
                sw[r] = i < P.k ? P.w[i] : 0.0;
            }
            __syncthreads();
            if (valid) {
                const double *pa = d + j * S + ra * 8 + g;
                const double *pb = d + j * S + cg * kMaxNT * 8 + g;
#pragma unroll 2
                for (int k0 = 0; k0 < kGramChunk; k0 += 4) {
                    const double a = pa[k0 * S] * sw[k0 + j];
#pragma unroll
                    for (int t = 0; t < kMaxNT; ++t)
                        if (cg * kMaxNT + t < NT) dmma(acc[t][0], acc[t][1], a, pb[k0 * S + t * 8]);
                }
            }
        }
        if (valid) {
#pragma unroll
            for (int t = 0; t < kMaxNT; ++t)
                if (cg * kMaxNT + t < NT) {
                    double *o = out + (ra * 8 + g) * DP + (cg * kMaxNT + t) * 8 + 2 * j;
                    o[0] = acc[t][0];
                    o[1] = acc[t][1];
                }
        }
    }
}

__global__ void gram_reduce_kernel(const double *partial, unsigned nparts, unsigned D, unsigned DP, double scale_div, double *out)
{
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D * D) return;
    const unsigned a = e / D, b = e - a * D;
    double s = 0.0;
    for (unsigned p = 0; p < nparts; ++p) s += partial[static_cast<size_t>(p) * DP * DP + a * DP + b];
    out[e] = s / scale_div; // cmaes.cpp:380  C /= sigma * sigma
}

// mean = e_0 w_0; mean += e_i w_i (cmaes.cpp:363-366), one thread per coordinate, same order and roundings as the reference
__global__ void weighted_mean_kernel(const double *rows, const unsigned *idx, const double *w, unsigned k, unsigned D, double *out)
{
    const unsigned a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= D) return;
    double m = 0.0;
    for (unsigned i = 0; i < k; ++i) {
        const size_t row = idx ? idx[i] : i;
        const double term = rows[row * D + a] * w[i];
        m = (i == 0) ? term : m + term;
    }
    out[a] = m;
}

struct SampleParams {
    const double *mean; // [D]
    const double *bd;   // [D x D] row-major: (B*D)[a][j]
    double sigma;
    unsigned lambda, D;
    unsigned long long seed;
    unsigned generation;
    double *z; // optional [lambda x D]
    double *x; // [lambda x D]
};

__device__ __forceinline__ double normal_at(unsigned long long seed, unsigned generation, unsigned i, unsigned j)
{
    const double u1 = 1.0 - philox_u01(seed, kTagCmaes, generation, i, 2 * j);
    const double u2 = philox_u01(seed, kTagCmaes, generation, i, 2 * j + 1);
    return sqrt(-2.0 * log(u1)) * cos(2.0 * 3.141592653589793238462643383279502884 * u2);
}

constexpr int kSampleWarps = 8;

__global__ void __launch_bounds__(kSampleWarps * 32, 1) cmaes_sample_kernel(const SampleParams P)
{
    const int D = static_cast<int>(P.D), DP = pad8(D), KP = pad4(D), NT = DP / 8, S = stride_mod16(KP, 4);
    extern __shared__ __align__(16) double smem[];
    double *sB = smem;              // [DP][S]: row a = output coordinate, column j
    double *sZ = sB + DP * S;       // [warps][8][S]
    double *sM = sZ + kSampleWarps * 8 * S;
    for (int e = threadIdx.x; e < DP * S; e += blockDim.x) {
        const int a = e / S, jj = e - a * S;
        sB[e] = (a < D && jj < D) ? P.bd[a * D + jj] : 0.0;
    }
    for (int a = threadIdx.x; a < DP; a += blockDim.x) sM[a] = a < D ? P.mean[a] : 0.0;
    for (int e = threadIdx.x; e < kSampleWarps * 8 * S; e += blockDim.x) sZ[e] = 0.0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, j = lane & 3;
    double *zt = sZ + warp * 8 * S;
    const unsigned ntiles = (P.lambda + 7) / 8;
    for (unsigned tile = blockIdx.x * kSampleWarps + warp; tile < ntiles; tile += gridDim.x * kSampleWarps) {
        const unsigned i0 = tile * 8;
        for (int e = lane; e < 8 * D; e += 32) {
            const int t = e / D, jj = e - t * D;
            const unsigned i = i0 + t;
            const double z = i < P.lambda ? normal_at(P.seed, P.generation, i, jj) : 0.0;
            zt[t * S + jj] = z;
            if (P.z && i < P.lambda) P.z[static_cast<size_t>(i) * D + jj] = z;
        }
        __syncwarp();
        double acc[kMaxNT][2];
#pragma unroll
        for (int t = 0; t < kMaxNT; ++t) acc[t][0] = acc[t][1] = 0.0;
        const double *pa = zt + g * S + j;  // A[row individual g][k j]
        const double *pb = sB + g * S + j;  // B[k j][col a] = BD[a][j]: row a0 + g, column k0 + j
#pragma unroll 2
        for (int k0 = 0; k0 < KP; k0 += 4) {
            const double a = pa[k0];
#pragma unroll
            for (int t = 0; t < kMaxNT; ++t)
                if (t < NT) dmma(acc[t][0], acc[t][1], a, pb[t * 8 * S + k0]);
        }
        __syncwarp();
        // accumulator tile t: lane holds y[individual g][coordinate t*8 + 2j + {0,1}]; x = mean + sigma * y
        const unsigned i = i0 + g;
#pragma unroll
        for (int t = 0; t < kMaxNT; ++t)
            if (t < NT && i < P.lambda) {
                const int c = t * 8 + 2 * j;
                if (c < D) P.x[static_cast<size_t>(i) * D + c] = sM[c] + P.sigma * acc[t][0];
                if (c + 1 < D) P.x[static_cast<size_t>(i) * D + c + 1] = sM[c + 1] + P.sigma * acc[t][1];
            }
        __syncwarp();
    }
}

// D > 128: B*D no longer fits in shared memory and an individual has more than 16 output tiles.  The CTA takes 8 tiles of
// individuals at a time (one per warp) and walks, in lock step, over groups of 128 output coordinates and blocks of kSampleKC
// inner indices: the block of B*D is staged co-operatively, every warp regenerates the matching slice of its normals (Philox
// makes z_ij a pure function of (i, j)) and accumulates.
constexpr int kSampleKC = 64;

__global__ void __launch_bounds__(kSampleWarps * 32, 1) cmaes_sample_big_kernel(const SampleParams P)
{
    const int D = static_cast<int>(P.D), DP = pad8(D), KP = pad4(D), NT = DP / 8, NCG = (NT + kMaxNT - 1) / kMaxNT;
    constexpr int S = kSampleKC + 4; // == 4 (mod 16)
    extern __shared__ __align__(16) double smem[];
    double *sB = smem;                          // [128][S]: rows = output coordinates of the group, columns = inner block
    double *sZ = sB + kMaxD * S;                // [warps][8][S]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, j = lane & 3;
    double *zt = sZ + warp * 8 * S;
    const unsigned ntiles = (P.lambda + 7) / 8;
    for (unsigned t0 = blockIdx.x * kSampleWarps; t0 < ntiles; t0 += gridDim.x * kSampleWarps) {
        const unsigned tile = t0 + warp, i0 = tile * 8;
        for (int cg = 0; cg < NCG; ++cg) {
            double acc[kMaxNT][2];
#pragma unroll
            for (int t = 0; t < kMaxNT; ++t) acc[t][0] = acc[t][1] = 0.0;
            for (int kc = 0; kc < KP; kc += kSampleKC) {
                __syncthreads();
                for (int e = threadIdx.x; e < kMaxD * kSampleKC; e += blockDim.x) {
                    const int r = e / kSampleKC, c = e - r * kSampleKC, a = cg * kMaxD + r, jj = kc + c;
                    sB[r * S + c] = (a < D && jj < D) ? P.bd[static_cast<size_t>(a) * D + jj] : 0.0;
                }
                for (int e = lane; e < 8 * kSampleKC; e += 32) {
                    const int t = e / kSampleKC, c = e - t * kSampleKC, jj = kc + c;
                    const unsigned i = i0 + t;
                    const double z = (i < P.lambda && jj < D) ? normal_at(P.seed, P.generation, i, jj) : 0.0;
                    zt[t * S + c] = z;
                    if (P.z && cg == 0 && i < P.lambda && jj < D) P.z[static_cast<size_t>(i) * D + jj] = z;
                }
                __syncthreads();
                const double *pa = zt + g * S + j, *pb = sB + g * S + j;
                const int kn = min(kSampleKC, KP - kc);
#pragma unroll 2
                for (int k0 = 0; k0 < kn; k0 += 4) {
                    const double a = pa[k0];
#pragma unroll
                    for (int t = 0; t < kMaxNT; ++t)
                        if (cg * kMaxNT + t < NT) dmma(acc[t][0], acc[t][1], a, pb[t * 8 * S + k0]);
                }
            }
            const unsigned i = i0 + g;
#pragma unroll
            for (int t = 0; t < kMaxNT; ++t)
                if (cg * kMaxNT + t < NT && i < P.lambda) {
                    const int c = (cg * kMaxNT + t) * 8 + 2 * j;
                    if (c < D) P.x[static_cast<size_t>(i) * D + c] = P.mean[c] + P.sigma * acc[t][0];
                    if (c + 1 < D) P.x[static_cast<size_t>(i) * D + c + 1] = P.mean[c + 1] + P.sigma * acc[t][1];
                }
        }
    }
}

} // namespace

int weighted_gram_device(pgc_ctx *ctx, const double *d_rows, const unsigned *d_idx, const double *d_center, const double *d_w, size_t k,
                         size_t D, double scale_div, double *d_out, cudaStream_t st)
{
    PGC_REQUIRE(D >= 1 && k >= 1, "weighted Gram matrix: empty input");
    const int DP = pad8(static_cast<int>(D)), S = stride_mod16(DP, 8);
    const size_t smem = sizeof(double) * (kGramChunk * S + kGramChunk + DP);
    if (smem > ctx->smem_optin) {
        set_error("weighted Gram matrix (cmaes rank-mu / xnes): dimension %zu needs %zu bytes of shared memory per CTA, the device offers %zu",
                  D, smem, ctx->smem_optin);
        return PGC_ERR_UNSUPPORTED;
    }
    const unsigned nchunks = static_cast<unsigned>((k + kGramChunk - 1) / kGramChunk);
    unsigned grid = nchunks < static_cast<unsigned>(ctx->sm_count) ? nchunks : static_cast<unsigned>(ctx->sm_count);
    double *partial = nullptr;
    PGC_CUDA(cudaMallocAsync(&partial, sizeof(double) * grid * DP * DP, st));
    GramParams P{d_rows, d_idx, d_center, d_w, static_cast<unsigned>(k), static_cast<unsigned>(D), partial};
    auto kern = D > static_cast<size_t>(kMaxD) ? gram_partial_big_kernel : gram_partial_kernel;
    PGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, kGramWarps * 32, smem, st>>>(P);
    PGC_CUDA(cudaGetLastError());
    const unsigned dd = static_cast<unsigned>(D * D);
    gram_reduce_kernel<<<(dd + 255) / 256, 256, 0, st>>>(partial, grid, static_cast<unsigned>(D), static_cast<unsigned>(DP), scale_div, d_out);
    PGC_CUDA(cudaGetLastError());
    PGC_CUDA(cudaFreeAsync(partial, st));
    ctx->launches.fetch_add(2, std::memory_order_relaxed);
    return PGC_OK;
}

int weighted_mean_device(pgc_ctx *ctx, const double *d_rows, const unsigned *d_idx, const double *d_w, size_t k, size_t D, double *d_out,
                         cudaStream_t st)
{
    PGC_REQUIRE(D >= 1 && k >= 1, "weighted mean: empty input");
    weighted_mean_kernel<<<static_cast<unsigned>((D + 127) / 128), 128, 0, st>>>(d_rows, d_idx, d_w, static_cast<unsigned>(k),
                                                                              static_cast<unsigned>(D), d_out);
    PGC_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

int cmaes_sample_device(pgc_ctx *ctx, const double *d_mean, const double *d_bd, double sigma, size_t lambda, size_t D, unsigned long long seed,
                        unsigned generation, double *d_z, double *d_x, cudaStream_t st)
{
    if (lambda == 0) return PGC_OK;
    PGC_REQUIRE(D >= 1, "cmaes sampling: empty dimension");
    const unsigned ntiles_all = static_cast<unsigned>((lambda + 7) / 8);
    if (D > static_cast<size_t>(kMaxD)) {
        const size_t smem_big = sizeof(double) * (static_cast<size_t>(kMaxD) * (kSampleKC + 4) + kSampleWarps * 8 * (kSampleKC + 4));
        PGC_CUDA(cudaFuncSetAttribute(cmaes_sample_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_big)));
        unsigned grid_big = (ntiles_all + kSampleWarps - 1) / kSampleWarps;
        if (grid_big > static_cast<unsigned>(ctx->sm_count)) grid_big = static_cast<unsigned>(ctx->sm_count);
        SampleParams PB{d_mean, d_bd, sigma, static_cast<unsigned>(lambda), static_cast<unsigned>(D), seed, generation, d_z, d_x};
        cmaes_sample_big_kernel<<<grid_big, kSampleWarps * 32, smem_big, st>>>(PB);
        PGC_CUDA(cudaGetLastError());
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        return PGC_OK;
    }
    const int DP = pad8(static_cast<int>(D)), KP = pad4(static_cast<int>(D)), S = stride_mod16(KP, 4);
    const size_t smem = sizeof(double) * (static_cast<size_t>(DP) * S + kSampleWarps * 8 * S + DP);
    PGC_CUDA(cudaFuncSetAttribute(cmaes_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const unsigned ntiles = static_cast<unsigned>((lambda + 7) / 8);
    unsigned grid = (ntiles + kSampleWarps - 1) / kSampleWarps;
    if (grid > static_cast<unsigned>(ctx->sm_count)) grid = static_cast<unsigned>(ctx->sm_count);
    SampleParams P{d_mean, d_bd, sigma, static_cast<unsigned>(lambda), static_cast<unsigned>(D), seed, generation, d_z, d_x};
    cmaes_sample_kernel<<<grid, kSampleWarps * 32, smem, st>>>(P);
    PGC_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

// ---- cmaes::evolve (src/algorithms/cmaes.cpp:111-407) around the device contractions ------------------------------------------
// The population, the normal draws, sampling x = mean + sigma * B * D * z, the batch evaluation, the recombination and the rank-mu
// Gram matrix run on the device; the O(D^2) bookkeeping of the evolution paths, the combination of C and its eigendecomposition
// (cmaes.cpp:385-401; Eigen's SelfAdjointEigenSolver there, a cyclic Jacobi solver here) run on the host, as SURVEY a22 scopes it.
namespace
{
__global__ void clamp_rows_kernel(double *x, size_t total, unsigned D, const double *__restrict__ lb, const double *__restrict__ ub)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const unsigned j = static_cast<unsigned>(e % D);
    const double v = x[e];
    if (v < lb[j]) x[e] = lb[j];
    else if (v > ub[j]) x[e] = ub[j]; // cmaes.cpp:304-313
}

// Cyclic Jacobi eigensolver for a symmetric matrix (row-major a[D x D], destroyed): eigenvalues ascending in w, the matching
// unit eigenvectors in the COLUMNS of v, each with its largest-magnitude component positive (the sign convention is free).
void jacobi_eigen(std::vector<double> &a, size_t D, std::vector<double> &w, std::vector<double> &v)
{
    v.assign(D * D, 0.);
    for (size_t i = 0; i < D; ++i) v[i * D + i] = 1.;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0., diag = 0.;
        for (size_t p = 0; p < D; ++p) {
            diag += a[p * D + p] * a[p * D + p];
            for (size_t q = p + 1; q < D; ++q) off += a[p * D + q] * a[p * D + q];
        }
        if (off <= 1e-32 * diag || off == 0.) break;
        for (size_t p = 0; p + 1 < D; ++p)
            for (size_t q = p + 1; q < D; ++q) {
                const double apq = a[p * D + q];
                if (apq == 0.) continue;
                const double theta = (a[q * D + q] - a[p * D + p]) / (2. * apq);
                const double t = (theta >= 0. ? 1. : -1.) / (std::fabs(theta) + std::sqrt(theta * theta + 1.));
                const double c = 1. / std::sqrt(t * t + 1.), sn = t * c;
                for (size_t k = 0; k < D; ++k) { // A <- A J (columns p, q)
                    const double akp = a[k * D + p], akq = a[k * D + q];
                    a[k * D + p] = c * akp - sn * akq;
                    a[k * D + q] = sn * akp + c * akq;
                }
                for (size_t k = 0; k < D; ++k) { // A <- J^T A (rows p, q)
                    const double apk = a[p * D + k], aqk = a[q * D + k];
                    a[p * D + k] = c * apk - sn * aqk;
                    a[q * D + k] = sn * apk + c * aqk;
                }
                for (size_t k = 0; k < D; ++k) { // V <- V J
                    const double vkp = v[k * D + p], vkq = v[k * D + q];
                    v[k * D + p] = c * vkp - sn * vkq;
                    v[k * D + q] = sn * vkp + c * vkq;
                }
            }
    }
    std::vector<size_t> order(D);
    for (size_t i = 0; i < D; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return a[x * D + x] < a[y * D + y]; });
    std::vector<double> vs(D * D);
    w.resize(D);
    for (size_t j = 0; j < D; ++j) {
        const size_t src = order[j];
        w[j] = a[src * D + src];
        size_t big = 0;
        for (size_t k = 1; k < D; ++k)
            if (std::fabs(v[k * D + src]) > std::fabs(v[big * D + src])) big = k;
        const double sgn = v[big * D + src] < 0. ? -1. : 1.;
        for (size_t k = 0; k < D; ++k) vs[k * D + j] = sgn * v[k * D + src];
    }
    v.swap(vs);
}
} // namespace

// doubles an evolution strategy built with memory = true keeps between evolve() calls (layouts in cmaes_evolve_device / xnes_evolve_device)
size_t es_state_doubles(int algo, size_t D) { return algo == PGC_ALGO_XNES ? 3 + D + D * D : 6 + 4 * D + 3 * D * D; }

int cmaes_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t lam, unsigned gens, double cc, double cs, double c1, double cmu,
                        double sigma0, double ftol, double xtol, int force_bounds, unsigned long long seed, unsigned first_generation,
                        unsigned *gens_done, double *sigma_out, int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t),
                        cudaStream_t st, double *es_state, size_t es_state_len)
{
    pgc_ctx *ctx = prob->ctx;
    const size_t D = prob->nx, mu = lam / 2u;
    PGC_REQUIRE(!es_state || es_state_len >= es_state_doubles(PGC_ALGO_CMAES, D), "cmaes: the memory array holds %zu doubles, %zu are needed", es_state_len,
                es_state_doubles(PGC_ALGO_CMAES, D));
    if (gens_done) *gens_done = 0;
    // constructor checks, cmaes.cpp:64-88, and evolve's, :126-150
    PGC_REQUIRE(((cc >= 0.) && (cc <= 1.)) || cc == -1., "cc must be in [0,1] or -1 if its value has to be initialized automatically, a value of %g was detected", cc);
    PGC_REQUIRE(((cs >= 0.) && (cs <= 1.)) || cs == -1., "cs needs to be in [0,1] or -1 if its value has to be initialized automatically, a value of %g was detected", cs);
    PGC_REQUIRE(((c1 >= 0.) && (c1 <= 1.)) || c1 == -1., "c1 needs to be in [0,1] or -1 if its value has to be initialized automatically, a value of %g was detected", c1);
    PGC_REQUIRE(((cmu >= 0.) && (cmu <= 1.)) || cmu == -1., "cmu needs to be in [0,1] or -1 if its value has to be initialized automatically, a value of %g was detected", cmu);
    PGC_REQUIRE(prob->nobj == 1, "Multiple objectives detected in %s instance. CMA-ES: Covariance Matrix Adaptation Evolutionary Strategy cannot deal with them", prob->name.c_str());
    PGC_REQUIRE(lam >= 5u, "CMA-ES: Covariance Matrix Adaptation Evolutionary Strategy needs at least 5 individuals in the population, %zu detected", lam);
    for (size_t j = 0; j < D; ++j)
        PGC_REQUIRE(std::isfinite(prob->lb[j]) && std::isfinite(prob->ub[j]), "A non-finite value is detected in the bounds, CMA-ES cannot deal with it.");
    if (gens == 0) return PGC_OK;
    const double N = static_cast<double>(D);
    // selection weights and adaptation constants, :160-186
    std::vector<double> weights(mu);
    double wsum = 0.;
    for (size_t i = 0; i < mu; ++i) {
        weights[i] = std::log(static_cast<double>(mu) + 0.5) - std::log(static_cast<double>(i) + 1.);
        wsum += weights[i];
    }
    double w2 = 0.;
    for (auto &w : weights) {
        w /= wsum;
        w2 += w * w;
    }
    const double mueff = 1. / w2;
    if (cc == -1) cc = (4. + mueff / N) / (N + 4. + 2. * mueff / N);
    if (cs == -1) cs = (mueff + 2.) / (N + mueff + 5.);
    if (c1 == -1) c1 = 2. / ((N + 1.3) * (N + 1.3) + mueff);
    if (cmu == -1) cmu = 2. * (mueff - 2. + 1. / mueff) / ((N + 2.) * (N + 2.) + mueff);
    const double damps = 1. + 2. * std::max(0., std::sqrt((mueff - 1.) / (N + 1.)) - 1.) + cs;
    const double chiN = std::sqrt(N) * (1. - 1. / (4. * N) + 1. / (21. * N * N));

    // population on the host: fitness only (lam doubles per generation) - the decision vectors stay on the device
    std::vector<double> f(lam), mean(D), meanold(D), pc(D, 0.), ps(D, 0.), dvec(D), C(D * D, 0.), Cold(D * D), Cmu(D * D), B(D * D, 0.),
        invsqrtC(D * D, 0.), BD(D * D), zlast(D), tmp(D);
    PGC_CUDA(cudaMemcpyAsync(f.data(), d_f, 8 * lam, cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    auto best_worst = [&](size_t &b, size_t &w) { // population::best_idx / worst_idx: first minimum / first maximum
        b = w = 0;
        for (size_t i = 1; i < lam; ++i) {
            if (f[i] < f[b]) b = i;
            if (f[i] > f[w]) w = i;
        }
    };
    size_t ib, iw;
    best_worst(ib, iw);
    double sigma = sigma0; // :196-225 (memory = false)
    PGC_CUDA(cudaMemcpyAsync(mean.data(), d_x + ib * D, 8 * D, cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    for (size_t j = 0; j < D; ++j) {
        dvec[j] = std::max(prob->ub[j] - prob->lb[j], 1e-6);
        B[j * D + j] = 1.;
        C[j * D + j] = dvec[j] * dvec[j];
        invsqrtC[j * D + j] = 1. / dvec[j];
    }
    unsigned long long counteval = 0, eigeneval = 0;
    // memory = true (cmaes.cpp:201-203): a state left by an earlier call on the same dimension and population size replaces the fresh start
    // layout: [1, D, lam, sigma, counteval, eigeneval | mean | pc | ps | dvec | B | C | invsqrtC]
    const bool resume = es_state && es_state[0] == 1. && es_state[1] == static_cast<double>(D) && es_state[2] == static_cast<double>(lam);
    if (resume) {
        const double *p = es_state + 6;
        sigma = es_state[3], counteval = static_cast<unsigned long long>(es_state[4]), eigeneval = static_cast<unsigned long long>(es_state[5]);
        const auto take = [&](std::vector<double> &v) {
            std::copy(p, p + v.size(), v.begin());
            p += v.size();
        };
        take(mean), take(pc), take(ps), take(dvec), take(B), take(C), take(invsqrtC);
    }
    struct SaveState { // written back however the loop ends (an exit test leaves the state as it found it)
        double *out;
        const size_t &D, &lam;
        const double &sigma;
        const unsigned long long &counteval, &eigeneval;
        const std::vector<double> &mean, &pc, &ps, &dvec, &B, &C, &invsqrtC;
        ~SaveState()
        {
            if (!out) return;
            out[0] = 1., out[1] = static_cast<double>(D), out[2] = static_cast<double>(lam), out[3] = sigma;
            out[4] = static_cast<double>(counteval), out[5] = static_cast<double>(eigeneval);
            double *p = out + 6;
            for (const std::vector<double> *v : {&mean, &pc, &ps, &dvec, &B, &C, &invsqrtC}) p = std::copy(v->begin(), v->end(), p);
        }
    } save_state{es_state, D, lam, sigma, counteval, eigeneval, mean, pc, ps, dvec, B, C, invsqrtC};

    struct Buf {
        cudaStream_t st;
        std::vector<void *> owned;
        ~Buf()
        {
            for (void *p : owned) cudaFreeAsync(p, st);
        }
        int get(void **out, size_t bytes)
        {
            PGC_CUDA(cudaMallocAsync(out, bytes ? bytes : 8, st));
            owned.push_back(*out);
            return PGC_OK;
        }
    } buf{st, {}};
    double *d_mean, *d_meanold, *d_bd, *d_z, *d_xn, *d_fn, *d_w, *d_C, *d_b;
    unsigned *d_idx;
    int rc;
    if ((rc = buf.get(reinterpret_cast<void **>(&d_mean), 8 * D)) || (rc = buf.get(reinterpret_cast<void **>(&d_meanold), 8 * D))
        || (rc = buf.get(reinterpret_cast<void **>(&d_bd), 8 * D * D)) || (rc = buf.get(reinterpret_cast<void **>(&d_z), 8 * lam * D))
        || (rc = buf.get(reinterpret_cast<void **>(&d_xn), 8 * lam * D)) || (rc = buf.get(reinterpret_cast<void **>(&d_fn), 8 * lam))
        || (rc = buf.get(reinterpret_cast<void **>(&d_w), 8 * mu)) || (rc = buf.get(reinterpret_cast<void **>(&d_C), 8 * D * D))
        || (rc = buf.get(reinterpret_cast<void **>(&d_b), 16 * D)) || (rc = buf.get(reinterpret_cast<void **>(&d_idx), 4 * mu)))
        return rc;
    PGC_CUDA(cudaMemcpyAsync(d_w, weights.data(), 8 * mu, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(d_b, prob->lb.data(), 8 * D, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(d_b + D, prob->ub.data(), 8 * D, cudaMemcpyHostToDevice, st));
    std::vector<unsigned> order(lam), idx(mu);
    unsigned done = 0;
    for (unsigned g = 0; g < gens; ++g) {
        const unsigned generation = first_generation + g;
        // 1 - lam new individuals: x_i = mean + sigma * B * D * z_i, :246-253
        for (size_t a = 0; a < D; ++a)
            for (size_t j = 0; j < D; ++j) BD[a * D + j] = B[a * D + j] * dvec[j];
        PGC_CUDA(cudaMemcpyAsync(d_mean, mean.data(), 8 * D, cudaMemcpyHostToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(d_bd, BD.data(), 8 * D * D, cudaMemcpyHostToDevice, st));
        if ((rc = cmaes_sample_device(ctx, d_mean, d_bd, sigma, lam, D, seed, generation, d_z, d_xn, st))) return rc;
        // 1bis - exit conditions, :257-273: the step of the LAST sampled individual, the spread of the current population
        PGC_CUDA(cudaMemcpyAsync(zlast.data(), d_z + (lam - 1) * D, 8 * D, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        double nrm = 0.;
        for (size_t a = 0; a < D; ++a) {
            double y = 0.;
            for (size_t j = 0; j < D; ++j) y += BD[a * D + j] * zlast[j];
            nrm += (sigma * y) * (sigma * y);
        }
        if (std::sqrt(nrm) < xtol) break;
        best_worst(ib, iw);
        if (std::fabs(f[ib] - f[iw]) < ftol) break;
        if (log_due(g + 1u)) { // 1bis - the log line, :276-296: (gen, fevals, best, dx, df, sigma)
            const double line[6] = {static_cast<double>(g + 1u), static_cast<double>(g) * static_cast<double>(lam), f[ib], std::sqrt(nrm),
                                    std::fabs(f[ib] - f[iw]), sigma};
            tls_log->host_rows.insert(tls_log->host_rows.end(), line, line + 6);
        }
        // 2 - bounds, :301-315
        if (force_bounds) {
            clamp_rows_kernel<<<static_cast<unsigned>((lam * D + 255) / 256), 256, 0, st>>>(d_xn, lam * D, static_cast<unsigned>(D), d_b, d_b + D);
            ctx->launches.fetch_add(1, std::memory_order_relaxed);
        }
        // 3 - evaluation and reinsertion (the new generation REPLACES the population, :323-350)
        if ((rc = eval(prob, d_xn, lam, d_fn, st))) return rc;
        PGC_CUDA(cudaMemcpyAsync(d_x, d_xn, 8 * lam * D, cudaMemcpyDeviceToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(d_f, d_fn, 8 * lam, cudaMemcpyDeviceToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(f.data(), d_fn, 8 * lam, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        counteval += lam;
        ++done;
        // 4 - the elite: the mu best by fitness, NaN last (:352-361; std::sort there, a stable sort here)
        for (size_t i = 0; i < lam; ++i) order[i] = static_cast<unsigned>(i);
        std::stable_sort(order.begin(), order.end(), [&](unsigned x, unsigned y) {
            const double a = f[x], b = f[y];
            return !std::isnan(a) && (std::isnan(b) || a < b);
        });
        for (size_t i = 0; i < mu; ++i) idx[i] = order[i];
        PGC_CUDA(cudaMemcpyAsync(d_idx, idx.data(), 4 * mu, cudaMemcpyHostToDevice, st));
        // 5 - new mean, :363-367; 7a - rank-mu matrix around the OLD mean, :375-380
        meanold = mean;
        PGC_CUDA(cudaMemcpyAsync(d_meanold, meanold.data(), 8 * D, cudaMemcpyHostToDevice, st));
        if ((rc = weighted_mean_device(ctx, d_xn, d_idx, d_w, mu, D, d_mean, st))) return rc;
        if ((rc = weighted_gram_device(ctx, d_xn, d_idx, d_meanold, d_w, mu, D, sigma * sigma, d_C, st))) return rc;
        PGC_CUDA(cudaMemcpyAsync(mean.data(), d_mean, 8 * D, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaMemcpyAsync(Cmu.data(), d_C, 8 * D * D, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        // 6 - evolution paths, :369-374
        for (size_t a = 0; a < D; ++a) {
            double y = 0.;
            for (size_t j = 0; j < D; ++j) y += invsqrtC[a * D + j] * (mean[j] - meanold[j]);
            tmp[a] = y;
        }
        double ps2 = 0.;
        for (size_t a = 0; a < D; ++a) {
            ps[a] = (1. - cs) * ps[a] + std::sqrt(cs * (2. - cs) * mueff) * tmp[a] / sigma;
            ps2 += ps[a] * ps[a];
        }
        const double hsig = (ps2 / N / (1. - std::pow((1. - cs), (2. * static_cast<double>(counteval) / static_cast<double>(lam))))) < (2. + 4. / (N + 1.)) ? 1. : 0.;
        for (size_t a = 0; a < D; ++a) pc[a] = (1. - cc) * pc[a] + hsig * std::sqrt(cc * (2. - cc) * mueff) * (mean[a] - meanold[a]) / sigma;
        // 7b - covariance matrix, :381
        Cold = C;
        for (size_t a = 0; a < D; ++a)
            for (size_t b = 0; b < D; ++b)
                C[a * D + b] = (1. - c1 - cmu) * Cold[a * D + b] + cmu * Cmu[a * D + b]
                               + c1 * ((pc[a] * pc[b]) + (1. - hsig) * cc * (2. - cc) * Cold[a * D + b]);
        // 8 - step size, :383
        sigma *= std::exp(std::min(0.6, (cs / damps) * (std::sqrt(ps2) / chiN - 1.)));
        // 9 - eigendecomposition every O(N) evaluations, :385-401
        if (static_cast<double>(counteval - eigeneval) > (static_cast<double>(lam) / (c1 + cmu) / N / 10.)) {
            eigeneval = counteval;
            for (size_t a = 0; a < D; ++a)
                for (size_t b = a + 1; b < D; ++b) C[a * D + b] = C[b * D + a] = (C[a * D + b] + C[b * D + a]) / 2.;
            std::vector<double> work(C), w, V;
            jacobi_eigen(work, D, w, V);
            B = V;
            for (size_t j = 0; j < D; ++j) dvec[j] = std::sqrt(std::max(1e-20, w[j]));
            for (size_t a = 0; a < D; ++a)
                for (size_t b = 0; b < D; ++b) {
                    double y = 0.;
                    for (size_t j = 0; j < D; ++j) y += B[a * D + j] * (1. / dvec[j]) * B[b * D + j];
                    invsqrtC[a * D + b] = y;
                }
        }
    }
    if (gens_done) *gens_done = done;
    if (sigma_out) *sigma_out = sigma;
    PGC_CUDA(cudaStreamSynchronize(st));
    return PGC_OK;
}

// xnes::evolve (reference src/algorithms/xnes.cpp:96-303, memory = false): the lam samples x = mean + A z and the two natural-gradient
// contractions d_center = sum u_i z_(i), sum u_i z_(i) z_(i)^T run on the device (cmaes_sample_device with BD = A, sigma = 1;
// weighted_mean_device / weighted_gram_device over all lam fitness-sorted samples); the D x D updates - mean, A <- A exp(d_A), sigma -
// on the host, with exp of the symmetric d_A through the Jacobi eigendecomposition where the reference uses Eigen's matrix exponential.
int xnes_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t lam, unsigned gens, double eta_mu, double eta_sigma, double eta_b,
                       double sigma0, double ftol, double xtol, int force_bounds, unsigned long long seed, unsigned first_generation,
                       unsigned *gens_done, double *sigma_out, int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t),
                       cudaStream_t st, double *es_state, size_t es_state_len)
{
    pgc_ctx *ctx = prob->ctx;
    const size_t D = prob->nx;
    if (gens_done) *gens_done = 0;
    PGC_REQUIRE(!es_state || es_state_len >= es_state_doubles(PGC_ALGO_XNES, D), "xnes: the memory array holds %zu doubles, %zu are needed", es_state_len,
                es_state_doubles(PGC_ALGO_XNES, D));
    // constructor checks, xnes.cpp:55-78, and evolve's, :110-137
    PGC_REQUIRE((eta_mu > 0. && eta_mu <= 1.) || eta_mu == -1., "eta_mu must be in ]0,1] or -1 if its value has to be initialized automatically, a value of %g was detected", eta_mu);
    PGC_REQUIRE((eta_sigma > 0. && eta_sigma <= 1.) || eta_sigma == -1., "eta_sigma needs to be in ]0,1] or -1 if its value has to be initialized automatically, a value of %g was detected", eta_sigma);
    PGC_REQUIRE((eta_b > 0. && eta_b <= 1.) || eta_b == -1., "eta_b needs to be in ]0,1] or -1 if its value has to be initialized automatically, a value of %g was detected", eta_b);
    PGC_REQUIRE((sigma0 > 0. && sigma0 <= 1.) || sigma0 == -1., "sigma0 needs to be in ]0,1] or -1 if its value has to be initialized automatically, a value of %g was detected", sigma0);
    PGC_REQUIRE(prob->nobj == 1, "Multiple objectives detected in %s instance. xNES: Exponential Natural Evolution Strategies cannot deal with them", prob->name.c_str());
    PGC_REQUIRE(lam >= 4u, "xNES: Exponential Natural Evolution Strategies needs at least 5 individuals in the population, %zu detected", lam);
    for (size_t j = 0; j < D; ++j)
        PGC_REQUIRE(std::isfinite(prob->lb[j]) && std::isfinite(prob->ub[j]), "A non-finite value is detected in the bounds, xNES cannot deal with it.");
    if (gens == 0) return PGC_OK;
    const double dim_d = static_cast<double>(D), lam_d = static_cast<double>(lam);
    if (eta_mu == -1) eta_mu = 1.;
    const double common_default = 0.6 * (3. + std::log(dim_d)) / (dim_d * std::sqrt(dim_d)); // :143-150
    if (eta_sigma == -1) eta_sigma = common_default;
    if (eta_b == -1) eta_b = common_default;
    std::vector<double> u(lam); // utilities, :151-161
    double sum = 0.;
    for (size_t i = 0; i < lam; ++i) u[i] = std::max(0., std::log(lam_d / 2. + 1.) - std::log(static_cast<double>(i + 1)));
    for (size_t i = 0; i < lam; ++i) sum += u[i];
    for (size_t i = 0; i < lam; ++i) u[i] = u[i] / sum - 1. / lam_d;
    double usum = 0.;
    for (size_t i = 0; i < lam; ++i) usum += u[i];
    double sigma = sigma0 == -1 ? 0.5 : sigma0;
    std::vector<double> f(lam), A(D * D, 0.), mean(D), z0(D), dc(D), G(D * D), dA(D * D), E(D * D), An(D * D), tmp(D);
    for (size_t j = 0; j < D; ++j) A[j * D + j] = std::max(prob->ub[j] - prob->lb[j], 1e-6) * sigma;
    PGC_CUDA(cudaMemcpyAsync(f.data(), d_f, 8 * lam, cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    size_t ib = 0, iw = 0;
    for (size_t i = 1; i < lam; ++i)
        if (f[i] < f[ib]) ib = i;
    PGC_CUDA(cudaMemcpyAsync(mean.data(), d_x + ib * D, 8 * D, cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    // memory = true (xnes.cpp:163): a state left by an earlier call on the same dimension replaces the fresh start.  [1, D, sigma | mean | A]
    if (es_state && es_state[0] == 1. && es_state[1] == static_cast<double>(D)) {
        sigma = es_state[2];
        std::copy(es_state + 3, es_state + 3 + D, mean.begin());
        std::copy(es_state + 3 + D, es_state + 3 + D + D * D, A.begin());
    }
    struct SaveState {
        double *out;
        const size_t &D;
        const double &sigma;
        const std::vector<double> &mean, &A;
        ~SaveState()
        {
            if (!out) return;
            out[0] = 1., out[1] = static_cast<double>(D), out[2] = sigma;
            std::copy(A.begin(), A.end(), std::copy(mean.begin(), mean.end(), out + 3));
        }
    } save_state{es_state, D, sigma, mean, A};
    struct Buf {
        cudaStream_t st;
        std::vector<void *> owned;
        ~Buf()
        {
            for (void *p : owned) cudaFreeAsync(p, st);
        }
        int get(void **out, size_t bytes)
        {
            PGC_CUDA(cudaMallocAsync(out, bytes ? bytes : 8, st));
            owned.push_back(*out);
            return PGC_OK;
        }
    } buf{st, {}};
    double *d_mean, *d_A, *d_z, *d_xn, *d_fn, *d_u, *d_dc, *d_G, *d_b;
    unsigned *d_idx;
    int rc;
    if ((rc = buf.get(reinterpret_cast<void **>(&d_mean), 8 * D)) || (rc = buf.get(reinterpret_cast<void **>(&d_A), 8 * D * D))
        || (rc = buf.get(reinterpret_cast<void **>(&d_z), 8 * lam * D)) || (rc = buf.get(reinterpret_cast<void **>(&d_xn), 8 * lam * D))
        || (rc = buf.get(reinterpret_cast<void **>(&d_fn), 8 * lam)) || (rc = buf.get(reinterpret_cast<void **>(&d_u), 8 * lam))
        || (rc = buf.get(reinterpret_cast<void **>(&d_dc), 8 * D)) || (rc = buf.get(reinterpret_cast<void **>(&d_G), 8 * D * D))
        || (rc = buf.get(reinterpret_cast<void **>(&d_b), 16 * D)) || (rc = buf.get(reinterpret_cast<void **>(&d_idx), 4 * lam)))
        return rc;
    PGC_CUDA(cudaMemcpyAsync(d_u, u.data(), 8 * lam, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(d_b, prob->lb.data(), 8 * D, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(d_b + D, prob->ub.data(), 8 * D, cudaMemcpyHostToDevice, st));
    std::vector<unsigned> order(lam);
    unsigned done = 0;
    for (unsigned g = 0; g < gens; ++g) {
        const unsigned generation = first_generation + g;
        // 1 - lam new individuals x_i = mean + A z_i, evaluated as they are written into the population (pop.set_x, :196-216)
        PGC_CUDA(cudaMemcpyAsync(d_mean, mean.data(), 8 * D, cudaMemcpyHostToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(d_A, A.data(), 8 * D * D, cudaMemcpyHostToDevice, st));
        if ((rc = cmaes_sample_device(ctx, d_mean, d_A, 1.0, lam, D, seed, generation, d_z, d_xn, st))) return rc;
        if (force_bounds) {
            clamp_rows_kernel<<<static_cast<unsigned>((lam * D + 255) / 256), 256, 0, st>>>(d_xn, lam * D, static_cast<unsigned>(D), d_b, d_b + D);
            ctx->launches.fetch_add(1, std::memory_order_relaxed);
        }
        if ((rc = eval(prob, d_xn, lam, d_fn, st))) return rc;
        PGC_CUDA(cudaMemcpyAsync(d_x, d_xn, 8 * lam * D, cudaMemcpyDeviceToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(d_f, d_fn, 8 * lam, cudaMemcpyDeviceToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(f.data(), d_fn, 8 * lam, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaMemcpyAsync(z0.data(), d_z, 8 * D, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        ++done;
        // 2 - exit conditions on the step of the FIRST sample and the spread of the new population, :219-236
        double nrm = 0.;
        for (size_t a = 0; a < D; ++a) {
            double y = 0.;
            for (size_t j = 0; j < D; ++j) y += A[a * D + j] * z0[j];
            nrm += y * y;
        }
        if (std::sqrt(nrm) < xtol) break;
        ib = iw = 0;
        for (size_t i = 1; i < lam; ++i) {
            if (f[i] < f[ib]) ib = i;
            if (f[i] > f[iw]) iw = i;
        }
        if (std::fabs(f[ib] - f[iw]) < ftol) break;
        if (log_due(g + 1u)) { // the log line, :238-256: (gen, fevals, best, dx, df, sigma)
            const double line[6] = {static_cast<double>(g + 1u), static_cast<double>(g + 1u) * lam_d, f[ib], std::sqrt(nrm), std::fabs(f[ib] - f[iw]), sigma};
            tls_log->host_rows.insert(tls_log->host_rows.end(), line, line + 6);
        }
        // 3 - the samples in order of fitness (plain <, :258-261; std::sort there, a stable sort here), the two gradients
        for (size_t i = 0; i < lam; ++i) order[i] = static_cast<unsigned>(i);
        std::stable_sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { return f[a] < f[b]; });
        PGC_CUDA(cudaMemcpyAsync(d_idx, order.data(), 4 * lam, cudaMemcpyHostToDevice, st));
        if ((rc = weighted_mean_device(ctx, d_z, d_idx, d_u, lam, D, d_dc, st))) return rc;               // d_center, :264-267
        if ((rc = weighted_gram_device(ctx, d_z, d_idx, nullptr, d_u, lam, D, 1.0, d_G, st))) return rc; // sum u_i z z^T, :268-271
        PGC_CUDA(cudaMemcpyAsync(dc.data(), d_dc, 8 * D, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaMemcpyAsync(G.data(), d_G, 8 * D * D, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        for (size_t a = 0; a < D; ++a) G[a * D + a] -= usum; // ... - sum u_i I
        double cov_trace = 0.;
        for (size_t a = 0; a < D; ++a) cov_trace += G[a * D + a];
        for (size_t a = 0; a < D; ++a) G[a * D + a] -= cov_trace / dim_d; // :273
        for (size_t a = 0; a < D; ++a)
            for (size_t b = 0; b < D; ++b) dA[a * D + b] = 0.5 * ((a == b ? eta_sigma * cov_trace / dim_d : 0.) + eta_b * G[a * D + b]); // :274
        // 4 - the updates, :276-291
        for (size_t a = 0; a < D; ++a) {
            double y = 0.;
            for (size_t j = 0; j < D; ++j) y += A[a * D + j] * dc[j];
            tmp[a] = y;
        }
        for (size_t a = 0; a < D; ++a) mean[a] = mean[a] + eta_mu * tmp[a];
        for (size_t a = 0; a < D; ++a)
            for (size_t b = a + 1; b < D; ++b) dA[a * D + b] = dA[b * D + a] = (dA[a * D + b] + dA[b * D + a]) / 2.;
        std::vector<double> work(dA), w, V;
        jacobi_eigen(work, D, w, V);
        for (size_t a = 0; a < D; ++a)
            for (size_t b = 0; b < D; ++b) {
                double y = 0.;
                for (size_t k = 0; k < D; ++k) y += V[a * D + k] * std::exp(w[k]) * V[b * D + k];
                E[a * D + b] = y;
            }
        for (size_t a = 0; a < D; ++a)
            for (size_t b = 0; b < D; ++b) {
                double y = 0.;
                for (size_t k = 0; k < D; ++k) y += A[a * D + k] * E[k * D + b];
                An[a * D + b] = y;
            }
        A = An;
        sigma = sigma * std::exp(eta_sigma / 2. * cov_trace / dim_d);
    }
    PGC_CUDA(cudaStreamSynchronize(st));
    if (gens_done) *gens_done = done;
    if (sigma_out) *sigma_out = sigma;
    return PGC_OK;
}

} // namespace pgc
