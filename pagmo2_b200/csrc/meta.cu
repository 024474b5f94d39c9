// meta.cu - meta-problems on the device: translate and decompose.
//
//   translate : fitness(x) = inner.fitness(x - t)                 reference src/problems/translate.cpp:100-153
//               (batch_fitness de-shifts every row with std::minus, :137-150, then calls the inner batch_fitness)
//   decompose : fitness(x) = decompose_objectives(inner.fitness(x), weight, z, method)
//               reference src/problems/decompose.cpp:139-154, src/utils/multi_objective.cpp:582-638
//
// Both wrap an existing pgc_problem (borrowed: the inner problem must outlive the wrapper) and run as one extra
// element-wise kernel before / after the inner evaluator on the same stream.  The translated rows keep the reference's
// two separate subtractions (x - t, then the inner problem's own shift): nothing is folded, so roundings match.
// decompose's ideal-point adaptation (m_adapt_ideal, decompose.cpp:143-149) mutates z after every single fitness call
// in call order - a sequential semantic the batch path cannot reproduce; it is refused (PGC_ERR_UNSUPPORTED).
#include <cmath>

#include "pgc_internal.cuh"

namespace pgc
{

namespace
{

inline int cuda_ok(cudaError_t e, const char *what) { return e == cudaSuccess ? PGC_OK : cuda_fail(e, what, __FILE__, __LINE__); }

// xs[i][j] - t[j], translate.cpp:145-147
__global__ void translate_rows_kernel(const double *__restrict__ x, const double *__restrict__ t, double *__restrict__ out, size_t total,
                                      unsigned nx)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride)
        out[e] = x[e] - t[e % nx];
}

// decompose_objectives per row, multi_objective.cpp:602-632 (same loop order, no fused multiply-add: -fmad=false)
__global__ void decompose_rows_kernel(const double *__restrict__ f, const double *__restrict__ w, const double *__restrict__ z,
                                      double *__restrict__ out, size_t n, unsigned m, int method)
{
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *fi = f + i * m;
    double fd = 0.0;
    if (method == PGC_DECOMPOSE_WEIGHTED) { // :603-606
        for (unsigned k = 0; k < m; ++k) fd += w[k] * fi[k];
    } else if (method == PGC_DECOMPOSE_TCHEBYCHEFF) { // :607-616
        for (unsigned k = 0; k < m; ++k) {
            const double fixed_weight = (w[k] == 0.0) ? 1e-4 : w[k];
            const double tmp = fixed_weight * fabs(fi[k] - z[k]);
            if (tmp > fd) fd = tmp;
        }
    } else { // boundary intersection, :617-632
        const double THETA = 5.0;
        double d1 = 0.0, weight_norm = 0.0;
        for (unsigned k = 0; k < m; ++k) {
            d1 += (fi[k] - z[k]) * w[k];
            weight_norm += w[k] * w[k]; // std::pow(weight, 2) is the exact square
        }
        weight_norm = sqrt(weight_norm);
        d1 = d1 / weight_norm;
        double d2 = 0.0;
        for (unsigned k = 0; k < m; ++k) {
            const double d = fi[k] - (z[k] + d1 * w[k] / weight_norm);
            d2 += d * d;
        }
        d2 = sqrt(d2);
        fd = d1 + THETA * d2;
    }
    out[i] = fd;
}

} // namespace

int meta_eval(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t s)
{
    if (n == 0) return PGC_OK;
    pgc_problem *inner = p->inner;
    pgc_ctx *ctx = p->ctx;
    if (p->desc.family == PGC_TRANSLATE) {
        double *tmp = nullptr;
        const size_t total = n * p->nx;
        PGC_CUDA(cudaMallocAsync(&tmp, sizeof(double) * total, s));
        const unsigned blocks = static_cast<unsigned>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(ctx->sm_count) * 16));
        translate_rows_kernel<<<blocks, 256, 0, s>>>(d_dvs, p->d_meta, tmp, total, static_cast<unsigned>(p->nx));
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        int rc = cuda_ok(cudaGetLastError(), "translate_rows_kernel");
        if (rc == PGC_OK) rc = problem_eval_device(inner, tmp, n, d_fvs, s);
        cudaFreeAsync(tmp, s);
        return rc;
    }
    // decompose
    const size_t m = inner->nobj;
    double *ftmp = nullptr;
    PGC_CUDA(cudaMallocAsync(&ftmp, sizeof(double) * n * m, s));
    int rc = problem_eval_device(inner, d_dvs, n, ftmp, s);
    if (rc == PGC_OK) {
        decompose_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(ftmp, p->d_meta, p->d_meta + m, d_fvs, n,
                                                                                      static_cast<unsigned>(m), p->meta_method);
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        rc = cuda_ok(cudaGetLastError(), "decompose_rows_kernel");
    }
    cudaFreeAsync(ftmp, s);
    return rc;
}

int meta_create(pgc_problem *inner, int family, const double *a, const double *b, size_t len, int method, pgc_problem **out)
{
    PGC_REQUIRE(inner && out, "meta problem: null argument");
    *out = nullptr;
    PGC_REQUIRE(a && (family == PGC_TRANSLATE || b), "meta problem: null vector");
    if (family == PGC_TRANSLATE) {
        // translate.cpp:83-87
        PGC_REQUIRE(len == inner->nx, "Length of shift vector is: %zu while the problem dimension is: %zu", len, inner->nx);
    } else {
        // decompose.cpp:68-124
        PGC_REQUIRE(inner->nobj >= 2, "Decomposition can only be applied to multi-objective problems");
        PGC_REQUIRE(method == PGC_DECOMPOSE_WEIGHTED || method == PGC_DECOMPOSE_TCHEBYCHEFF || method == PGC_DECOMPOSE_BI,
                    "Decomposition method requested is: %d while only one of ['weighted', 'tchebycheff', 'bi'] are allowed", method);
        PGC_REQUIRE(len == inner->nobj,
                    "Weight vector size must be equal to the number of objectives. The size of the weight vector is %zu while the "
                    "problem has %zu objectives",
                    len, inner->nobj);
        double sum = 0.0;
        for (size_t i = 0; i < len; ++i) {
            PGC_REQUIRE(std::isfinite(a[i]), "Weight contains non finite numbers");
            PGC_REQUIRE(std::isfinite(b[i]), "Reference point contains non finite numbers");
            sum += a[i];
        }
        PGC_REQUIRE(std::fabs(sum - 1.0) <= 1e-8, "The weight vector must sum to 1 with a tolerance of 1E-8. The sum of the weight "
                                                   "vector components was detected to be: %f", sum);
        for (size_t i = 0; i < len; ++i)
            PGC_REQUIRE(a[i] >= 0.0, "The weight vector may contain only non negative values. A value of %f was detected at index %zu",
                        a[i], i);
    }
    pgc_problem *p = new (std::nothrow) pgc_problem;
    if (!p) return PGC_ERR_OUT_OF_MEMORY;
    p->ctx = inner->ctx;
    p->desc = inner->desc;
    p->desc.family = family;
    p->inner = inner;
    p->meta_method = method;
    p->nx = inner->nx;
    p->nix = inner->nix; // translate.cpp:169-172 / decompose.cpp:180-183: get_nix forwards to the inner problem
    p->lb = inner->lb;
    p->ub = inner->ub;
    p->flops_per_eval = inner->flops_per_eval;
    p->transc_per_eval = inner->transc_per_eval;
    std::vector<double> host(a, a + len);
    if (family == PGC_TRANSLATE) {
        // translate.cpp:175-181: the bounds move with the problem
        for (size_t i = 0; i < len; ++i) {
            p->lb[i] = inner->lb[i] + a[i];
            p->ub[i] = inner->ub[i] + a[i];
        }
        p->nobj = inner->nobj;
        p->name = inner->name + " [translated]"; // translate.cpp:355-358
        p->flops_per_eval += static_cast<double>(len);
    } else {
        host.insert(host.end(), b, b + len);
        p->nobj = 1;
        p->name = inner->name + " [decomposed]"; // decompose.cpp:226-229
        p->flops_per_eval += 2.0 * static_cast<double>(len);
    }
    cudaError_t e = cudaSetDevice(p->ctx->device);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_meta, sizeof(double) * host.size());
    if (e == cudaSuccess) e = cudaMemcpy(p->d_meta, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(p->d_meta);
        delete p;
        return cuda_fail(e, "meta problem tables", __FILE__, __LINE__);
    }
    *out = p;
    return PGC_OK;
}

void meta_destroy(pgc_problem *p)
{
    cudaFree(p->d_meta);
    p->d_meta = nullptr;
}

} // namespace pgc
