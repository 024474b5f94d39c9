// meta.cu - meta-problems on the device: translate, decompose and unconstrain.
//
//   translate : fitness(x) = inner.fitness(x - t)                 reference src/problems/translate.cpp:100-153
//               (batch_fitness de-shifts every row with std::minus, :137-150, then calls the inner batch_fitness)
//   decompose : fitness(x) = decompose_objectives(inner.fitness(x), weight, z, method)
//               reference src/problems/decompose.cpp:139-154, src/utils/multi_objective.cpp:582-638
//
//   unconstrain : fitness(x) = penalize(inner.fitness(x)): the constraints folded into the objectives by one of five methods
//               reference src/problems/unconstrain.cpp:136-223 (batch_fitness :244-263 penalizes row by row)
//
// All wrap an existing pgc_problem (borrowed: the inner problem must outlive the wrapper) and run as one extra
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

__device__ __forceinline__ double max0(double a) { return a < 0. ? 0. : a; } // std::max(a, 0.): a NaN stays a NaN (never satisfied)

// unconstrain::penalize per row, unconstrain.cpp:136-223.  tw = [c_tol (nc) | weights (nc)]; rows of f are [nobj | nec | nic].
__global__ void unconstrain_rows_kernel(const double *__restrict__ f, const double *__restrict__ tw, double *__restrict__ out, size_t n,
                                        unsigned nobj, unsigned nec, unsigned nic, int method)
{
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned nc = nec + nic;
    const double *fi = f + i * (nobj + nc);
    const double *c = fi + nobj;
    const double *tol = tw, *w = tw + nc;
    // test_eq_constraints / test_ineq_constraints (constrained.hpp:49-80): satisfied counts and the two violation norms
    unsigned sat = 0;
    double l2e = 0., l2i = 0.;
    for (unsigned k = 0; k < nec; ++k) {
        const double err = max0(fabs(c[k]) - tol[k]);
        l2e += err * err;
        sat += (err <= 0.) ? 1u : 0u;
    }
    for (unsigned k = nec; k < nc; ++k) {
        const double err = max0(c[k] - tol[k]);
        l2i += err * err;
        sat += (err <= 0.) ? 1u : 0u;
    }
    if (method == PGC_UNCONSTRAIN_IGNORE_O) { // :206-221: one objective, the norm of the violation
        out[i] = sqrt(l2e) + sqrt(l2i);
        return;
    }
    double *o = out + i * nobj;
    const bool feasible = sat == nc; // problem::feasibility_f, problem.cpp:709-721
    double add = 0.;
    bool overwrite = false;
    double value = 0.;
    if (method == PGC_UNCONSTRAIN_DEATH) { // :150-158
        overwrite = !feasible;
        value = 1.7976931348623157e308;
    } else if (method == PGC_UNCONSTRAIN_KURI) { // :159-179
        overwrite = !feasible;
        value = 1.7976931348623157e308 * (1. - static_cast<double>(sat) / static_cast<double>(nc));
    } else if (method == PGC_UNCONSTRAIN_WEIGHTED) { // :180-204: `!(c <= 0)` so a NaN constraint is penalized too
        for (unsigned k = 0; k < nc; ++k) {
            const double ck = (k < nec ? fabs(c[k]) : c[k]) - tol[k];
            if (!(ck <= 0.)) add += w[k] * ck;
        }
    }
    for (unsigned k = 0; k < nobj; ++k) o[k] = overwrite ? value : (method == PGC_UNCONSTRAIN_WEIGHTED ? fi[k] + add : fi[k]);
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
    if (p->desc.family == PGC_UNCONSTRAIN) {
        double *ftmp = nullptr;
        PGC_CUDA(cudaMallocAsync(&ftmp, sizeof(double) * n * inner->nf(), s));
        int rc = problem_eval_device(inner, d_dvs, n, ftmp, s);
        if (rc == PGC_OK) {
            unconstrain_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(
                ftmp, p->d_meta, d_fvs, n, static_cast<unsigned>(inner->nobj), static_cast<unsigned>(inner->nec),
                static_cast<unsigned>(inner->nic), p->meta_method);
            ctx->launches.fetch_add(1, std::memory_order_relaxed);
            rc = cuda_ok(cudaGetLastError(), "unconstrain_rows_kernel");
        }
        cudaFreeAsync(ftmp, s);
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
    PGC_REQUIRE(family == PGC_UNCONSTRAIN || (a && (family == PGC_TRANSLATE || b)), "meta problem: null vector");
    if (family == PGC_UNCONSTRAIN) {
        // unconstrain.cpp:66-92 (generic_ctor_impl); `a` = weights, len = their number
        const size_t nc = inner->nec + inner->nic;
        PGC_REQUIRE(nc != 0, "Unconstrain can only be applied to constrained problems, the instance of %s is not one.", inner->name.c_str());
        PGC_REQUIRE(!(len != nc && method == PGC_UNCONSTRAIN_WEIGHTED), "Length of weight vector is: %zu while the problem constraints are: %zu",
                    len, nc);
        PGC_REQUIRE(method >= PGC_UNCONSTRAIN_DEATH && method <= PGC_UNCONSTRAIN_IGNORE_O,
                    "The method %d is not supported (did you misspell?)", method);
        PGC_REQUIRE(!(len != 0 && method != PGC_UNCONSTRAIN_WEIGHTED), "The weight vector needs to be empty to use the unconstrain method %d",
                    method);
        PGC_REQUIRE(len == 0 || a, "meta problem: null vector");
    } else if (family == PGC_TRANSLATE) {
        // translate.cpp:83-87
        PGC_REQUIRE(len == inner->nx, "Length of shift vector is: %zu while the problem dimension is: %zu", len, inner->nx);
    } else {
        // decompose.cpp:68-124
        PGC_REQUIRE(inner->nobj >= 2, "Decomposition can only be applied to multi-objective problems");
        PGC_REQUIRE(inner->nec + inner->nic == 0, "Decomposition can only be applied to unconstrained problems, it seems you are trying to "
                    "decompose a problem with %zu constraints", inner->nec + inner->nic);
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
    std::vector<double> host;
    if (a) host.assign(a, a + len);
    if (family == PGC_UNCONSTRAIN) {
        // [c_tol | weights]: the tolerances are the inner problem's at construction (unconstrain copies its inner problem)
        const size_t nc = inner->nec + inner->nic;
        host.assign(inner->c_tol.begin(), inner->c_tol.end());
        host.resize(2 * nc, 0.);
        for (size_t i = 0; i < len; ++i) host[nc + i] = a[i];
        p->nobj = (method == PGC_UNCONSTRAIN_IGNORE_O) ? 1 : inner->nobj; // unconstrain.cpp:269-276
        p->nec = p->nic = 0;
        p->name = inner->name + " [unconstrained]"; // unconstrain.cpp:353-356
        p->flops_per_eval += 4.0 * static_cast<double>(nc);
    } else if (family == PGC_TRANSLATE) {
        // translate.cpp:175-181: the bounds move with the problem
        for (size_t i = 0; i < len; ++i) {
            p->lb[i] = inner->lb[i] + a[i];
            p->ub[i] = inner->ub[i] + a[i];
        }
        p->nobj = inner->nobj;
        p->nec = inner->nec; // translate.cpp: get_nec / get_nic / the tolerances are the inner problem's
        p->nic = inner->nic;
        p->c_tol = inner->c_tol;
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
