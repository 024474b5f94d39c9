// eval_constrained.cu - the constrained UDPs with a device evaluator (what the `unconstrain` meta-problem, SURVEY.md section 8f
// row 1, wraps) and problem::feasibility_f per row.
//
//   hock_schittkowski_71 : nx 4; f = [objective | 1 equality | 1 inequality]      reference src/problems/hock_schittkowski_71.cpp:48-55
//   luksan_vlcek1        : nx = dim >= 3; f = [objective | dim - 2 equalities]     reference src/problems/luksan_vlcek1.cpp:60-77
//   feasibility          : all constraints within the tolerances                   reference src/problem.cpp:709-721,
//                                                                                   include/pagmo/utils/constrained.hpp:49-80
//
// Fitness rows are nf = nobj + nec + nic wide, row-major, as problem::batch_fitness lays them out (problem.cpp:383-410).  One
// thread per output element: element 0 of a row walks the objective's terms in the reference's order (a sequential sum), the
// other elements are one constraint each.  -fmad=false keeps the reference's separate roundings; hock_schittkowski_71 is exact.
#include <cmath>

#include "pgc_internal.cuh"

namespace pgc
{

namespace
{

__device__ __forceinline__ double max0(double a) { return a < 0. ? 0. : a; } // std::max(a, 0.): a NaN stays a NaN (never satisfied)

__global__ void hs71_kernel(const double *__restrict__ xs, double *__restrict__ fs, size_t n)
{
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *x = xs + 4 * i;
    const double x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3];
    double *f = fs + 3 * i;
    f[0] = x0 * x3 * (x0 + x1 + x2) + x2;                    // :51
    f[1] = x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3 - 40.;      // :52
    f[2] = 25. - x0 * x1 * x2 * x3;                          // :53
}

// one thread per (row, element): element 0 = objective (:67-71), element j >= 1 = equality j - 1 (:72-76)
__global__ void luksan_vlcek1_kernel(const double *__restrict__ xs, double *__restrict__ fs, size_t n, unsigned dim)
{
    const unsigned nf = dim - 1u;
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n * nf) return;
    const size_t r = e / nf;
    const unsigned j = static_cast<unsigned>(e - r * nf);
    const double *x = xs + r * dim;
    if (j == 0u) {
        double f0 = 0.;
        double xi = x[0];
        for (unsigned i = 0; i + 1u < dim; ++i) {
            const double xn = x[i + 1];
            const double a1 = xi * xi - xn;
            const double a2 = xi - 1.;
            f0 += 100. * a1 * a1 + a2 * a2;
            xi = xn;
        }
        fs[e] = f0;
    } else {
        const unsigned i = j - 1u;
        const double x0 = x[i], x1 = x[i + 1], x2 = x[i + 2];
        fs[e] = (3. * pow(x1, 3.) + 2. * x2 - 5. + sin(x1 - x2) * sin(x1 + x2) + 4. * x1 - x0 * exp(x0 - x1) - 3.);
    }
}

// problem::feasibility_f: test_eq_constraints / test_ineq_constraints count the constraints whose violation max(|c| - tol, 0)
// (equalities) or max(c - tol, 0) (inequalities) is <= 0; NaN constraints compare false and make the row infeasible.
__global__ void feasibility_kernel(const double *__restrict__ f, const double *__restrict__ tol, unsigned char *__restrict__ out, size_t n,
                                   unsigned nobj, unsigned nec, unsigned nic)
{
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *c = f + i * (nobj + nec + nic) + nobj;
    unsigned sat = 0;
    for (unsigned k = 0; k < nec; ++k) sat += (max0(fabs(c[k]) - tol[k]) <= 0.) ? 1u : 0u;
    for (unsigned k = nec; k < nec + nic; ++k) sat += (max0(c[k] - tol[k]) <= 0.) ? 1u : 0u;
    out[i] = sat == nec + nic ? 1 : 0;
}

} // namespace

int constrained_create(pgc_problem *p)
{
    if (p->desc.family == PGC_HOCK_SCHITTKOWSKI_71) {
        p->nx = 4;
        p->nobj = 1;
        p->nec = 1;
        p->nic = 1;
        p->lb.assign(4, 1.); // hock_schittkowski_71.cpp:64-67
        p->ub.assign(4, 5.);
        p->name = "Hock Schittkowski 71";
        p->flops_per_eval = 18;
        p->transc_per_eval = 0;
    } else if (p->desc.family == PGC_LUKSAN_VLCEK1) {
        const unsigned D = p->desc.dim;
        PGC_REQUIRE(D >= 3u, "luksan_vlcek1 must have minimum 3 dimension, %u requested", D); // luksan_vlcek1.cpp:46-49
        p->nx = D;
        p->nobj = 1;
        p->nec = D - 2u;
        p->nic = 0;
        p->lb.assign(D, -5.); // luksan_vlcek1.cpp:85-88
        p->ub.assign(D, 5.);
        p->name = "luksan_vlcek1";
        p->flops_per_eval = 8.0 * (D - 1) + 14.0 * (D - 2);
        p->transc_per_eval = 4.0 * (D - 2);
    } else {
        set_error("constrained_create: family %d is not a constrained UDP", p->desc.family);
        return PGC_ERR_INVALID_ARGUMENT;
    }
    p->c_tol.assign(p->nec + p->nic, 0.); // problem.cpp:232: tolerances default to zero
    return PGC_OK;
}

int constrained_eval(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream)
{
    if (n == 0) return PGC_OK;
    if (p->desc.family == PGC_HOCK_SCHITTKOWSKI_71) {
        hs71_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(d_dvs, d_fvs, n);
    } else {
        const size_t total = n * (p->nx - 1);
        PGC_REQUIRE(total / 256 < 0x7fffffffull, "luksan_vlcek1: batch too large for one launch");
        luksan_vlcek1_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(d_dvs, d_fvs, n, static_cast<unsigned>(p->nx));
    }
    PGC_CUDA(cudaGetLastError());
    p->ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

int feasibility_rows(pgc_problem *p, const double *d_f, size_t n, unsigned char *d_feasible, cudaStream_t stream)
{
    if (n == 0) return PGC_OK;
    const size_t nc = p->nec + p->nic;
    if (nc == 0) { // an unconstrained problem: every fitness vector is feasible
        PGC_CUDA(cudaMemsetAsync(d_feasible, 1, n, stream));
        return PGC_OK;
    }
    StreamScratch scratch(stream);
    double *d_tol = nullptr;
    PGC_CUDA(scratch.get(&d_tol, sizeof(double) * nc));
    // the tolerances are a host-side attribute of the problem (set_c_tol may change them between calls): pageable copy, the
    // vector is copied into the driver's staging buffer before the call returns
    PGC_CUDA(cudaMemcpyAsync(d_tol, p->c_tol.data(), sizeof(double) * nc, cudaMemcpyHostToDevice, stream));
    feasibility_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(d_f, d_tol, d_feasible, n, static_cast<unsigned>(p->nobj),
                                                                                    static_cast<unsigned>(p->nec), static_cast<unsigned>(p->nic));
    PGC_CUDA(cudaGetLastError());
    p->ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

} // namespace pgc
