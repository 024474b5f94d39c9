// eval_mo.cu - batch fitness of the multi-objective UDPs on sm_100a:
//   zdt1..zdt6   reference src/problems/zdt.cpp:233-356 (ctor checks :54-66, bounds :98-124)
//   dtlz1..dtlz7 reference src/problems/dtlz.cpp:258-409 (g-functions :207-245, h7 :247-256, ctor checks :53-76)
//
// O(D) work per individual and 8*(D + nobj) bytes: HBM-bound.  A CTA stages a tile of individuals (whole rows) in
// shared memory with coalesced loads, then thread t evaluates individual t with the reference's loop order, so the
// sums and products round exactly as the reference's do; what remains are libdevice-vs-glibc ulps of
// sqrt/sin/cos/exp/pow.  Compiled with -fmad=false (the reference build has no FMA contraction).
#include <cmath>

#include "pgc_internal.cuh"

namespace pgc
{

namespace
{

constexpr double kPi = 3.141592653589793238462643383279502884;   // pagmo::detail::pi()
constexpr double kPiHalf = 1.570796326794896619231321691639751442; // pagmo::detail::pi_half()

struct MoParams {
    const double *x;
    double *f;
    long long n;
    int D;      // decision-vector length
    int nobj;   // fitness length
    int prob_id;
    int alpha;  // dtlz4
    int tile;   // individuals per CTA == blockDim.x
    int stride; // shared-memory row stride (odd)
};

// ---- ZDT ------------------------------------------------------------------------------------------------------
__device__ void zdt_eval(int id, const double *x, int N, double *f)
{
    double g = 0.;
    switch (id) {
        case 1: // zdt.cpp:233-247
        case 2: // :249-264
        case 3: // :266-281
            for (int i = 1; i < N; ++i) g += x[i];
            g = 1. + (9. * g) / static_cast<double>(N - 1);
            f[0] = x[0];
            if (id == 1) f[1] = g * (1. - sqrt(x[0] / g));
            else if (id == 2) f[1] = g * (1. - (x[0] / g) * (x[0] / g));
            else f[1] = g * (1. - sqrt(x[0] / g) - x[0] / g * sin(10. * kPi * x[0]));
            return;
        case 4: // :283-298
            g = 1 + 10 * static_cast<double>(N - 1);
            for (int i = 1; i < N; ++i) g += x[i] * x[i] - 10. * cos(4. * kPi * x[i]);
            f[0] = x[0];
            f[1] = g * (1. - sqrt(x[0] / g));
            return;
        case 5: { // :300-343 (integer-valued: the decision vector is rounded first)
            const int n_vectors = (N - 30) / 5 + 1;
            int u0 = 0;
            for (int i = 0; i < 30; ++i) u0 += (round(x[i]) == 1.) ? 1 : 0;
            int k = 30;
            for (int i = 1; i < n_vectors; ++i) {
                int u = 0;
                for (int j = 0; j < 5; ++j) u += (round(x[k++]) == 1.) ? 1 : 0;
                g += static_cast<double>((u < 5) ? 2 + u : 1);
            }
            f[0] = 1.0 + static_cast<double>(u0);
            f[1] = g * (1. / f[0]);
            return;
        }
        default: // 6, :345-360
            f[0] = 1 - exp(-4 * x[0]) * pow(sin(6 * kPi * x[0]), 6.0);
            for (int i = 1; i < N; ++i) g += x[i];
            g = 1 + 9 * pow((g / static_cast<double>(N - 1)), 0.25);
            f[1] = g * (1 - (f[0] / g) * (f[0] / g));
            return;
    }
}

// ---- DTLZ -----------------------------------------------------------------------------------------------------
// g-function over x_M = x[M-1 .. N), dtlz.cpp:182-245
__device__ double dtlz_g(int id, const double *xm, int len)
{
    double y = 0.;
    switch (id) {
        case 1:
        case 3: // g13 :208-215
            for (int i = 0; i < len; ++i) {
                const double d = xm[i] - 0.5;
                y += d * d - cos(20. * kPi * d);
            }
            return 100. * (y + static_cast<double>(len));
        case 2:
        case 4:
        case 5: // g245 :217-224
            for (int i = 0; i < len; ++i) {
                const double d = xm[i] - 0.5;
                y += d * d;
            }
            return y;
        case 6: // g6 :226-233
            for (int i = 0; i < len; ++i) y += pow(xm[i], 0.1);
            return y;
        default: // g7 :235-245
            for (int i = 0; i < len; ++i) y += xm[i];
            return (9. / static_cast<double>(len)) * y;
    }
}

// angle of variable j for the spherical shape functions of dtlz2..6
__device__ __forceinline__ double dtlz_angle(int id, const double *x, int j, double g, double alpha)
{
    if (id == 4) return pow(x[j], alpha) * kPiHalf;                        // :324
    if (id == 5 || id == 6) {                                              // theta, :351-358
        if (j == 0) return x[0] * kPiHalf;
        const double t = 1. / (2. * (1. + g));
        return (t + ((g * x[j]) / (1.0 + g))) * kPiHalf;
    }
    return x[j] * kPiHalf;                                                  // :296
}

__device__ void dtlz_eval(int id, const double *x, int N, int M, int alpha, double *f)
{
    const double g = dtlz_g(id, x + (M - 1), N - (M - 1));
    if (id == 1) { // f1_objfun_impl :258-284
        f[0] = 0.5 * (1. + g);
        for (int i = 0; i < M - 1; ++i) f[0] *= x[i];
        for (int i = 1; i < M - 1; ++i) {
            double v = 0.5 * (1.0 + g);
            for (int j = 0; j < M - (i + 1); ++j) v *= x[j];
            v *= 1. - x[M - (i + 1)];
            f[i] = v;
        }
        f[M - 1] = 0.5 * (1. - x[0]) * (1. + g);
        return;
    }
    if (id == 7) { // f7_objfun_impl :383-403, h7 :247-256 (g already holds 1 + g7)
        const double g1 = 1. + g;
        double y = 0.;
        for (int i = 0; i < M - 1; ++i) {
            f[i] = x[i];
            y += (x[i] / (1.0 + g1)) * (1.0 + sin(3 * kPi * x[i]));
        }
        f[M - 1] = (1. + g1) * (static_cast<double>(M) - y);
        return;
    }
    // f23 :286-311, f4 :313-339, f56 :341-381: same structure, different angle
    const double a = static_cast<double>(alpha);
    double v = (1. + g);
    for (int i = 0; i < M - 1; ++i) v *= cos(dtlz_angle(id, x, i, g, a));
    f[0] = v;
    for (int i = 1; i < M - 1; ++i) {
        v = (1. + g);
        for (int j = 0; j < M - (i + 1); ++j) v *= cos(dtlz_angle(id, x, j, g, a));
        v *= sin(dtlz_angle(id, x, M - (i + 1), g, a));
        f[i] = v;
    }
    f[M - 1] = (1. + g) * sin(dtlz_angle(id, x, 0, g, a));
}

template <int FAM> __global__ void mo_kernel(const MoParams P)
{
    extern __shared__ double tile[];
    const int tid = threadIdx.x, T = P.tile, S = P.stride, D = P.D;
    double *fout = tile + static_cast<size_t>(T) * S; // [T][nobj] staging for coalesced stores
    const long long ntiles = (P.n + T - 1) / T;
    for (long long tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
        const long long t0 = tb * T;
        const int nt = (P.n - t0 < T) ? static_cast<int>(P.n - t0) : T;
        const double *src = P.x + t0 * D;
        for (int e = tid; e < nt * D; e += T) {
            const int t = e / D, j = e - t * D;
            tile[t * S + j] = __ldcs(src + e);
        }
        __syncthreads();
        if (tid < nt) {
            double *f = fout + tid * P.nobj;
            if (FAM == PGC_ZDT) zdt_eval(P.prob_id, tile + tid * S, D, f);
            else dtlz_eval(P.prob_id, tile + tid * S, D, P.nobj, P.alpha, f);
        }
        __syncthreads();
        for (int e = tid; e < nt * P.nobj; e += T) P.f[t0 * P.nobj + e] = fout[e];
        __syncthreads();
    }
}

} // namespace

int mo_create(pgc_problem *p)
{
    const pgc_problem_desc &d = p->desc;
    if (d.family == PGC_ZDT) {
        const unsigned param = d.dim;
        PGC_REQUIRE(param >= 2u,
                    "ZDT test problems must have a minimum value of 2 for the constructing parameter (representing the "
                    "dimension except for ZDT5), %u requested",
                    param); // zdt.cpp:56-60
        PGC_REQUIRE(d.prob_id >= 1u && d.prob_id <= 6u,
                    "ZDT test suite contains six (prob_id=[1 ... 6]) problems, prob_id=%u was detected", d.prob_id);
        const size_t D = (d.prob_id == 5u) ? 30u + 5u * (param - 1u) : param; // :115-118
        p->nx = D;
        p->nobj = 2;
        p->lb.assign(D, 0.);
        p->ub.assign(D, 1.);
        if (d.prob_id == 4u) { // :104-111
            p->lb.assign(D, -5.);
            p->ub.assign(D, 5.);
            p->lb[0] = 0.0;
            p->ub[0] = 1.0;
        }
        p->name = "ZDT" + std::to_string(d.prob_id); // :161-164
        p->flops_per_eval = 2.0 * D + 8;
        p->transc_per_eval = (d.prob_id == 4u) ? static_cast<double>(D) : 2.0;
        return PGC_OK;
    }
    // DTLZ, dtlz.cpp:53-76
    PGC_REQUIRE(d.prob_id >= 1u && d.prob_id <= 7u,
                "DTLZ test suite contains seven (prob_id = [1 ... 7]) problems, prob_id=%u was detected", d.prob_id);
    PGC_REQUIRE(d.nobj >= 2u, "DTLZ test problem have a minimum of 2 objectives: fdim=%u was detected", d.nobj);
    PGC_REQUIRE(d.dim > d.nobj, "The problem dimension has to be larger than the number of objectives.");
    p->nx = d.dim;
    p->nobj = d.nobj;
    p->lb.assign(d.dim, 0.);
    p->ub.assign(d.dim, 1.);
    p->name = "DTLZ" + std::to_string(d.prob_id); // dtlz.cpp:170-173
    p->flops_per_eval = 4.0 * d.dim + 3.0 * d.nobj * d.nobj;
    p->transc_per_eval = static_cast<double>(d.nobj) * d.nobj / 2 + ((d.prob_id == 1u || d.prob_id == 3u) ? d.dim : 0);
    return PGC_OK;
}

int mo_eval(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream)
{
    if (n == 0) return PGC_OK;
    MoParams mp;
    mp.x = d_dvs;
    mp.f = d_fvs;
    mp.n = static_cast<long long>(n);
    mp.D = static_cast<int>(p->nx);
    mp.nobj = static_cast<int>(p->nobj);
    mp.prob_id = static_cast<int>(p->desc.prob_id);
    mp.alpha = static_cast<int>(p->desc.param);
    mp.stride = mp.D | 1;
    const size_t limit = 200 * 1024;
    int tile = 128;
    while (tile > 32 && sizeof(double) * tile * (mp.stride + mp.nobj) > limit) tile >>= 1;
    const size_t smem = sizeof(double) * tile * (mp.stride + mp.nobj);
    if (smem > limit) {
        set_error("multi-objective evaluator: decision vectors of length %d do not fit the shared-memory tile", mp.D);
        return PGC_ERR_UNSUPPORTED;
    }
    mp.tile = tile;
    auto kern = (p->desc.family == PGC_ZDT) ? mo_kernel<PGC_ZDT> : mo_kernel<PGC_DTLZ>;
    PGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(limit)));
    const long long ntiles = (mp.n + tile - 1) / tile;
    const long long per_sm = std::max<long long>(1, std::min<long long>(16, (220 * 1024) / static_cast<long long>(smem + 1024)));
    long long blocks = std::min<long long>(ntiles, static_cast<long long>(p->ctx->sm_count) * per_sm);
    kern<<<static_cast<unsigned>(blocks), tile, smem, stream>>>(mp);
    PGC_CUDA(cudaGetLastError());
    p->ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

} // namespace pgc
