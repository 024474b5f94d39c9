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


// ---- WFG1..9, wfg.cpp:304-1066 (shape functions :161-222, transformations :225-302) ---------------------------------------
// One thread per individual, the working vector y in place in the thread's shared-memory row; t / par (M values) in the
// thread's staging rows.  Every problem is: normalise -> element-wise transformations -> reduction to M values -> shapes.
__device__ __forceinline__ double wfg_s_linear(double y, double a) { return fabs(y - a) / (fabs(floor(a - y) + a)); }
__device__ __forceinline__ double wfg_b_flat(double y, double a, double b, double c)
{
    return a + fmin(0.0, floor(y - b)) * a * (b - y) / (b)-fmin(0.0, floor(c - y)) * (1.0 - a) * (y - c) / (1 - c);
}
__device__ __forceinline__ double wfg_b_param(double y, double u, double a, double b, double c)
{
    const double v = a - (1.0 - 2 * u) * fabs(floor(0.5 - u) + a);
    return pow(y, b + (c - b) * v);
}
__device__ __forceinline__ double wfg_s_decept(double y, double a, double b, double c)
{
    return 1.0
           + (fabs(y - a) - b)
                 * ((floor(y - a + b) * (1.0 - c + (a - b) / b)) / (a - b) + (floor(a + b - y) * (1.0 - c + (1.0 - a - b) / b)) / (1.0 - a - b)
                    + 1.0 / b);
}
__device__ __forceinline__ double wfg_s_multi(double y, double a, double b, double c)
{
    const double r = fabs(y - c) / (2.0 * (floor(c - y) + c));
    return (1 + cos((4.0 * a + 2.0) * kPi * (0.5 - r)) + 4.0 * b * (r * r)) / (b + 2.0);
}
__device__ double wfg_r_sum(const double *y, int lo, int hi, bool weighted)
{
    double g1 = 0., g2 = 0.;
    for (int j = lo; j < hi; ++j) {
        const double w = weighted ? 2. * (static_cast<double>(j) + 1) : 1.0;
        g1 += w * y[j];
        g2 += w;
    }
    return g1 / g2;
}
__device__ double wfg_r_nonsep(const double *y, int lo, int hi, int A)
{
    if (A == 1) return wfg_r_sum(y, lo, hi, false);
    const int len = hi - lo;
    double g = 0.;
    for (int j = 0; j < len; ++j) {
        g += y[lo + j];
        for (int i = 0; i <= A - 2; ++i) g += fabs(y[lo + j] - y[lo + (1 + j + i) % len]);
    }
    // :297-300: ceil(A / 2) is taken on the INTEGER quotient in the reference
    return g / (static_cast<double>(len) / static_cast<double>(A) * static_cast<double>(A / 2)
                * (1.0 + 2.0 * static_cast<double>(A) - 2.0 * ceil(static_cast<double>(A) / 2.0)));
}
// shape: 0 convex, 1 linear, 2 concave (m = 1..M)
__device__ double wfg_shape(int shape, const double *p, int m, int M)
{
    double g = 1.;
    const int lim = (m == 1) ? M - 1 : M - m;
    if (shape == 1) {
        if (m == M) return 1.0 - p[0];
        for (int i = 0; i < lim; ++i) g *= p[i];
        return m == 1 ? g : g * (1.0 - p[M - m]);
    }
    if (shape == 0) {
        for (int i = 0; i < lim; ++i) g *= 1.0 - cos(p[i] * kPi / 2.0);
        return m == 1 ? g : g * (1 - sin(p[M - m] * kPi / 2.0));
    }
    if (m == M) return cos(p[0] * kPi / 2.0);
    for (int i = 0; i < lim; ++i) g *= sin(p[i] * kPi / 2.0);
    return m == 1 ? g : g * cos(p[M - m] * kPi / 2.0);
}

// y: the individual's row (overwritten), par: M scratch values, f: M outputs
__device__ void wfg_eval(int id, double *y, int n, int M, int k, double *par, double *f)
{
    const int l = n - k;
    for (int i = 0; i < n; ++i) y[i] = y[i] / (2.0 * (static_cast<double>(i) + 1)); // get_bounds().second[i], :138-146
    int red_n = n;
    switch (id) {
        case 1: // :326-352
            for (int i = k; i < n; ++i) y[i] = wfg_b_flat(wfg_s_linear(y[i], 0.35), 0.8, 0.75, 0.85);
            for (int i = 0; i < n; ++i) y[i] = pow(y[i], 0.02);
            break;
        case 2:
        case 3: // :418-444, :515-541
            for (int i = k; i < n; ++i) y[i] = wfg_s_linear(y[i], 0.35);
            for (int i = k + 1; i <= k + l / 2; ++i) {
                const int head = k + 2 * (i - k) - 2;
                y[i - 1] = wfg_r_nonsep(y, head, head + 2, 2);
            }
            red_n = k + l / 2;
            break;
        case 4: // :611-614
            for (int i = 0; i < n; ++i) y[i] = wfg_s_multi(y[i], 30.0, 10.0, 0.35);
            break;
        case 5: // :681-684
            for (int i = 0; i < n; ++i) y[i] = wfg_s_decept(y[i], 0.35, 0.001, 0.05);
            break;
        case 6: // :754-761
            for (int i = k; i < n; ++i) y[i] = wfg_s_linear(y[i], 0.35);
            break;
        case 7: // :819-843; position i-1 reads x_norm[i-1 ...] only, so in place is safe
            for (int i = 1; i <= k; ++i) y[i - 1] = wfg_b_param(y[i - 1], wfg_r_sum(y, i, n, false), 0.98 / 49.98, 0.02, 50);
            for (int i = k; i < n; ++i) y[i] = wfg_s_linear(y[i], 0.35);
            break;
        case 8: // :905-928: position i uses the transformed prefix
            for (int i = k; i < n; ++i) y[i] = wfg_b_param(y[i], wfg_r_sum(y, 0, i, false), 0.98 / 49.98, 0.02, 50);
            for (int i = k; i < n; ++i) y[i] = wfg_s_linear(y[i], 0.35);
            break;
        default: // 9, :992-1016
            for (int i = 0; i + 1 < n; ++i) y[i] = wfg_b_param(y[i], wfg_r_sum(y, i + 1, n, false), 0.98 / 49.98, 0.02, 50);
            for (int i = 0; i < n; ++i) y[i] = i < k ? wfg_s_decept(y[i], 0.35, 0.001, 0.05) : wfg_s_multi(y[i], 30.0, 95.0, 0.35);
            break;
    }
    const bool nonsep = id == 6 || id == 9, weighted = id == 1;
    for (int i = 1; i <= M - 1; ++i) {
        const int head = (i - 1) * k / (M - 1), tail = i * k / (M - 1);
        par[i - 1] = nonsep ? wfg_r_nonsep(y, head, tail, k / (M - 1)) : wfg_r_sum(y, head, tail, weighted);
    }
    const double t_last = nonsep ? wfg_r_nonsep(y, k, n, l) : wfg_r_sum(y, k, red_n, weighted);
    for (int i = 0; i < M - 1; ++i) {
        const double lo = (id == 3 && i > 0) ? 0.0 : 1.0; // WFG3, :568-575
        par[i] = fmax(t_last, lo) * (par[i] - 0.5) + 0.5;
    }
    par[M - 1] = t_last;
    for (int i = 0; i < M; ++i) {
        double sh;
        if (id <= 2 && i + 1 == M) {
            const double p0 = par[0];
            if (id == 1) sh = pow((1.0 - p0 - cos(2 * 5.0 * kPi * p0 + kPi / 2.0) / (2.0 * 5.0 * kPi)), 1.0); // mixed(p0, 1, 5) :199-206
            else {
                const double c = cos(5.0 * pow(p0, 1.0) * kPi);                                               // disconnected(p0, 1, 1, 5) :208-214
                sh = 1.0 - pow(p0, 1.0) * (c * c);
            }
        } else
            sh = wfg_shape(id <= 2 ? 0 : (id == 3 ? 1 : 2), par, i + 1, M);
        f[i] = t_last + 2.0 * (static_cast<double>(i) + 1) * sh;
    }
}

template <int FAM> __global__ void mo_kernel(const MoParams P)
{
    extern __shared__ double tile[];
    const int tid = threadIdx.x, T = P.tile, S = P.stride, D = P.D;
    double *fout = tile + static_cast<size_t>(T) * S; // [T][nobj] staging for coalesced stores
    double *scratch = fout + static_cast<size_t>(T) * P.nobj; // WFG: [T][nobj] parameters
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
            else if (FAM == PGC_WFG) wfg_eval(P.prob_id, tile + tid * S, D, P.nobj, P.alpha, scratch + tid * P.nobj, f);
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
        p->nix = d.prob_id == 5u ? D : 0u; // zdt.cpp:126-143: zdt5 is integer valued
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
    if (d.family == PGC_WFG) { // wfg.cpp:64-93
        const unsigned k = d.param;
        PGC_REQUIRE(d.prob_id >= 1u && d.prob_id <= 9u, "WFG test suite contains nine (prob_id=[1 ... 9]) problems, prob_id=%u was detected",
                    d.prob_id);
        PGC_REQUIRE(d.dim >= 1u, "WFG problem suite must have minimum 1 dimension for the decision vector, %u requested", d.dim);
        PGC_REQUIRE(d.nobj >= 2u, "WFG test problems must have a minimum value of 2 for the objective vector dimension, %u requested", d.nobj);
        PGC_REQUIRE(k < d.dim && k >= 1u && k % (d.nobj - 1u) == 0u,
                    "WFG test problems must have a dim_k parameter which is within [1,dim_dvs), and such that dim_k mod(dim_obj-1) == 0 %u "
                    "requested",
                    k);
        PGC_REQUIRE(!(d.prob_id == 2u || d.prob_id == 3u) || (d.dim - k) % 2u == 0u,
                    "For problems WFG2 and WFG3 the dim_k parameter and the decision vector size must satisfy (dim_dvs-dim_k) mod(2)=0%u was "
                    "detected",
                    (d.dim - k) % 2u);
        p->nx = d.dim;
        p->nobj = d.nobj;
        p->lb.assign(d.dim, 0.);
        p->ub.resize(d.dim);
        for (unsigned i = 0; i < d.dim; ++i) p->ub[i] = 2.0 * (static_cast<double>(i) + 1); // :138-146
        p->name = "WFG" + std::to_string(d.prob_id);                                       // :149-152
        p->flops_per_eval = 30.0 * d.dim + 10.0 * d.nobj * d.nobj;
        p->transc_per_eval = 2.0 * d.dim + d.nobj * d.nobj;
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
    while (tile > 32 && sizeof(double) * tile * (mp.stride + 2 * mp.nobj) > limit) tile >>= 1;
    const size_t smem = sizeof(double) * tile * (mp.stride + 2 * mp.nobj);
    if (smem > limit) {
        set_error("multi-objective evaluator: decision vectors of length %d do not fit the shared-memory tile", mp.D);
        return PGC_ERR_UNSUPPORTED;
    }
    mp.tile = tile;
    auto kern = (p->desc.family == PGC_ZDT) ? mo_kernel<PGC_ZDT> : (p->desc.family == PGC_WFG) ? mo_kernel<PGC_WFG> : mo_kernel<PGC_DTLZ>;
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
