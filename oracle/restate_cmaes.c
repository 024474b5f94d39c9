/* oracle/restate_cmaes.c - plain-C restatement of the dense contractions of CMA-ES / xNES.  TEST INFRASTRUCTURE ONLY.
 * Follows reference src/algorithms/cmaes.cpp: sampling :246-253 (newpop[i] = mean + sigma * B * D * tmp), recombination :362-366
 * (mean = elite[0] * w0; mean += elite[i] * w_i), rank-mu :375-380 (C = (e0 - m)(e0 - m)^T w0; C += ...; C /= sigma * sigma), and
 * src/algorithms/xnes.cpp:302-305 (cov_grad = u0 (z0 z0^T - I) + ...; the "- I" part is the caller's).
 * PARITY UNPINNED: cmaes/xnes need Eigen 3.3 (CMakeLists.txt:182), which is not in this image, so the reference algorithms cannot
 * be compiled here and the reference tests (tests/cmaes.cpp) only check determinism and exit conditions.  The loops below are the
 * same sums in the same order with plain doubles (Eigen evaluates these expressions coefficient-wise without reassociation).
 * Normal draws: Box-Muller on the Philox stream (seed, TAG_CMAES, generation, i, 2j / 2j+1), as the device does.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "philox.h"

#define ROW(i) (rows + (size_t)(idx ? idx[i] : (i)) * D)

int oracle_weighted_mean(const double *rows, const uint32_t *idx, const double *w, size_t k, size_t D, double *out)
{
    if (!k) return -1;
    for (size_t a = 0; a < D; ++a) out[a] = ROW(0)[a] * w[0];
    for (size_t i = 1; i < k; ++i)
        for (size_t a = 0; a < D; ++a) out[a] += ROW(i)[a] * w[i];
    return 0;
}

int oracle_weighted_gram(const double *rows, const uint32_t *idx, const double *center, const double *w, size_t k, size_t D, double scale_div,
                         double *out)
{
    if (!k) return -1;
    double *d = (double *)malloc(D * sizeof(double));
    for (size_t i = 0; i < k; ++i) {
        for (size_t a = 0; a < D; ++a) d[a] = ROW(i)[a] - (center ? center[a] : 0.0);
        for (size_t a = 0; a < D; ++a)
            for (size_t b = 0; b < D; ++b) {
                const double t = d[a] * d[b] * w[i];
                out[a * D + b] = i ? out[a * D + b] + t : t;
            }
    }
    for (size_t e = 0; e < D * D; ++e) out[e] /= scale_div;
    free(d);
    return 0;
}

int oracle_cmaes_sample(const double *mean, const double *bd, double sigma, size_t lambda, size_t D, uint64_t seed, uint32_t generation,
                        double *z, double *x)
{
    double *zi = (double *)malloc(D * sizeof(double));
    for (size_t i = 0; i < lambda; ++i) {
        for (size_t j = 0; j < D; ++j) {
            const double u1 = 1.0 - oracle_philox_u01(seed, ORACLE_TAG_CMAES, generation, (uint32_t)i, (uint32_t)(2 * j));
            const double u2 = oracle_philox_u01(seed, ORACLE_TAG_CMAES, generation, (uint32_t)i, (uint32_t)(2 * j + 1));
            zi[j] = sqrt(-2.0 * log(u1)) * cos(2.0 * 3.141592653589793238462643383279502884 * u2);
            if (z) z[i * D + j] = zi[j];
        }
        for (size_t a = 0; a < D; ++a) {
            double y = 0.0;
            for (size_t j = 0; j < D; ++j) y += bd[a * D + j] * zi[j];
            x[i * D + a] = mean[a] + sigma * y;
        }
    }
    free(zi);
    return 0;
}

/* ---- cmaes::evolve, src/algorithms/cmaes.cpp:111-407 (memory = false), with the Philox normals the device draws ----
 * Statement by statement after the reference; Eigen's SelfAdjointEigenSolver (absent here) is stood in for by the cyclic
 * Jacobi method (Golub & Van Loan 8.5: rotations with t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)) over all pairs p < q until
 * the off-diagonal mass vanishes), eigenvalues ascending, each eigenvector with its largest component positive.  PARITY UNPINNED
 * for that step (the reference cannot be compiled without Eigen); everything else is the reference's arithmetic. */
static void jacobi_eigen(double *a, size_t D, double *w, double *v)
{
    for (size_t i = 0; i < D * D; ++i) v[i] = 0.;
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
                const double t = (theta >= 0. ? 1. : -1.) / (fabs(theta) + sqrt(theta * theta + 1.));
                const double c = 1. / sqrt(t * t + 1.), sn = t * c;
                for (size_t k = 0; k < D; ++k) {
                    const double akp = a[k * D + p], akq = a[k * D + q];
                    a[k * D + p] = c * akp - sn * akq;
                    a[k * D + q] = sn * akp + c * akq;
                }
                for (size_t k = 0; k < D; ++k) {
                    const double apk = a[p * D + k], aqk = a[q * D + k];
                    a[p * D + k] = c * apk - sn * aqk;
                    a[q * D + k] = sn * apk + c * aqk;
                }
                for (size_t k = 0; k < D; ++k) {
                    const double vkp = v[k * D + p], vkq = v[k * D + q];
                    v[k * D + p] = c * vkp - sn * vkq;
                    v[k * D + q] = sn * vkp + c * vkq;
                }
            }
    }
    size_t *order = (size_t *)malloc(D * sizeof(size_t));
    for (size_t i = 0; i < D; ++i) order[i] = i;
    for (size_t i = 1; i < D; ++i) { /* stable insertion sort by eigenvalue */
        const size_t o = order[i];
        size_t j = i;
        while (j > 0 && a[o * D + o] < a[order[j - 1] * D + order[j - 1]]) { order[j] = order[j - 1]; --j; }
        order[j] = o;
    }
    double *vs = (double *)malloc(D * D * sizeof(double));
    for (size_t j = 0; j < D; ++j) {
        const size_t src = order[j];
        w[j] = a[src * D + src];
        size_t big = 0;
        for (size_t k = 1; k < D; ++k)
            if (fabs(v[k * D + src]) > fabs(v[big * D + src])) big = k;
        const double sgn = v[big * D + src] < 0. ? -1. : 1.;
        for (size_t k = 0; k < D; ++k) vs[k * D + j] = sgn * v[k * D + src];
    }
    memcpy(v, vs, D * D * sizeof(double));
    free(order); free(vs);
}

static int less_nan_last(double a, double b) { return !isnan(a) && (isnan(b) || a < b); }

int oracle_cmaes_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t lam, size_t D,
                        unsigned gens, double cc, double cs, double c1, double cmu, double sigma0, double ftol, double xtol, int force_bounds,
                        uint64_t seed, uint32_t first_generation, unsigned *gens_done, double *sigma_out)
{
    const size_t mu = lam / 2u;
    if (gens_done) *gens_done = 0;
    if (lam < 5u) return -1;
    if (gens == 0) return 0;
    const double N = (double)D;
    double *weights = (double *)malloc(mu * sizeof(double)), wsum = 0., w2 = 0.;
    for (size_t i = 0; i < mu; ++i) { weights[i] = log((double)mu + 0.5) - log((double)i + 1.); wsum += weights[i]; }
    for (size_t i = 0; i < mu; ++i) { weights[i] /= wsum; w2 += weights[i] * weights[i]; }
    const double mueff = 1. / w2;
    if (cc == -1) cc = (4. + mueff / N) / (N + 4. + 2. * mueff / N);
    if (cs == -1) cs = (mueff + 2.) / (N + mueff + 5.);
    if (c1 == -1) c1 = 2. / ((N + 1.3) * (N + 1.3) + mueff);
    if (cmu == -1) cmu = 2. * (mueff - 2. + 1. / mueff) / ((N + 2.) * (N + 2.) + mueff);
    const double damps = 1. + 2. * fmax(0., sqrt((mueff - 1.) / (N + 1.)) - 1.) + cs;
    const double chiN = sqrt(N) * (1. - 1. / (4. * N) + 1. / (21. * N * N));
#define VEC(n) ((double *)calloc((n), sizeof(double)))
    double *mean = VEC(D), *meanold = VEC(D), *pc = VEC(D), *ps = VEC(D), *dvec = VEC(D), *C = VEC(D * D), *Cold = VEC(D * D), *Cmu = VEC(D * D),
           *B = VEC(D * D), *invsqrtC = VEC(D * D), *BD = VEC(D * D), *z = VEC(lam * D), *xn = VEC(lam * D), *fn = VEC(lam), *tmp = VEC(D),
           *work = VEC(D * D), *wv = VEC(D), *V = VEC(D * D);
    uint32_t *idx = (uint32_t *)malloc(lam * sizeof(uint32_t));
    size_t ib = 0, iw = 0;
#define BEST_WORST()                                                                                                   \
    do {                                                                                                               \
        ib = iw = 0;                                                                                                   \
        for (size_t i = 1; i < lam; ++i) {                                                                             \
            if (f[i] < f[ib]) ib = i;                                                                                  \
            if (f[i] > f[iw]) iw = i;                                                                                  \
        }                                                                                                              \
    } while (0)
    BEST_WORST();
    double sigma = sigma0;
    memcpy(mean, x + ib * D, D * sizeof(double));
    for (size_t j = 0; j < D; ++j) {
        dvec[j] = fmax(ub[j] - lb[j], 1e-6);
        B[j * D + j] = 1.;
        C[j * D + j] = dvec[j] * dvec[j];
        invsqrtC[j * D + j] = 1. / dvec[j];
    }
    unsigned long long counteval = 0, eigeneval = 0;
    unsigned done = 0;
    int rc = 0;
    for (unsigned g = 0; g < gens && !rc; ++g) {
        const uint32_t generation = first_generation + g;
        for (size_t a = 0; a < D; ++a)
            for (size_t j = 0; j < D; ++j) BD[a * D + j] = B[a * D + j] * dvec[j];
        oracle_cmaes_sample(mean, BD, sigma, lam, D, seed, generation, z, xn);
        double nrm = 0.;
        for (size_t a = 0; a < D; ++a) {
            double y = 0.;
            for (size_t j = 0; j < D; ++j) y += BD[a * D + j] * z[(lam - 1) * D + j];
            nrm += (sigma * y) * (sigma * y);
        }
        if (sqrt(nrm) < xtol) break;
        BEST_WORST();
        if (fabs(f[ib] - f[iw]) < ftol) break;
        if (force_bounds)
            for (size_t i = 0; i < lam; ++i)
                for (size_t j = 0; j < D; ++j) {
                    if (xn[i * D + j] < lb[j]) xn[i * D + j] = lb[j];
                    else if (xn[i * D + j] > ub[j]) xn[i * D + j] = ub[j];
                }
        if ((rc = oracle_problem_eval(prob, xn, lam, fn))) break;
        memcpy(x, xn, lam * D * sizeof(double));
        memcpy(f, fn, lam * sizeof(double));
        counteval += lam;
        ++done;
        for (size_t i = 0; i < lam; ++i) idx[i] = (uint32_t)i;
        for (size_t i = 1; i < lam; ++i) { /* stable insertion sort by fitness, NaN last */
            const uint32_t o = idx[i];
            size_t j = i;
            while (j > 0 && less_nan_last(f[o], f[idx[j - 1]])) { idx[j] = idx[j - 1]; --j; }
            idx[j] = o;
        }
        memcpy(meanold, mean, D * sizeof(double));
        oracle_weighted_mean(x, idx, weights, mu, D, mean);
        oracle_weighted_gram(x, idx, meanold, weights, mu, D, sigma * sigma, Cmu);
        for (size_t a = 0; a < D; ++a) {
            double y = 0.;
            for (size_t j = 0; j < D; ++j) y += invsqrtC[a * D + j] * (mean[j] - meanold[j]);
            tmp[a] = y;
        }
        double ps2 = 0.;
        for (size_t a = 0; a < D; ++a) {
            ps[a] = (1. - cs) * ps[a] + sqrt(cs * (2. - cs) * mueff) * tmp[a] / sigma;
            ps2 += ps[a] * ps[a];
        }
        const double hsig = (ps2 / N / (1. - pow((1. - cs), (2. * (double)counteval / (double)lam)))) < (2. + 4. / (N + 1.)) ? 1. : 0.;
        for (size_t a = 0; a < D; ++a) pc[a] = (1. - cc) * pc[a] + hsig * sqrt(cc * (2. - cc) * mueff) * (mean[a] - meanold[a]) / sigma;
        memcpy(Cold, C, D * D * sizeof(double));
        for (size_t a = 0; a < D; ++a)
            for (size_t b = 0; b < D; ++b)
                C[a * D + b] = (1. - c1 - cmu) * Cold[a * D + b] + cmu * Cmu[a * D + b]
                               + c1 * ((pc[a] * pc[b]) + (1. - hsig) * cc * (2. - cc) * Cold[a * D + b]);
        sigma *= exp(fmin(0.6, (cs / damps) * (sqrt(ps2) / chiN - 1.)));
        if ((double)(counteval - eigeneval) > ((double)lam / (c1 + cmu) / N / 10.)) {
            eigeneval = counteval;
            for (size_t a = 0; a < D; ++a)
                for (size_t b = a + 1; b < D; ++b) C[a * D + b] = C[b * D + a] = (C[a * D + b] + C[b * D + a]) / 2.;
            memcpy(work, C, D * D * sizeof(double));
            jacobi_eigen(work, D, wv, V);
            memcpy(B, V, D * D * sizeof(double));
            for (size_t j = 0; j < D; ++j) dvec[j] = sqrt(fmax(1e-20, wv[j]));
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
    free(weights); free(mean); free(meanold); free(pc); free(ps); free(dvec); free(C); free(Cold); free(Cmu); free(B); free(invsqrtC); free(BD);
    free(z); free(xn); free(fn); free(tmp); free(work); free(wv); free(V); free(idx);
    return rc;
}

/* ---- xnes::evolve, src/algorithms/xnes.cpp:96-303 (memory = false), with the Philox normals the device draws ------------------
 * Statement by statement after the reference: utilities u (:146-157), A = diag(max(ub - lb, 1e-6) * sigma) and the mean at the best
 * individual (:160-175), per generation the lam samples x = mean + A z with their evaluation (pop.set_x, :196-216), the exit tests on
 * ||A z_0|| and on the NEW population's fitness spread (:219-236), the natural gradients d_center = sum u_i z_(i), cov_grad = sum u_i
 * (z_(i) z_(i)^T - I) over the fitness-sorted samples, and the updates mean += eta_mu A d_center, A <- A exp(d_A), sigma (:264-291).
 * Eigen's matrix exponential (absent here) is stood in for by exp of the symmetric d_A through the Jacobi eigendecomposition above:
 * PARITY UNPINNED for that step, like cmaes' eigendecomposition. */
int oracle_xnes_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t lam, size_t D, unsigned gens,
                       double eta_mu, double eta_sigma, double eta_b, double sigma0, double ftol, double xtol, int force_bounds, uint64_t seed,
                       uint32_t first_generation, unsigned *gens_done, double *sigma_out)
{
    if (gens_done) *gens_done = 0;
    if (lam < 4u) return -1;
    if (gens == 0) return 0;
    const double dim_d = (double)D, lam_d = (double)lam;
    if (eta_mu == -1) eta_mu = 1.;
    const double common_default = 0.6 * (3. + log(dim_d)) / (dim_d * sqrt(dim_d));
    if (eta_sigma == -1) eta_sigma = common_default;
    if (eta_b == -1) eta_b = common_default;
    double *u = (double *)malloc(lam * sizeof(double)), sum = 0.;
    for (size_t i = 0; i < lam; ++i) u[i] = fmax(0., log(lam_d / 2. + 1.) - log((double)(i + 1)));
    for (size_t i = 0; i < lam; ++i) sum += u[i];
    for (size_t i = 0; i < lam; ++i) u[i] = u[i] / sum - 1. / lam_d;
    double sigma = sigma0 == -1 ? 0.5 : sigma0;
    double *A = VEC(D * D), *mean = VEC(D), *z = VEC(lam * D), *xn = VEC(lam * D), *fn = VEC(lam), *dc = VEC(D), *G = VEC(D * D), *dA = VEC(D * D),
           *work = VEC(D * D), *wv = VEC(D), *V = VEC(D * D), *E = VEC(D * D), *An = VEC(D * D), *tmp = VEC(D);
    uint32_t *idx = (uint32_t *)malloc(lam * sizeof(uint32_t));
    for (size_t j = 0; j < D; ++j) A[j * D + j] = fmax(ub[j] - lb[j], 1e-6) * sigma;
    size_t ib = 0, iw = 0;
    for (size_t i = 1; i < lam; ++i)
        if (f[i] < f[ib]) ib = i;
    memcpy(mean, x + ib * D, D * sizeof(double));
    unsigned done = 0;
    int rc = 0;
    for (unsigned g = 0; g < gens && !rc; ++g) {
        const uint32_t generation = first_generation + g;
        oracle_cmaes_sample(mean, A, 1.0, lam, D, seed, generation, z, xn);
        if (force_bounds)
            for (size_t i = 0; i < lam; ++i)
                for (size_t j = 0; j < D; ++j) {
                    if (xn[i * D + j] < lb[j]) xn[i * D + j] = lb[j];
                    else if (xn[i * D + j] > ub[j]) xn[i * D + j] = ub[j];
                }
        if ((rc = oracle_problem_eval(prob, xn, lam, fn))) break;
        memcpy(x, xn, lam * D * sizeof(double));
        memcpy(f, fn, lam * sizeof(double));
        ++done;
        double nrm = 0.;
        for (size_t a = 0; a < D; ++a) {
            double y = 0.;
            for (size_t j = 0; j < D; ++j) y += A[a * D + j] * z[j];
            nrm += y * y;
        }
        if (sqrt(nrm) < xtol) break;
        ib = iw = 0;
        for (size_t i = 1; i < lam; ++i) {
            if (f[i] < f[ib]) ib = i;
            if (f[i] > f[iw]) iw = i;
        }
        if (fabs(f[ib] - f[iw]) < ftol) break;
        for (size_t i = 0; i < lam; ++i) idx[i] = (uint32_t)i;
        for (size_t i = 1; i < lam; ++i) { /* stable insertion sort by plain < (:258-261) */
            const uint32_t o = idx[i];
            size_t j = i;
            while (j > 0 && f[o] < f[idx[j - 1]]) { idx[j] = idx[j - 1]; --j; }
            idx[j] = o;
        }
        oracle_weighted_mean(z, idx, u, lam, D, dc);
        oracle_weighted_gram(z, idx, NULL, u, lam, D, 1.0, G);
        double usum = 0.;
        for (size_t i = 0; i < lam; ++i) usum += u[i];
        for (size_t a = 0; a < D; ++a) G[a * D + a] -= usum; /* sum u_i (z z^T - I) */
        double cov_trace = 0.;
        for (size_t a = 0; a < D; ++a) cov_trace += G[a * D + a];
        for (size_t a = 0; a < D; ++a) G[a * D + a] -= cov_trace / dim_d;
        for (size_t a = 0; a < D; ++a)
            for (size_t b = 0; b < D; ++b) dA[a * D + b] = 0.5 * ((a == b ? eta_sigma * cov_trace / dim_d : 0.) + eta_b * G[a * D + b]);
        for (size_t a = 0; a < D; ++a) {
            double y = 0.;
            for (size_t j = 0; j < D; ++j) y += A[a * D + j] * dc[j];
            tmp[a] = y;
        }
        for (size_t a = 0; a < D; ++a) mean[a] = mean[a] + eta_mu * tmp[a];
        /* A <- A exp(d_A): d_A is symmetric, exp(d_A) = V diag(exp(w)) V^T */
        for (size_t a = 0; a < D; ++a)
            for (size_t b = a + 1; b < D; ++b) dA[a * D + b] = dA[b * D + a] = (dA[a * D + b] + dA[b * D + a]) / 2.;
        memcpy(work, dA, D * D * sizeof(double));
        jacobi_eigen(work, D, wv, V);
        for (size_t a = 0; a < D; ++a)
            for (size_t b = 0; b < D; ++b) {
                double y = 0.;
                for (size_t k = 0; k < D; ++k) y += V[a * D + k] * exp(wv[k]) * V[b * D + k];
                E[a * D + b] = y;
            }
        for (size_t a = 0; a < D; ++a)
            for (size_t b = 0; b < D; ++b) {
                double y = 0.;
                for (size_t k = 0; k < D; ++k) y += A[a * D + k] * E[k * D + b];
                An[a * D + b] = y;
            }
        memcpy(A, An, D * D * sizeof(double));
        sigma = sigma * exp(eta_sigma / 2. * cov_trace / dim_d);
    }
    if (gens_done) *gens_done = done;
    if (sigma_out) *sigma_out = sigma;
    free(u); free(A); free(mean); free(z); free(xn); free(fn); free(dc); free(G); free(dA); free(work); free(wv); free(V); free(E); free(An);
    free(tmp); free(idx);
    return rc;
}
