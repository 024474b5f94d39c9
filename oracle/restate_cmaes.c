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
