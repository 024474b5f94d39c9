/* oracle/restate_meta.c - CPU restatement of the translate and decompose meta-problems.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h).  Pinned against the unmodified reference (pagmo::translate, pagmo::decompose, decompose_objectives in
 * oracle/_ref) by tests/test_oracle.py. */
#include <math.h>

#include "oracle.h"

/* translate::batch_fitness de-shifting, reference src/problems/translate.cpp:137-150: out[i][j] = xs[i][j] - t[j] */
int oracle_translate_rows(const double *xs, size_t n, size_t nx, const double *t, double *out)
{
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < nx; ++j) out[i * nx + j] = xs[i * nx + j] - t[j];
    return 0;
}

/* decompose_objectives, reference src/utils/multi_objective.cpp:582-638.  method: 0 weighted, 1 tchebycheff, 2 bi */
int oracle_decompose_objectives(const double *f, size_t m, const double *weight, const double *ref_point, int method, double *out)
{
    double fd = 0.;
    if (m == 0) return -1; /* :598-601 */
    if (method == 0) {     /* :603-606 */
        for (size_t i = 0; i < m; ++i) fd += weight[i] * f[i];
    } else if (method == 1) { /* :607-616 */
        for (size_t i = 0; i < m; ++i) {
            const double fixed_weight = (weight[i] == 0.) ? 1e-4 : weight[i];
            const double tmp = fixed_weight * fabs(f[i] - ref_point[i]);
            if (tmp > fd) fd = tmp;
        }
    } else if (method == 2) { /* :617-632 */
        const double THETA = 5.;
        double d1 = 0., weight_norm = 0., d2 = 0.;
        for (size_t i = 0; i < m; ++i) {
            d1 += (f[i] - ref_point[i]) * weight[i];
            weight_norm += pow(weight[i], 2);
        }
        weight_norm = sqrt(weight_norm);
        d1 = d1 / weight_norm;
        for (size_t i = 0; i < m; ++i) d2 += pow(f[i] - (ref_point[i] + d1 * weight[i] / weight_norm), 2);
        d2 = sqrt(d2);
        fd = d1 + THETA * d2;
    } else {
        return -1; /* :633-636 */
    }
    *out = fd;
    return 0;
}

int oracle_decompose_rows(const double *fs, size_t n, size_t m, const double *weight, const double *ref_point, int method, double *out)
{
    for (size_t i = 0; i < n; ++i) {
        const int rc = oracle_decompose_objectives(fs + i * m, m, weight, ref_point, method, out + i);
        if (rc) return rc;
    }
    return 0;
}
