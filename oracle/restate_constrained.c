/* oracle/restate_constrained.c - CPU restatement of the two constrained UDPs with a device evaluator and of the
 * `unconstrain` meta-problem.  TEST INFRASTRUCTURE ONLY (see oracle.h).  Pinned bit for bit against the
 * unmodified reference (pagmo::hock_schittkowski_71, pagmo::luksan_vlcek1, pagmo::unconstrain in oracle/_ref) by
 * tests/test_oracle.py. */
#include <float.h>
#include <math.h>

#include "oracle.h"

/* hock_schittkowski_71::fitness, reference src/problems/hock_schittkowski_71.cpp:48-55: [objective | 1 equality | 1 inequality] */
int oracle_hs71_batch(const double *xs, size_t n, double *fs)
{
    for (size_t i = 0; i < n; ++i) {
        const double *x = xs + 4 * i;
        double *f = fs + 3 * i;
        f[0] = x[0] * x[3] * (x[0] + x[1] + x[2]) + x[2];
        f[1] = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3] - 40.;
        f[2] = 25. - x[0] * x[1] * x[2] * x[3];
    }
    return 0;
}

/* luksan_vlcek1::fitness, reference src/problems/luksan_vlcek1.cpp:60-77: [objective | dim - 2 equalities] */
int oracle_luksan_vlcek1_batch(size_t dim, const double *xs, size_t n, double *fs)
{
    if (dim < 3) return -1; /* :46-49 */
    for (size_t r = 0; r < n; ++r) {
        const double *x = xs + dim * r;
        double *f = fs + (dim - 1) * r;
        f[0] = 0.;
        for (size_t i = 0; i < dim - 1; ++i) {
            const double a1 = x[i] * x[i] - x[i + 1];
            const double a2 = x[i] - 1.;
            f[0] += 100. * a1 * a1 + a2 * a2;
        }
        for (size_t i = 0; i < dim - 2; ++i)
            f[i + 1] = (3. * pow(x[i + 1], 3.) + 2. * x[i + 2] - 5. + sin(x[i + 1] - x[i + 2]) * sin(x[i + 1] + x[i + 2]) + 4. * x[i + 1]
                        - x[i] * exp(x[i] - x[i + 1]) - 3.);
    }
    return 0;
}

/* detail::test_eq_constraints / test_ineq_constraints, reference include/pagmo/utils/constrained.hpp:49-80:
 * number of satisfied constraints and the L2 norm of the violation */
static inline double max0(double a) { return a < 0. ? 0. : a; } /* std::max(a, 0.): a NaN stays a NaN (never satisfied) */

static void test_eq(const double *c, size_t k, const double *tol, size_t *sat, double *l2norm)
{
    double l2 = 0.;
    size_t n = 0;
    for (size_t i = 0; i < k; ++i) {
        const double err = max0(fabs(c[i]) - tol[i]);
        l2 += err * err;
        if (err <= 0.) ++n;
    }
    *sat = n;
    *l2norm = sqrt(l2);
}

static void test_ineq(const double *c, size_t k, const double *tol, size_t *sat, double *l2norm)
{
    double l2 = 0.;
    size_t n = 0;
    for (size_t i = 0; i < k; ++i) {
        const double err = max0(c[i] - tol[i]);
        l2 += err * err;
        if (err <= 0.) ++n;
    }
    *sat = n;
    *l2norm = sqrt(l2);
}

/* unconstrain::penalize per row, reference src/problems/unconstrain.cpp:136-223.  method: 0 death penalty, 1 kuri, 2 weighted,
 * 3 ignore_c, 4 ignore_o.  fs rows are [nobj | nec | nic]; out rows are nobj wide (1 wide for ignore_o, :220). */
int oracle_unconstrain_rows(const double *fs, size_t n, size_t nobj, size_t nec, size_t nic, const double *c_tol, int method,
                            const double *weights, double *out)
{
    const size_t nc = nec + nic, nf = nobj + nc;
    if (nc == 0 || method < 0 || method > 4) return -1;
    for (size_t r = 0; r < n; ++r) {
        const double *f = fs + r * nf;
        size_t sat_ec, sat_ic;
        double norm_ec, norm_ic;
        test_eq(f + nobj, nec, c_tol, &sat_ec, &norm_ec);
        test_ineq(f + nobj + nec, nic, c_tol + nec, &sat_ic, &norm_ic);
        const int feasible = (sat_ec + sat_ic == nc); /* problem::feasibility_f, src/problem.cpp:709-721 */
        if (method == 4) {                             /* :206-221 */
            out[r] = norm_ec + norm_ic;
            continue;
        }
        double *o = out + r * nobj;
        for (size_t k = 0; k < nobj; ++k) o[k] = f[k];
        if (method == 0) { /* :150-158 */
            if (!feasible)
                for (size_t k = 0; k < nobj; ++k) o[k] = DBL_MAX;
        } else if (method == 1) { /* :159-179 */
            if (!feasible) {
                const double penalty = DBL_MAX * (1. - (double)(sat_ec + sat_ic) / (double)nc);
                for (size_t k = 0; k < nobj; ++k) o[k] = penalty;
            }
        } else if (method == 2) { /* :180-204 */
            double penalty = 0.;
            for (size_t i = 0; i < nc; ++i) {
                const double c = (i < nec) ? fabs(f[nobj + i]) - c_tol[i] : f[nobj + i] - c_tol[i];
                if (!(c <= 0.)) penalty += weights[i] * c;
            }
            for (size_t k = 0; k < nobj; ++k) o[k] += penalty;
        } /* method 3: the objectives alone, :203-205 */
    }
    return 0;
}
