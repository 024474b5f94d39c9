/* oracle/restate_wfg.c - plain-C restatement of pagmo::wfg::fitness (WFG1..WFG9).  TEST INFRASTRUCTURE ONLY.
 * Follows reference src/problems/wfg.cpp: shape functions :161-222, transformation functions :225-302, problems :304-1066.
 * Same operation order and libm calls as the reference, so the result is BIT-IDENTICAL to the reference compiled from source
 * (oracle/_ref); asserted by tests/test_oracle.py together with the reference's own 45 known answers (tests/wfg.cpp:75-181).
 * Structure is ours: every problem is  normalise -> element-wise transformations -> one reduction to M values -> shape functions,
 * described by a small table instead of nine near-identical functions.  Kept quirk: r_nonsep's denominator takes
 * ceil(A / 2) with INTEGER division (:298), i.e. floor(A/2), next to a true ceil(A / 2.0).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define PI 3.141592653589793238462643383279502884

static double s_linear(double y, double a) { return fabs(y - a) / (fabs(floor(a - y) + a)); }
static double b_flat(double y, double a, double b, double c)
{
    return a + fmin(0.0, floor(y - b)) * a * (b - y) / (b)-fmin(0.0, floor(c - y)) * (1.0 - a) * (y - c) / (1 - c);
}
static double b_param(double y, double u, double a, double b, double c)
{
    double v = a - (1.0 - 2 * u) * fabs(floor(0.5 - u) + a);
    return pow(y, b + (c - b) * v);
}
static double s_decept(double y, double a, double b, double c)
{
    return 1.0
           + (fabs(y - a) - b)
                 * ((floor(y - a + b) * (1.0 - c + (a - b) / b)) / (a - b) + (floor(a + b - y) * (1.0 - c + (1.0 - a - b) / b)) / (1.0 - a - b)
                    + 1.0 / b);
}
static double s_multi(double y, double a, double b, double c)
{
    return (1 + cos((4.0 * a + 2.0) * PI * (0.5 - (fabs(y - c)) / (2.0 * (floor(c - y) + c))))
            + 4.0 * b * pow(fabs(y - c) / (2 * (floor(c - y) + c)), 2))
           / (b + 2.0);
}
/* r_sum with weights w0 + dw*index (dw = 0: constant weights) over y[lo, hi) */
static double r_sum(const double *y, size_t lo, size_t hi, int weighted)
{
    double g1 = 0., g2 = 0.;
    for (size_t j = lo; j < hi; ++j) {
        const double w = weighted ? 2. * ((double)j + 1) : 1.0;
        g1 += w * y[j];
        g2 += w;
    }
    return g1 / g2;
}
static double r_nonsep(const double *y, size_t lo, size_t hi, size_t A)
{
    if (A == 1) return r_sum(y, lo, hi, 0);
    const size_t len = hi - lo;
    double g = 0.;
    for (size_t j = 0; j < len; ++j) {
        g += y[lo + j];
        for (size_t i = 0; i <= A - 2; ++i) g += fabs(y[lo + j] - y[lo + (1 + j + i) % len]);
    }
    return g / ((double)len / (double)A * ceil((double)(A / 2)) * (1.0 + 2.0 * (double)A - 2.0 * ceil((double)A / 2.0)));
}

enum shape { CONVEX, LINEAR, CONCAVE };
static double shape_fn(enum shape s, const double *p, size_t m, size_t M)
{
    double g = 1.;
    const size_t lim = (m == 1) ? M - 1 : M - m;
    switch (s) {
        case LINEAR:
            if (m == M) return 1.0 - p[0];
            for (size_t i = 0; i < lim; ++i) g *= p[i];
            return m == 1 ? g : g * (1.0 - p[M - m]);
        case CONVEX:
            for (size_t i = 0; i < lim; ++i) g *= 1.0 - cos(p[i] * PI / 2.0);
            return m == 1 ? g : g * (1 - sin(p[M - m] * PI / 2.0));
        default:
            if (m == M) return cos(p[0] * PI / 2.0);
            for (size_t i = 0; i < lim; ++i) g *= sin(p[i] * PI / 2.0);
            return m == 1 ? g : g * cos(p[M - m] * PI / 2.0);
    }
}
static double mixed(double p0, double alpha, double A) { return pow((1.0 - p0 - cos(2 * A * PI * p0 + PI / 2.0) / (2.0 * A * PI)), alpha); }
static double disconnected(double p0, double alpha, double beta, double A)
{
    return 1.0 - pow(p0, alpha) * pow(cos(A * pow(p0, beta) * PI), 2);
}

int oracle_wfg_check(unsigned prob_id, size_t n, size_t M, size_t k) /* ctor checks :64-93 */
{
    if (prob_id == 0 || prob_id > 9 || n < 1 || M < 2) return -1;
    if (k >= n || k < 1 || k % (M - 1) != 0) return -1;
    if ((prob_id == 2 || prob_id == 3) && (n - k) % 2 != 0) return -1;
    return 0;
}

int oracle_wfg_fitness(unsigned prob_id, size_t n, size_t M, size_t k, const double *x, double *f)
{
    if (oracle_wfg_check(prob_id, n, M, k)) return -1;
    const size_t l = n - k;
    double *y = (double *)malloc(2 * n * sizeof(double)), *xn = y + n;
    double *t = (double *)malloc(2 * M * sizeof(double)), *par = t + M;
    for (size_t i = 0; i < n; ++i) xn[i] = x[i] / (2.0 * ((double)i + 1)); /* get_bounds().second[i], :138-146 */
    memcpy(y, xn, n * sizeof(double));
    size_t red_n = n; /* length of the vector the final reduction sees beyond k */
    switch (prob_id) {
        case 1: /* :326-352 */
            for (size_t i = k; i < n; ++i) y[i] = s_linear(y[i], 0.35);
            for (size_t i = k; i < n; ++i) y[i] = b_flat(y[i], 0.8, 0.75, 0.85);
            for (size_t i = 0; i < n; ++i) y[i] = pow(y[i], 0.02);
            break;
        case 2:
        case 3: /* :418-444, :515-541 */
            for (size_t i = k; i < n; ++i) y[i] = s_linear(y[i], 0.35);
            for (size_t i = k + 1; i <= k + l / 2; ++i) {
                const size_t head = k + 2 * (i - k) - 2;
                y[i - 1] = r_nonsep(y, head, head + 2, 2); /* in place: reads only indices >= i - 1 */
            }
            red_n = k + l / 2;
            break;
        case 4: for (size_t i = 0; i < n; ++i) y[i] = s_multi(y[i], 30.0, 10.0, 0.35); break;   /* :611-614 */
        case 5: for (size_t i = 0; i < n; ++i) y[i] = s_decept(y[i], 0.35, 0.001, 0.05); break; /* :681-684 */
        case 6: for (size_t i = k; i < n; ++i) y[i] = s_linear(y[i], 0.35); break;              /* :754-761 */
        case 7: /* :819-843 */
            for (size_t i = 1; i <= k; ++i) y[i - 1] = b_param(xn[i - 1], r_sum(xn, i, n, 0), 0.98 / 49.98, 0.02, 50);
            for (size_t i = k; i < n; ++i) y[i] = s_linear(y[i], 0.35);
            break;
        case 8: /* :905-928: position i depends on the already transformed prefix */
            for (size_t i = k; i < n; ++i) y[i] = b_param(xn[i], r_sum(y, 0, i, 0), 0.98 / 49.98, 0.02, 50);
            for (size_t i = k; i < n; ++i) y[i] = s_linear(y[i], 0.35);
            break;
        case 9: /* :992-1016 */
            for (size_t i = 0; i + 1 < n; ++i) y[i] = b_param(xn[i], r_sum(xn, i + 1, n, 0), 0.98 / 49.98, 0.02, 50);
            for (size_t i = 0; i < n; ++i) y[i] = i < k ? s_decept(y[i], 0.35, 0.001, 0.05) : s_multi(y[i], 30.0, 95.0, 0.35);
            break;
    }
    /* reduction to M values: M-1 groups of the first k, one group of the rest */
    const int nonsep = prob_id == 6 || prob_id == 9, weighted = prob_id == 1;
    for (size_t i = 1; i <= M - 1; ++i) {
        const size_t head = (i - 1) * k / (M - 1), tail = i * k / (M - 1);
        t[i - 1] = nonsep ? r_nonsep(y, head, tail, k / (M - 1)) : r_sum(y, head, tail, weighted);
    }
    t[M - 1] = nonsep ? r_nonsep(y, k, n, l) : r_sum(y, k, red_n, weighted);
    for (size_t i = 0; i < M; ++i) {
        const double floor_ = (prob_id == 3 && i > 0) ? 0.0 : 1.0; /* WFG3 degenerate front, :568-575 */
        par[i] = fmax(t[M - 1], floor_) * (t[i] - 0.5) + 0.5;
    }
    par[M - 1] = t[M - 1];
    for (size_t i = 0; i < M; ++i) {
        double sh;
        if (prob_id == 1) sh = i + 1 < M ? shape_fn(CONVEX, par, i + 1, M) : mixed(par[0], 1.0, 5.0);
        else if (prob_id == 2) sh = i + 1 < M ? shape_fn(CONVEX, par, i + 1, M) : disconnected(par[0], 1.0, 1.0, 5.0);
        else sh = shape_fn(prob_id == 3 ? LINEAR : CONCAVE, par, i + 1, M);
        f[i] = par[M - 1] + 2.0 * ((double)i + 1) * sh;
    }
    free(y);
    free(t);
    return 0;
}

int oracle_wfg_batch(unsigned prob_id, size_t n, size_t M, size_t k, const double *xs, size_t count, double *fs)
{
    for (size_t q = 0; q < count; ++q)
        if (oracle_wfg_fitness(prob_id, n, M, k, xs + q * n, fs + q * M)) return -1;
    return 0;
}
