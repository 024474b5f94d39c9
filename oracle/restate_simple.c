/* oracle/restate_simple.c - plain-C restatement of pagmo's simple single-objective UDPs.
 * TEST INFRASTRUCTURE ONLY: the checker the CUDA path is compared against; never linked into the product.
 * Pinned by tests/test_oracle_vs_reference.py against oracle/_ref (the unmodified reference, bit-exact) and
 * against the known answers in the reference's own tests (tests/rastrigin.cpp:56-57, ackley.cpp:56-57,
 * griewank.cpp:55-56, schwefel.cpp:55-56, rosenbrock.cpp:60-61) via tests/golden/.
 */
#include <math.h>
#include <stddef.h>

#include "oracle.h"

static const double kPi = 3.141592653589793238462643383279502884; /* pagmo::detail::pi(), detail/constants.hpp */

/* rastrigin.cpp:62-72 */
static double f_rastrigin(const double *x, size_t n)
{
    const double omega = 2. * kPi;
    double f = 0.;
    for (size_t i = 0; i < n; ++i) f += x[i] * x[i] - 10. * cos(omega * x[i]);
    f += 10. * (double)n;
    return f;
}

/* ackley.cpp:61-76 */
static double f_ackley(const double *x, size_t n)
{
    const double omega = 2. * kPi, nepero = exp(1.0);
    double s1 = 0., s2 = 0.;
    for (size_t i = 0; i < n; ++i) {
        s1 += x[i] * x[i];
        s2 += cos(omega * x[i]);
    }
    return -20 * exp(-0.2 * sqrt(1.0 / (double)n * s1)) - exp(1.0 / (double)n * s2) + 20 + nepero;
}

/* griewank.cpp:60-75 */
static double f_griewank(const double *x, size_t n)
{
    const double fr = 4000.;
    double retval = 0., p = 1.;
    for (size_t i = 0; i < n; ++i) retval += x[i] * x[i];
    for (size_t i = 0; i < n; ++i) p *= cos(x[i] / sqrt((double)i + 1.0));
    return (retval / fr - p + 1.);
}

/* schwefel.cpp:60-69 */
static double f_schwefel(const double *x, size_t n)
{
    double f = 0.;
    for (size_t i = 0; i < n; ++i) f += x[i] * sin(sqrt(fabs(x[i])));
    return 418.9828872724338 * (double)n - f;
}

/* rosenbrock.cpp:59-66 */
static double f_rosenbrock(const double *x, size_t n)
{
    double retval = 0.;
    for (size_t i = 0; i + 1 < n; ++i)
        retval += 100. * (x[i] * x[i] - x[i + 1]) * (x[i] * x[i] - x[i + 1]) + (x[i] - 1) * (x[i] - 1);
    return retval;
}

int oracle_simple_fitness(int family, size_t dim, const double *x, double *f)
{
    switch (family) {
        case ORACLE_RASTRIGIN: *f = f_rastrigin(x, dim); return 0;
        case ORACLE_ACKLEY: *f = f_ackley(x, dim); return 0;
        case ORACLE_GRIEWANK: *f = f_griewank(x, dim); return 0;
        case ORACLE_SCHWEFEL: *f = f_schwefel(x, dim); return 0;
        case ORACLE_ROSENBROCK: *f = f_rosenbrock(x, dim); return 0;
        default: return -1;
    }
}

int oracle_simple_batch(int family, size_t dim, const double *xs, size_t n, double *fs)
{
    for (size_t i = 0; i < n; ++i) {
        const int rc = oracle_simple_fitness(family, dim, xs + i * dim, fs + i);
        if (rc) return rc;
    }
    return 0;
}

/* box bounds: rastrigin.cpp:80-85, ackley.cpp:85-90, griewank.cpp:84-89, schwefel.cpp:77-82, rosenbrock.cpp:72-75 */
int oracle_simple_bounds(int family, double *lo, double *hi)
{
    switch (family) {
        case ORACLE_RASTRIGIN: *lo = -5.12; *hi = 5.12; return 0;
        case ORACLE_ACKLEY: *lo = -15; *hi = 30; return 0;
        case ORACLE_GRIEWANK: *lo = -600; *hi = 600; return 0;
        case ORACLE_SCHWEFEL: *lo = -500; *hi = 500; return 0;
        case ORACLE_ROSENBROCK: *lo = -5.; *hi = 10.; return 0;
        default: return -1;
    }
}
