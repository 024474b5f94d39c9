/* oracle/restate_mo.c - plain-C restatement of the multi-objective UDPs ZDT1-6 and DTLZ1-7.
 * TEST INFRASTRUCTURE ONLY: the checker the CUDA path is compared against; never linked into the product.
 * Follows reference src/problems/zdt.cpp:233-356 and src/problems/dtlz.cpp:182-409 operation by operation.
 * Pinned bit-exactly against oracle/_ref (tests/test_oracle.py) and against the reference's own known answers
 * (tests/zdt.cpp:72-176, tests/dtlz.cpp:70-...) through tests/golden/mo_ref.npz.
 */
#include <math.h>
#include <stddef.h>

#include "oracle.h"

static const double kPi = 3.141592653589793238462643383279502884;
static const double kPiHalf = 1.570796326794896619231321691639751442;

/* zdt.cpp:233-360; N = length of x */
int oracle_zdt_fitness(unsigned id, const double *x, size_t N, double *f)
{
    double g = 0.;
    size_t i;
    if (id < 1 || id > 6 || N < 2) return -1;
    switch (id) {
        case 1:
            for (i = 1; i < N; ++i) g += x[i];
            g = 1. + (9. * g) / (double)(N - 1u);
            f[0] = x[0];
            f[1] = g * (1. - sqrt(x[0] / g));
            return 0;
        case 2:
            for (i = 1; i < N; ++i) g += x[i];
            g = 1. + (9. * g) / (double)(N - 1u);
            f[0] = x[0];
            f[1] = g * (1. - (x[0] / g) * (x[0] / g));
            return 0;
        case 3:
            for (i = 1; i < N; ++i) g += x[i];
            g = 1. + (9. * g) / (double)(N - 1u);
            f[0] = x[0];
            f[1] = g * (1. - sqrt(x[0] / g) - x[0] / g * sin(10. * kPi * x[0]));
            return 0;
        case 4:
            g = 1 + 10 * (double)(N - 1u);
            for (i = 1; i < N; ++i) g += x[i] * x[i] - 10. * cos(4. * kPi * x[i]);
            f[0] = x[0];
            f[1] = g * (1. - sqrt(x[0] / g));
            return 0;
        case 5: {
            const size_t n_vectors = ((N - 30u) / 5u) + 1u;
            size_t k = 30, u0 = 0;
            for (i = 0; i < 30; ++i) u0 += (round(x[i]) == 1.);
            for (i = 1; i < n_vectors; ++i) {
                size_t u = 0;
                for (int j = 0; j < 5; ++j) u += (round(x[k++]) == 1.);
                g += (double)(u < 5u ? 2u + u : 1u);
            }
            f[0] = 1.0 + (double)u0;
            f[1] = g * (1. / f[0]);
            return 0;
        }
        default:
            f[0] = 1 - exp(-4 * x[0]) * pow(sin(6 * kPi * x[0]), 6);
            for (i = 1; i < N; ++i) g += x[i];
            g = 1 + 9 * pow((g / (double)(N - 1u)), 0.25);
            f[1] = g * (1 - (f[0] / g) * (f[0] / g));
            return 0;
    }
}

/* dtlz.cpp:182-245 */
static double dtlz_g(unsigned id, const double *xm, size_t len)
{
    double y = 0.;
    size_t i;
    switch (id) {
        case 1: case 3:
            for (i = 0; i < len; ++i) y += pow(xm[i] - 0.5, 2) - cos(20. * kPi * (xm[i] - 0.5));
            return 100. * (y + (double)len);
        case 2: case 4: case 5:
            for (i = 0; i < len; ++i) y += pow(xm[i] - 0.5, 2);
            return y;
        case 6:
            for (i = 0; i < len; ++i) y += pow(xm[i], 0.1);
            return y;
        default:
            for (i = 0; i < len; ++i) y += xm[i];
            return (9. / (double)len) * y;
    }
}

/* dtlz.cpp:258-403; N = dim, M = fdim */
int oracle_dtlz_fitness(unsigned id, const double *x, size_t N, size_t M, unsigned alpha, double *f)
{
    size_t i, j;
    if (id < 1 || id > 7 || M < 2 || N <= M) return -1;
    double g = dtlz_g(id, x + (M - 1), N - (M - 1));
    if (id == 1) {
        f[0] = 0.5 * (1. + g);
        for (i = 0; i < M - 1u; ++i) f[0] *= x[i];
        for (i = 1; i < M - 1u; ++i) {
            f[i] = 0.5 * (1.0 + g);
            for (j = 0; j < M - (i + 1); ++j) f[i] *= x[j];
            f[i] *= 1. - x[M - (i + 1u)];
        }
        f[M - 1u] = 0.5 * (1. - x[0]) * (1. + g);
        return 0;
    }
    if (id == 7) {
        double y = 0.;
        g = 1. + g;
        for (i = 0; i < M - 1u; ++i) f[i] = x[i];
        for (i = 0; i < M - 1u; ++i) y += (f[i] / (1.0 + g)) * (1.0 + sin(3 * kPi * f[i]));
        f[M - 1u] = (1. + g) * ((double)M - y);
        return 0;
    }
    {
        double ang[64]; /* angle of variable i, i < M */
        if (M > 64) return -1;
        for (i = 0; i < M; ++i) {
            if (id == 4) ang[i] = pow(x[i], alpha) * kPiHalf;
            else if (id == 5 || id == 6) {
                const double t = 1. / (2. * (1. + g));
                const double theta = (i == 0) ? x[0] : t + ((g * x[i]) / (1.0 + g));
                ang[i] = theta * kPiHalf;
            } else ang[i] = x[i] * kPiHalf;
        }
        f[0] = (1. + g);
        for (i = 0; i < M - 1u; ++i) f[0] *= cos(ang[i]);
        for (i = 1; i < M - 1u; ++i) {
            f[i] = (1. + g);
            for (j = 0; j < M - (i + 1u); ++j) f[i] *= cos(ang[j]);
            f[i] *= sin(ang[M - (i + 1u)]);
        }
        f[M - 1u] = (1. + g) * sin(ang[0]);
    }
    return 0;
}

int oracle_zdt_batch(unsigned id, const double *xs, size_t n, size_t N, double *fs)
{
    for (size_t i = 0; i < n; ++i)
        if (oracle_zdt_fitness(id, xs + i * N, N, fs + i * 2)) return -1;
    return 0;
}

int oracle_dtlz_batch(unsigned id, const double *xs, size_t n, size_t N, size_t M, unsigned alpha, double *fs)
{
    for (size_t i = 0; i < n; ++i)
        if (oracle_dtlz_fitness(id, xs + i * N, N, M, alpha, fs + i * M)) return -1;
    return 0;
}
