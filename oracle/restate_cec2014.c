/* oracle/restate_cec2014.c - plain-C restatement of pagmo::cec2014::fitness (CEC2014 f1..f30).
 * TEST INFRASTRUCTURE ONLY: the checker the CUDA path is compared against; never linked into the product.
 *
 * Follows reference src/problems/cec2014.cpp operation by operation (same evaluation order, same libm calls),
 * so that it is BIT-IDENTICAL to the reference compiled from source (oracle/_ref) - asserted by
 * tests/test_oracle_vs_reference.py for every function and dimension.  Structure is ours: primitives are pure
 * functions of an already shifted/scaled/rotated vector; hybrids and compositions are table driven.
 * The reference's own value tests (tests/cec2014.cpp:76-123) need the data tables that are missing from the
 * checkout, so parity is pinned on the compiled reference + synthetic tables (oracle/cec_synth.c) instead.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define PI 3.141592653589793238462643383279502884 /* boost::math::constants::pi<double>() */
#define E_ 2.718281828459045235360287471352662498  /* boost::math::constants::e<double>() */
#define MAXD 100

enum prim { ELLIPS, BENT_CIGAR, DISCUS, ROSENBROCK, ACKLEY, WEIERSTRASS, GRIEWANK, RASTRIGIN, SCHWEFEL, KATSUURA,
            HAPPYCAT, HGBAT, GRIE_ROSEN, ESCAFFER6 };

/* sh_rate handed to sr_func by each primitive (cec2014.cpp:381,393,407,437,474,498,522,539,573,602,700,724,749,772) */
static double rate_of(enum prim p)
{
    switch (p) {
        case ROSENBROCK: return 2.048 / 100.0;
        case WEIERSTRASS: return 0.5 / 100.0;
        case GRIEWANK: return 600.0 / 100.0;
        case RASTRIGIN: return 5.12 / 100.0;
        case SCHWEFEL: return 1000.0 / 100.0;
        case KATSUURA: case HAPPYCAT: case HGBAT: case GRIE_ROSEN: return 5.0 / 100.0;
        default: return 1.0;
    }
}

/* sr_func, cec2014.cpp:1238-1274 (shiftfunc :1215-1222, rotatefunc :1224-1235).  y is scratch (the UDP's m_y). */
static void shift_scale_rotate(const double *x, double *out, unsigned nx, const double *Os, const double *Mr,
                               double rate, int s_flag, int r_flag, double *y)
{
    double *dst = r_flag ? y : out;
    for (unsigned i = 0; i < nx; ++i) {
        double v = s_flag ? x[i] - Os[i] : x[i];
        dst[i] = v * rate;
    }
    if (r_flag) {
        for (unsigned i = 0; i < nx; ++i) {
            out[i] = 0;
            for (unsigned j = 0; j < nx; ++j) out[i] = out[i] + y[j] * Mr[i * nx + j];
        }
    }
}

/* The primitive bodies AFTER sr_func; z is modified in place exactly where the reference does. */
static double apply_prim(enum prim p, double *z, unsigned nx)
{
    double f = 0.0;
    unsigned i, j;
    switch (p) {
        case ELLIPS: /* :382-384 */
            for (i = 0; i < nx; i++) f += pow(10.0, 6.0 * i / (nx - 1)) * z[i] * z[i];
            return f;
        case BENT_CIGAR: /* :395-398 */
            f = z[0] * z[0];
            for (i = 1; i < nx; i++) f += pow(10.0, 6.0) * z[i] * z[i];
            return f;
        case DISCUS: /* :408-411 */
            f = pow(10.0, 6.0) * z[0] * z[0];
            for (i = 1; i < nx; i++) f += z[i] * z[i];
            return f;
        case ROSENBROCK: { /* :438-444 */
            z[0] += 1.0;
            for (i = 0; i + 1 < nx; i++) {
                z[i + 1] += 1.0;
                const double tmp1 = z[i] * z[i] - z[i + 1], tmp2 = z[i] - 1.0;
                f += 100.0 * tmp1 * tmp1 + tmp2 * tmp2;
            }
            return f;
        }
        case ACKLEY: { /* :476-482 */
            double sum1 = 0.0, sum2 = 0.0;
            for (i = 0; i < nx; i++) {
                sum1 += z[i] * z[i];
                sum2 += cos(2.0 * PI * z[i]);
            }
            sum1 = -0.2 * sqrt(sum1 / nx);
            sum2 /= nx;
            return E_ - 20.0 * exp(sum1) - exp(sum2) + 20.0;
        }
        case WEIERSTRASS: { /* :491-509 */
            const double a = 0.5, b = 3.0;
            double sum2 = 0.0;
            for (i = 0; i < nx; i++) {
                double sum = 0.0;
                sum2 = 0.0;
                for (j = 0; j <= 20; j++) {
                    sum += pow(a, j) * cos(2.0 * PI * pow(b, j) * (z[i] + 0.5));
                    sum2 += pow(a, j) * cos(2.0 * PI * pow(b, j) * 0.5);
                }
                f += sum;
            }
            f -= nx * sum2;
            return f;
        }
        case GRIEWANK: { /* :524-528 */
            double s = 0.0, pr = 1.0;
            for (i = 0; i < nx; i++) {
                s += z[i] * z[i];
                pr *= cos(z[i] / sqrt(1.0 + i));
            }
            return 1.0 + s / 4000.0 - pr;
        }
        case RASTRIGIN: /* :541-543 */
            for (i = 0; i < nx; i++) f += (z[i] * z[i] - 10.0 * cos(2.0 * PI * z[i]) + 10.0);
            return f;
        case SCHWEFEL: { /* :575-589 */
            for (i = 0; i < nx; i++) {
                double tmp;
                z[i] += 4.209687462275036e+002;
                if (z[i] > 500) {
                    f -= (500.0 - fmod(z[i], 500)) * sin(pow(500.0 - fmod(z[i], 500), 0.5));
                    tmp = (z[i] - 500.0) / 100;
                    f += tmp * tmp / nx;
                } else if (z[i] < -500) {
                    f -= (-500.0 + fmod(fabs(z[i]), 500)) * sin(pow(500.0 - fmod(fabs(z[i]), 500), 0.5));
                    tmp = (z[i] + 500.0) / 100;
                    f += tmp * tmp / nx;
                } else {
                    f -= z[i] * sin(pow(fabs(z[i]), 0.5));
                }
            }
            f += 4.189828872724338e+002 * nx;
            return f;
        }
        case KATSUURA: { /* :599-614 */
            const double tmp3 = pow(1.0 * nx, 1.2);
            f = 1.0;
            for (i = 0; i < nx; i++) {
                double temp = 0.0;
                for (j = 1; j <= 32; j++) {
                    const double tmp1 = pow(2.0, j), tmp2 = tmp1 * z[i];
                    temp += fabs(tmp2 - floor(tmp2 + 0.5)) / tmp1;
                }
                f *= pow(1.0 + (i + 1) * temp, 10.0 / tmp3);
            }
            {
                const double tmp1 = 10.0 / nx / nx;
                f = f * tmp1 - tmp1;
            }
            return f;
        }
        case HAPPYCAT: case HGBAT: { /* :747-759, :770-782 */
            double r2 = 0.0, sum_z = 0.0;
            for (i = 0; i < nx; i++) {
                z[i] = z[i] - 1.0;
                r2 += z[i] * z[i];
                sum_z += z[i];
            }
            if (p == HAPPYCAT) return pow(fabs(r2 - nx), 2 * (1.0 / 8.0)) + (0.5 * r2 + sum_z) / nx + 0.5;
            return pow(fabs(pow(r2, 2.0) - pow(sum_z, 2.0)), 2 * (1.0 / 4.0)) + (0.5 * r2 + sum_z) / nx + 0.5;
        }
        case GRIE_ROSEN: { /* :702-713 */
            double temp, tmp1, tmp2;
            z[0] += 1.0;
            for (i = 0; i + 1 < nx; i++) {
                z[i + 1] += 1.0;
                tmp1 = z[i] * z[i] - z[i + 1];
                tmp2 = z[i] - 1.0;
                temp = 100.0 * tmp1 * tmp1 + tmp2 * tmp2;
                f += (temp * temp) / 4000.0 - cos(temp) + 1.0;
            }
            tmp1 = z[nx - 1] * z[nx - 1] - z[0];
            tmp2 = z[nx - 1] - 1.0;
            temp = 100.0 * tmp1 * tmp1 + tmp2 * tmp2;
            f += (temp * temp) / 4000.0 - cos(temp) + 1.0;
            return f;
        }
        case ESCAFFER6: { /* :726-736 */
            double temp1, temp2;
            for (i = 0; i + 1 < nx; i++) {
                temp1 = sin(sqrt(z[i] * z[i] + z[i + 1] * z[i + 1]));
                temp1 = temp1 * temp1;
                temp2 = 1.0 + 0.001 * (z[i] * z[i] + z[i + 1] * z[i + 1]);
                f += 0.5 + (temp1 - 0.5) / (temp2 * temp2);
            }
            temp1 = sin(sqrt(z[nx - 1] * z[nx - 1] + z[0] * z[0]));
            temp1 = temp1 * temp1;
            temp2 = 1.0 + 0.001 * (z[nx - 1] * z[nx - 1] + z[0] * z[0]);
            f += 0.5 + (temp1 - 0.5) / (temp2 * temp2);
            return f;
        }
    }
    return f;
}

/* one basic function: sr_func with the primitive's rate, then the primitive (e.g. ellips_func :375-385) */
static double basic(enum prim p, const double *x, unsigned nx, const double *Os, const double *Mr, int s_flag, int r_flag)
{
    double y[MAXD], z[MAXD];
    shift_scale_rotate(x, z, nx, Os, Mr, rate_of(p), s_flag, r_flag, y);
    return apply_prim(p, z, nx);
}

/* hybrid functions hf01..hf06, :786-1034 */
static const struct { int n; double Gp[5]; enum prim pr[5]; } HF[6] = {
    {3, {0.3, 0.3, 0.4}, {SCHWEFEL, RASTRIGIN, ELLIPS}},
    {3, {0.3, 0.3, 0.4}, {BENT_CIGAR, HGBAT, RASTRIGIN}},
    {4, {0.2, 0.2, 0.3, 0.3}, {GRIEWANK, WEIERSTRASS, ROSENBROCK, ESCAFFER6}},
    {4, {0.2, 0.2, 0.3, 0.3}, {HGBAT, DISCUS, GRIE_ROSEN, RASTRIGIN}},
    {5, {0.1, 0.2, 0.2, 0.2, 0.3}, {ESCAFFER6, HGBAT, ROSENBROCK, SCHWEFEL, ELLIPS}},
    {5, {0.1, 0.2, 0.2, 0.2, 0.3}, {KATSUURA, HAPPYCAT, GRIE_ROSEN, SCHWEFEL, ACKLEY}},
};

static double hybrid(int hf, const double *x, unsigned nx, const double *Os, const double *Mr, const int *S, int s_flag, int r_flag)
{
    const int cf_num = HF[hf].n;
    unsigned G[5], G_nx[5], tmp = 0;
    double y[MAXD], z[MAXD], f = 0.0;
    int i;
    for (i = 0; i < cf_num - 1; i++) {
        G_nx[i] = (unsigned)ceil(HF[hf].Gp[i] * nx);
        tmp += G_nx[i];
    }
    G_nx[cf_num - 1] = nx - tmp;
    G[0] = 0;
    for (i = 1; i < cf_num; i++) G[i] = G[i - 1] + G_nx[i - 1];
    shift_scale_rotate(x, z, nx, Os, Mr, 1.0, s_flag, r_flag, y);
    for (unsigned j = 0; j < nx; j++) y[j] = z[(unsigned)(S[j] - 1)]; /* :807-809 */
    for (i = 0; i < cf_num; i++) f += basic(HF[hf].pr[i], &y[G[i]], G_nx[i], Os, Mr, 0, 0);
    return f;
}

/* cf_cal, :1319-1353 */
static double cf_cal(const double *x, unsigned nx, const double *Os, const double *delta, const double *bias, double *fit, int cf_num)
{
    double w[5], w_max = 0, w_sum = 0, f = 0.0;
    int i;
    for (i = 0; i < cf_num; i++) {
        fit[i] += bias[i];
        w[i] = 0;
        for (unsigned j = 0; j < nx; j++) w[i] += pow(x[j] - Os[i * nx + j], 2.0);
        if (w[i] != 0)
            w[i] = pow(1.0 / w[i], 0.5) * exp(-w[i] / 2.0 / nx / pow(delta[i], 2.0));
        else
            w[i] = DBL_MAX;
        if (w[i] > w_max) w_max = w[i];
    }
    for (i = 0; i < cf_num; i++) w_sum = w_sum + w[i];
    if (w_max == 0) {
        for (i = 0; i < cf_num; i++) w[i] = 1;
        w_sum = cf_num;
    }
    for (i = 0; i < cf_num; i++) f = f + w[i] / w_sum * fit[i];
    return f;
}

/* compositions cf01..cf08, :1037-1213: per component (primitive or -hybrid-1, r_flag, mul, div) */
struct comp { int prim; int r; double mul, div; };
static const struct { int n; double delta[5]; struct comp c[5]; } CF[8] = {
    {5, {10, 20, 30, 40, 50}, {{ROSENBROCK, 1, 10000, 1e+4}, {ELLIPS, 1, 10000, 1e+10}, {BENT_CIGAR, 1, 10000, 1e+30}, {DISCUS, 1, 10000, 1e+10}, {ELLIPS, 0, 10000, 1e+10}}},
    {3, {20, 20, 20}, {{SCHWEFEL, 0, 0, 0}, {RASTRIGIN, 1, 0, 0}, {HGBAT, 1, 0, 0}}},
    {3, {10, 30, 50}, {{SCHWEFEL, 1, 1000, 4e+3}, {RASTRIGIN, 1, 1000, 1e+3}, {ELLIPS, 1, 1000, 1e+10}}},
    {5, {10, 10, 10, 10, 10}, {{SCHWEFEL, 1, 1000, 4e+3}, {HAPPYCAT, 1, 1000, 1e+3}, {ELLIPS, 1, 1000, 1e+10}, {WEIERSTRASS, 1, 1000, 400}, {GRIEWANK, 1, 1000, 100}}},
    {5, {10, 10, 10, 20, 20}, {{HGBAT, 1, 10000, 1000}, {RASTRIGIN, 1, 10000, 1e+3}, {SCHWEFEL, 1, 10000, 4e+3}, {WEIERSTRASS, 1, 10000, 400}, {ELLIPS, 1, 10000, 1e+10}}},
    {5, {10, 20, 30, 40, 50}, {{GRIE_ROSEN, 1, 10000, 4e+3}, {HAPPYCAT, 1, 10000, 1e+3}, {SCHWEFEL, 1, 10000, 4e+3}, {ESCAFFER6, 1, 10000, 2e+7}, {ELLIPS, 1, 10000, 1e+10}}},
    {3, {10, 30, 50}, {{-1, 1, 0, 0}, {-2, 1, 0, 0}, {-3, 1, 0, 0}}},
    {3, {10, 30, 50}, {{-4, 1, 0, 0}, {-5, 1, 0, 0}, {-6, 1, 0, 0}}},
};

static double composition(int cf, const double *x, unsigned nx, const double *Os, const double *Mr, const int *S)
{
    static const double bias[5] = {0, 100, 200, 300, 400};
    double fit[5];
    for (int i = 0; i < CF[cf].n; i++) {
        const struct comp *c = &CF[cf].c[i];
        if (c->prim < 0)
            fit[i] = hybrid(-c->prim - 1, x, nx, &Os[i * nx], &Mr[i * nx * nx], &S[i * nx], 1, c->r);
        else
            fit[i] = basic((enum prim)c->prim, x, nx, &Os[i * nx], &Mr[i * nx * nx], 1, c->r);
        if (c->mul != 0) fit[i] = c->mul * fit[i] / c->div;
    }
    return cf_cal(x, nx, Os, CF[cf].delta, bias, fit, CF[cf].n);
}

/* cec2014::fitness, :119-247 */
int oracle_cec2014_fitness(unsigned func, unsigned dim, const double *Mr, const double *Os, const int *S, const double *x, double *f)
{
    static const struct { enum prim p; int r; } B[16] = {
        {ELLIPS, 1}, {BENT_CIGAR, 1}, {DISCUS, 1}, {ROSENBROCK, 1}, {ACKLEY, 1}, {WEIERSTRASS, 1}, {GRIEWANK, 1}, {RASTRIGIN, 0},
        {RASTRIGIN, 1}, {SCHWEFEL, 0}, {SCHWEFEL, 1}, {KATSUURA, 1}, {HAPPYCAT, 1}, {HGBAT, 1}, {GRIE_ROSEN, 1}, {ESCAFFER6, 1}};
    if (!(dim == 2u || dim == 10u || dim == 20u || dim == 30u || dim == 50u || dim == 100u)) return -1; /* :51-55 */
    if (func < 1u || func > 30u) return -1;                                                              /* :56-60 */
    if (dim == 2 && ((func >= 17u && func <= 22u) || (func >= 29u && func <= 30u))) return -1;           /* :62-64 */
    double v;
    if (func <= 16) v = basic(B[func - 1].p, x, dim, Os, Mr, 1, B[func - 1].r);
    else if (func <= 22) v = hybrid((int)func - 17, x, dim, Os, Mr, S, 1, 1);
    else v = composition((int)func - 23, x, dim, Os, Mr, S);
    *f = v + 100.0 * func; /* f[0] += 100.0 ... 3000.0 */
    return 0;
}

size_t oracle_cec2014_compact_shift(const double *lines, size_t nlines, unsigned dim, double *out)
{
    size_t k = 0;
    for (size_t i = 0; i < nlines * 100; ++i)
        if ((i % 100) < dim) out[k++] = lines[i];
    return k;
}

struct job { unsigned func, dim; const double *Mr, *Os; const int *S; const double *xs; double *fs; size_t lo, hi; int rc; };

static void *worker(void *arg)
{
    struct job *jb = (struct job *)arg;
    for (size_t i = jb->lo; i < jb->hi; ++i) {
        const int rc = oracle_cec2014_fitness(jb->func, jb->dim, jb->Mr, jb->Os, jb->S, jb->xs + i * jb->dim, jb->fs + i);
        if (rc) jb->rc = rc;
    }
    return NULL;
}

int oracle_cec2014_batch(unsigned func, unsigned dim, const double *Mr, const double *Os, const int *S, const double *xs,
                         size_t n, double *fs, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    struct job *jobs = (struct job *)calloc((size_t)nthreads, sizeof(struct job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    const size_t chunk = n / (size_t)nthreads, rem = n % (size_t)nthreads;
    size_t lo = 0;
    int rc = 0;
    for (int t = 0; t < nthreads; ++t) {
        const size_t hi = lo + chunk + ((size_t)t < rem ? 1 : 0);
        jobs[t] = (struct job){func, dim, Mr, Os, S, xs, fs, lo, hi, 0};
        lo = hi;
        if (nthreads == 1) worker(&jobs[t]);
        else pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    free(jobs);
    free(th);
    return rc;
}
