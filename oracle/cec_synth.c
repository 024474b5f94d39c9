/* oracle/cec_synth.c - see cec_synth.h.  TEST INFRASTRUCTURE ONLY (never linked into the product). */
#include "cec_synth.h"

#include <math.h>
#include <stdlib.h>

static uint64_t sm64_next(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static uint64_t mix_seed(uint64_t a, uint64_t b, uint64_t c, uint64_t d)
{
    uint64_t s = a;
    s = sm64_next(&s) ^ (b * 0xD6E8FEB86659FD93ull);
    s = sm64_next(&s) ^ (c * 0xCA5A826395121157ull);
    s = sm64_next(&s) ^ (d * 0x9FB21C651E98DF25ull);
    return sm64_next(&s);
}

static double u01(uint64_t *s)
{
    return (double)(sm64_next(s) >> 11) * (1.0 / 9007199254740992.0);
}

void cec_synth_uniform(uint64_t seed, double lo, double hi, double *out, size_t n)
{
    uint64_t s = seed;
    for (size_t i = 0; i < n; ++i) out[i] = lo + (hi - lo) * u01(&s);
}

/* Rows = modified Gram-Schmidt (two passes) of a matrix with i.i.d. U(-1,1) entries. */
void cec_synth_rotation(uint64_t seed, unsigned dim, double *out)
{
    uint64_t s = seed;
    const size_t d = dim;
    for (size_t i = 0; i < d * d; ++i) out[i] = 2.0 * u01(&s) - 1.0;
    for (size_t i = 0; i < d; ++i) {
        double *ri = out + i * d;
        for (int pass = 0; pass < 2; ++pass) {
            for (size_t k = 0; k < i; ++k) {
                const double *rk = out + k * d;
                double dot = 0.0;
                for (size_t j = 0; j < d; ++j) dot += ri[j] * rk[j];
                for (size_t j = 0; j < d; ++j) ri[j] -= dot * rk[j];
            }
        }
        double nrm = 0.0;
        for (size_t j = 0; j < d; ++j) nrm += ri[j] * ri[j];
        nrm = sqrt(nrm);
        if (nrm == 0.0) { /* cannot happen for random input; keep the row valid anyway */
            for (size_t j = 0; j < d; ++j) ri[j] = (j == i) ? 1.0 : 0.0;
        } else {
            for (size_t j = 0; j < d; ++j) ri[j] /= nrm;
        }
    }
}

void cec_synth_perm(uint64_t seed, unsigned dim, int *out)
{
    uint64_t s = seed;
    for (unsigned i = 0; i < dim; ++i) out[i] = (int)i + 1;
    for (unsigned i = dim; i > 1; --i) { /* Fisher-Yates */
        const unsigned j = (unsigned)(sm64_next(&s) % i);
        const int t = out[i - 1];
        out[i - 1] = out[j];
        out[j] = t;
    }
}

enum { SUITE_2014_ROT = 20142, SUITE_2014_SHIFT = 20141, SUITE_2014_SHUF = 20143, SUITE_2013_ROT = 20132, SUITE_2013_SHIFT = 20131 };

void cec2014_synth_rotation(unsigned func, unsigned dim, double *out)
{
    for (unsigned k = 0; k < CEC_SYNTH_NCOMP; ++k)
        cec_synth_rotation(mix_seed(SUITE_2014_ROT, func, dim, k), dim, out + (size_t)k * dim * dim);
}

void cec2014_synth_shift(unsigned func, double *out)
{
    cec_synth_uniform(mix_seed(SUITE_2014_SHIFT, func, 0, 0), -80.0, 80.0, out, (size_t)CEC_SYNTH_NCOMP * 100);
}

void cec2014_synth_shuffle(unsigned func, unsigned dim, int *out)
{
    for (unsigned k = 0; k < CEC_SYNTH_NCOMP; ++k)
        cec_synth_perm(mix_seed(SUITE_2014_SHUF, func, dim, k), dim, out + (size_t)k * dim);
}

void cec2013_synth_md(unsigned dim, double *out)
{
    for (unsigned k = 0; k < CEC_SYNTH_NCOMP; ++k)
        cec_synth_rotation(mix_seed(SUITE_2013_ROT, 0, dim, k), dim, out + (size_t)k * dim * dim);
}

void cec2013_synth_shift(double *out)
{
    cec_synth_uniform(mix_seed(SUITE_2013_SHIFT, 0, 0, 0), -80.0, 80.0, out, (size_t)CEC_SYNTH_NCOMP * 100);
}
