/* restate_moead_gen.c - TEST INFRASTRUCTURE ONLY (see oracle.h): plain-C restatement of pagmo::moead_gen::evolve,
 * reference src/algorithms/moead_gen.cpp:128-345 (select_parents :395-426).
 *
 * moead_gen is the reference's own generational MOEA/D: all NP candidates of a generation are built from the population as the
 * previous generation left it (neighbourhood / whole-population parent selection, DE/rand/1 with binomial crossover, the
 * reference's bound repair, polynomial mutation), evaluated as ONE batch, and only then inserted one after the other - ideal point
 * update, replacement of the own sub-problem, then of up to `limit` sub-problems of a shuffled neighbourhood.
 * The weight vectors (decomposition_weights) and the neighbourhoods (kNN of the weights) are INPUTS here: they are utilities outside
 * evolve(), and the tests take them from the compiled reference itself.
 * Draws, in the reference's order, per generation: std::shuffle of the (persistent) index vector; then per individual n of that
 * order: the diversity draw, the parent picks (uniform_int(0, NP - 1), repeated until two distinct parents), one draw per gene for
 * the crossover (plus one per violated bound), the polynomial mutation's draws; after the batch evaluation, per individual, one
 * std::shuffle of its neighbourhood (or of 0 .. NP-1).  The draw source is dispatched (philox.h):
 *   mt19937 mode - the reference's own sequence, bit for bit (oracle_moead_gen_evolve_mt);
 *   Philox mode  - the generation's order is the stable argsort of Philox keys (tag MOEAD_ORDER), individual n draws from the
 *                  substream (seed, MOEAD, generation, n), and the neighbourhood shuffle of the individual at position q is the
 *                  stable argsort of the keys (seed, MOEAD_INSERT, generation, q, slot = element). */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "philox.h"

static void polymut(double *child, size_t nx, const double *lb, const double *ub, double p_m, double eta_m, oracle_stream *rs)
{ /* polynomial_mutation_impl, genetic_operators.cpp:148-197 (continuous genes) */
    for (size_t j = 0; j < nx; ++j) {
        if (oracle_next(rs) < p_m && lb[j] != ub[j]) {
            double y = child[j], yl = lb[j], yu = ub[j], deltaq, xy, val;
            const double delta1 = (y - yl) / (yu - yl), delta2 = (yu - y) / (yu - yl);
            const double rnd = oracle_next(rs), mut_pow = 1. / (eta_m + 1.);
            if (rnd < 0.5) {
                xy = 1. - delta1;
                val = 2. * rnd + (1. - 2. * rnd) * (pow(xy, (eta_m + 1.)));
                deltaq = pow(val, mut_pow) - 1.;
            } else {
                xy = 1. - delta2;
                val = 2. * (1. - rnd) + 2. * (rnd - 0.5) * (pow(xy, (eta_m + 1.)));
                deltaq = 1. - (pow(val, mut_pow));
            }
            y = y + deltaq * (yu - yl);
            if (y < yl) y = yl;
            if (y > yu) y = yu;
            child[j] = y;
        }
    }
}

struct kv { uint64_t k; size_t v; };
static int kv_cmp(const void *a, const void *b)
{
    const struct kv *x = (const struct kv *)a, *y = (const struct kv *)b;
    if (x->k != y->k) return x->k < y->k ? -1 : 1;
    return x->v < y->v ? -1 : (x->v > y->v);
}
/* Philox stand-in for std::shuffle of iota(n): stable argsort of the keys (seed, tag, generation, index, slot = element) */
static void philox_perm(size_t n, uint64_t seed, uint32_t tag, uint32_t generation, uint32_t index, size_t *perm, struct kv *scratch)
{
    for (size_t i = 0; i < n; ++i) {
        scratch[i].k = oracle_philox_u64(seed, tag, generation, index, (uint32_t)i);
        scratch[i].v = i;
    }
    qsort(scratch, n, sizeof(struct kv), kv_cmp);
    for (size_t i = 0; i < n; ++i) perm[i] = scratch[i].v;
}

int oracle_moead_gen_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t dim, size_t m,
                            unsigned gens, const double *weights /* [NP x m] */, const size_t *neigh /* [NP x T] */, size_t T,
                            int decomposition /* 0 weighted, 1 tchebycheff, 2 bi */, double CR, double F, double eta_m, double realb, unsigned limit,
                            int preserve_diversity, uint64_t seed, uint32_t first_generation, size_t burn_draws)
{
    if (NP == 0 || m < 2 || T < 2 || T > NP - 1 || decomposition < 0 || decomposition > 2) return -1;
    if (gens == 0) return 0;
    int rc = 0;
    double *ideal = (double *)malloc(m * sizeof(double)), *cand = (double *)malloc(NP * dim * sizeof(double)),
           *fnew = (double *)malloc(NP * m * sizeof(double));
    size_t *shuffle = (size_t *)malloc(NP * sizeof(size_t)), *shuffle2 = (size_t *)malloc(NP * sizeof(size_t));
    unsigned char *whole = (unsigned char *)malloc(NP);
    struct kv *scratch = (struct kv *)malloc(NP * sizeof(struct kv));
    if (oracle_mt_active)
        for (size_t k = 0; k < burn_draws; ++k) (void)oracle_mt_u01(oracle_mt_active); /* decomposition_weights("random") drew first, :155 */
    for (size_t k = 0; k < m; ++k) { /* ideal(pop.get_f()), :171: first minimum under less_than_f */
        size_t b = 0;
        for (size_t i = 1; i < NP; ++i) {
            const double a = f[i * m + k], c = f[b * m + k];
            if (!isnan(a) && (isnan(c) || a < c)) b = i;
        }
        ideal[k] = f[b * m + k];
    }
    for (size_t i = 0; i < NP; ++i) shuffle[i] = i; /* :173-174: persistent across generations */
    for (unsigned g = 0; g < gens && !rc; ++g) {
        const uint32_t generation = first_generation + g;
        /* 1 - shuffle the population indexes, :213 */
        if (oracle_mt_active) oracle_mt_shuffle(oracle_mt_active, shuffle, NP);
        else philox_perm(NP, seed, ORACLE_TAG_MOEAD_ORDER, generation, 0, shuffle, scratch);
        /* 2 - the candidates, :227-269 */
        for (size_t q = 0; q < NP; ++q) {
            const size_t n = shuffle[q];
            oracle_stream rs = {seed, ORACLE_TAG_MOEAD, generation, (uint32_t)n, 0};
            /* 3 - neighbourhood or whole population (the draw is taken whatever m_preserve_diversity says) */
            const double u = oracle_next(&rs);
            whole[q] = !(u < realb || !preserve_diversity);
            /* 4 - two distinct parents, select_parents :395-426 */
            size_t parents[2], np_ = 0;
            while (np_ < 2) {
                const size_t r = oracle_next_below(&rs, NP);
                const size_t p = whole[q] ? r : neigh[n * T + r % T];
                if (np_ == 1 && parents[0] == p) continue;
                parents[np_++] = p;
            }
            /* 5 - DE/rand/1, binomial crossover, the reference's bound repair */
            double *c = cand + q * dim;
            const double *xn = x + n * dim, *x0 = x + parents[0] * dim, *x1 = x + parents[1] * dim;
            for (size_t kk = 0; kk < dim; ++kk) {
                if (oracle_next(&rs) < CR) {
                    c[kk] = xn[kk] + F * (x0[kk] - x1[kk]);
                    if (c[kk] < lb[kk]) c[kk] = lb[kk] + oracle_next(&rs) * (xn[kk] - lb[kk]);
                    if (c[kk] > ub[kk]) c[kk] = ub[kk] - oracle_next(&rs) * (ub[kk] - xn[kk]);
                } else {
                    c[kk] = xn[kk];
                }
            }
            /* 6 - polynomial mutation, p_m = 1 / dim */
            polymut(c, dim, lb, ub, 1.0 / (double)dim, eta_m, &rs);
        }
        if ((rc = oracle_problem_eval(prob, cand, NP, fnew))) break; /* the bfe branch, :270-288 */
        /* 8, 9 - insertion, one individual after the other, :297-344 */
        for (size_t q = 0; q < NP; ++q) {
            const size_t n = shuffle[q];
            const double *nf = fnew + q * m, *c = cand + q * dim;
            for (size_t j = 0; j < m; ++j)
                if (nf[j] < ideal[j]) ideal[j] = nf[j]; /* std::min(new_f[j], ideal[j]) twice, :303-307: min(a, b) = (b < a) ? b : a */
            unsigned time = 0;
            double f1, f2;
            oracle_decompose_objectives(f + n * m, m, weights + n * m, ideal, decomposition, &f1);
            oracle_decompose_objectives(nf, m, weights + n * m, ideal, decomposition, &f2);
            if (f2 < f1) {
                memcpy(x + n * dim, c, dim * sizeof(double));
                memcpy(f + n * m, nf, m * sizeof(double));
                ++time;
            }
            const size_t size = whole[q] ? NP : T;
            if (oracle_mt_active) {
                for (size_t k = 0; k < size; ++k) shuffle2[k] = k;
                oracle_mt_shuffle(oracle_mt_active, shuffle2, size);
            } else {
                philox_perm(size, seed, ORACLE_TAG_MOEAD_INSERT, generation, (uint32_t)q, shuffle2, scratch);
            }
            for (size_t k = 0; k < size; ++k) {
                const size_t pick = whole[q] ? shuffle2[k] : neigh[n * T + shuffle2[k]];
                oracle_decompose_objectives(f + pick * m, m, weights + pick * m, ideal, decomposition, &f1);
                oracle_decompose_objectives(nf, m, weights + pick * m, ideal, decomposition, &f2);
                if (f2 < f1) {
                    memcpy(x + pick * dim, c, dim * sizeof(double));
                    memcpy(f + pick * m, nf, m * sizeof(double));
                    ++time;
                }
                if (time >= limit && preserve_diversity) break;
            }
        }
    }
    free(ideal); free(cand); free(fnew); free(shuffle); free(shuffle2); free(whole); free(scratch);
    return rc;
}

/* moead_gen::evolve on the reference's own stream: std::mt19937(seed); burn_draws = the draws decomposition_weights("random")
 * took from the same engine before the loop ((NP - m) * (m - 1)), 0 for "grid" and "low discrepancy" */
int oracle_moead_gen_evolve_mt(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t dim,
                               size_t m, unsigned gens, const double *weights, const size_t *neigh, size_t T, int decomposition, double CR, double F,
                               double eta_m, double realb, unsigned limit, int preserve_diversity, uint32_t seed, size_t burn_draws)
{
    ORACLE_MT_BEGIN(seed);
    const int rc = oracle_moead_gen_evolve(prob, lb, ub, x, f, NP, dim, m, gens, weights, neigh, T, decomposition, CR, F, eta_m, realb, limit,
                                           preserve_diversity, 0, 0, burn_draws);
    ORACLE_MT_END();
    return rc;
}
