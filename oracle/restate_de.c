/* oracle/restate_de.c - plain-C restatement of pagmo's differential-evolution family as the GENERATIONAL loop the device runs.
 * TEST INFRASTRUCTURE ONLY.  Follows reference src/algorithms/de.cpp:76-345, sade.cpp:78-560, de1220.cpp:80-600:
 * index selection (Durstenfeld, de1220.cpp:181-187), variant / F / CR adaptation (:192-197 and the per-variant iDE formulas),
 * the 18 mutation + crossover variants (:199-505; de.cpp:154-275 for de's own forms of variants 4,5,9,10), feasibility
 * (:507-513), selection and global-best bookkeeping (:515-536), exit conditions (de.cpp:302-321).
 * Two modes of the SAME statements:
 *  - oracle_de_evolve_mt: the reference's sequential mt19937 stream and its one-individual-at-a-time order (evaluate and select
 *    inside the loop, so the adapted F / CR / variant of an individual are visible to the ones after it, sade.cpp:504-516).
 *    PINNED: reproduces the compiled de / sade / de1220 ::evolve bit for bit (tests/test_oracle_pin.py).
 *  - oracle_de_evolve (what the device is compared with): the GENERATIONAL loop - all trials are built from the previous
 *    generation, evaluated in one batch, then selected - with every draw taken from the Philox substream
 *    (seed, TAG_DE, generation, i) in the reference's per-individual order; uniform_int(0,n-1) = floor(u*n), normal = Box-Muller.
 *    For de, and for sade / de1220 with variant_adptv = 1, the two loops are the same function of the draws (a trial reads only
 *    the previous generation and its own F / CR); with variant_adptv = 2 the generational loop reads the other individuals'
 *    F / CR as they were at the start of the generation.
 * The normal draws inside one reference expression are evaluated left to right, F before CR (GCC's order; pinned by the test).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "philox.h"

#define normal01(rs) oracle_next_normal(rs)
#define uint_below(rs, n) ((unsigned)oracle_next_below((rs), (n)))

static double mutate(unsigned algo, unsigned base, double t, double gb, const double *p, double pi, double F)
{
    switch (base) {
        case 1: return gb + F * (p[1] - p[2]);
        case 2: return p[0] + F * (p[1] - p[2]);
        case 3: return t + F * (gb - t) + F * (p[0] - p[1]);
        case 4: return algo == 0 ? gb + (p[0] + p[1] - p[2] - p[3]) * F : gb + (p[0] - p[1]) * F + (p[2] - p[3]) * F;
        case 5: return algo == 0 ? p[4] + (p[0] + p[1] - p[2] - p[3]) * F : p[4] + (p[0] - p[1]) * F + (p[2] - p[3]) * F;
        case 6: return p[0] + (p[1] - p[2]) * F + (p[3] - p[4]) * F + (p[5] - p[6]) * F;
        case 7: return gb + (p[1] - p[2]) * F + (p[3] - p[4]) * F + (p[5] - p[6]) * F;
        case 8: return p[0] + (p[1] - pi) * F + (p[2] - p[3]) * F;
        default: return p[0] + (p[1] - pi) * F - (p[2] - gb) * F;
    }
}

static int de_evolve_impl(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t dim,
                          unsigned gens, unsigned algo, unsigned variant_in, unsigned variant_adptv, double F0, double CR0,
                          const unsigned *allowed, unsigned n_allowed, double ftol, double xtol, uint64_t seed, uint32_t first_generation,
                          unsigned *gens_done, double *F_state, double *CR_state, unsigned *variant_state, int sequential)
{
    if (gens_done) *gens_done = 0;
    if (gens == 0) return 0;
    if (algo > 2 || NP < (algo == 0 ? 5u : 7u)) return -1;
    double *trial = (double *)malloc(NP * dim * sizeof(double)), *ftrial = (double *)malloc(NP * sizeof(double));
    double *gbIter = (double *)malloc(dim * sizeof(double));
    double *mF = (double *)malloc(NP * sizeof(double)), *mC = (double *)malloc(NP * sizeof(double));
    double *Ftry = (double *)malloc(NP * sizeof(double)), *Ctry = (double *)malloc(NP * sizeof(double));
    unsigned *mV = (unsigned *)malloc(NP * sizeof(unsigned)), *Vtry = (unsigned *)malloc(NP * sizeof(unsigned));
    double *xnew = sequential ? (double *)malloc(NP * dim * sizeof(double)) : NULL;
    if (algo != 0) { /* de1220.cpp:147-165 (no memory): F / CR of every individual first, then (de1220) every variant */
        oracle_stream *irs = (oracle_stream *)malloc(NP * sizeof(oracle_stream));
        for (size_t i = 0; i < NP; ++i) {
            const oracle_stream rs0 = {seed, ORACLE_TAG_INIT, first_generation, (uint32_t)i, 0};
            oracle_stream *rs = &irs[i];
            *rs = rs0;
            if (variant_adptv == 1) {
                const double c = oracle_next(rs), ff = oracle_next(rs);
                mC[i] = c;
                mF[i] = ff * 0.9 + 0.1;
            } else {
                const double c = normal01(rs), ff = normal01(rs);
                mC[i] = c * 0.15 + 0.5;
                mF[i] = ff * 0.15 + 0.5;
            }
        }
        if (algo == 2)
            for (size_t i = 0; i < NP; ++i) mV[i] = allowed[uint_below(&irs[i], n_allowed)];
        free(irs);
    }
    /* global best: pop.best_idx() = first minimum */
    size_t gbidx = 0;
    for (size_t i = 1; i < NP; ++i)
        if (f[i] < f[gbidx]) gbidx = i;
    double gbfit = f[gbidx], gbF = algo ? mF[0] : 0, gbCR = algo ? mC[0] : 0;
    memcpy(gbIter, x + gbidx * dim, dim * sizeof(double));
    int rc = 0;
    unsigned done = 0;
    if (sequential) memcpy(xnew, x, NP * dim * sizeof(double));
/* de.cpp:281-297, sade.cpp:504-521, de1220.cpp:515-536; dst = the array of the NEW generation */
#define SELECT(i, dst)                                                                                                 \
    do {                                                                                                               \
        if (ftrial[i] <= f[i]) {                                                                                       \
            f[i] = ftrial[i];                                                                                          \
            memcpy((dst) + (i) * dim, trial + (i) * dim, dim * sizeof(double));                                        \
            if (algo) { mC[i] = Ctry[i]; mF[i] = Ftry[i]; }                                                            \
            if (algo == 2) mV[i] = Vtry[i];                                                                            \
            if (ftrial[i] <= gbfit) {                                                                                  \
                gbfit = ftrial[i];                                                                                     \
                gbidx = (i);                                                                                           \
                gbF = Ftry[i];                                                                                         \
                gbCR = Ctry[i];                                                                                        \
            }                                                                                                          \
        }                                                                                                              \
    } while (0)
    for (unsigned g = 0; g < gens && !rc; ++g) {
        const uint32_t generation = first_generation + g;
        const double gbIterF = gbF, gbIterCR = gbCR;
        for (size_t i = 0; i < NP; ++i) {
            oracle_stream rs = {seed, ORACLE_TAG_DE, generation, (uint32_t)i, 0};
            const unsigned npick = algo == 0 ? 5u : 7u;
            size_t r[7] = {0, 0, 0, 0, 0, 0, 0};
            {
                size_t *idxs = (size_t *)malloc(NP * sizeof(size_t));
                for (size_t k = 0; k < NP; ++k) idxs[k] = k;
                for (unsigned j = 0; j < npick; ++j) {
                    const size_t idx = uint_below(&rs, (unsigned)(NP - j));
                    r[j] = idxs[idx];
                    const size_t t = idxs[idx];
                    idxs[idx] = idxs[NP - 1u - j];
                    idxs[NP - 1u - j] = t;
                }
                free(idxs);
            }
            double F = F0, CR = CR0;
            unsigned VARIANT = variant_in;
            if (algo == 2) VARIANT = (oracle_next(&rs) < 0.9) ? mV[i] : allowed[uint_below(&rs, n_allowed)];
            if (algo != 0 && variant_adptv == 1) {
                F = (oracle_next(&rs) < 0.9) ? mF[i] : oracle_next(&rs) * 0.9 + 0.1;
                CR = (oracle_next(&rs) < 0.9) ? mC[i] : oracle_next(&rs);
            }
            unsigned base;
            int expo;
            if (VARIANT <= 10u) { expo = VARIANT <= 5u; base = expo ? VARIANT : VARIANT - 5u; }
            else { expo = (VARIANT & 1u) != 0u; base = 6u + (VARIANT - 11u) / 2u; }
            if (algo != 0 && variant_adptv == 2) {
                const double gF = gbIterF, gC = gbIterCR;
                double a1, a2, a3, c1, c2;
                switch (base) {
                    case 1: a1 = normal01(&rs); c1 = normal01(&rs);
                        F = gF + a1 * 0.5 * (mF[r[1]] - mF[r[2]]); CR = gC + c1 * 0.5 * (mC[r[1]] - mC[r[2]]); break;
                    case 2: a1 = normal01(&rs); c1 = normal01(&rs);
                        F = mF[r[0]] + a1 * 0.5 * (mF[r[1]] - mF[r[2]]); CR = mC[r[0]] + c1 * 0.5 * (mC[r[1]] - mC[r[2]]); break;
                    case 3: a1 = normal01(&rs); a2 = normal01(&rs); c1 = normal01(&rs); c2 = normal01(&rs);
                        F = mF[i] + a1 * 0.5 * (gF - mF[i]) + a2 * 0.5 * (mF[r[0]] - mF[r[1]]);
                        CR = mC[i] + c1 * 0.5 * (gC - mC[i]) + c2 * 0.5 * (mC[r[0]] - mC[r[1]]); break;
                    case 4: a1 = normal01(&rs); a2 = normal01(&rs); c1 = normal01(&rs); c2 = normal01(&rs);
                        F = gF + a1 * 0.5 * (mF[r[0]] - mF[r[1]]) + a2 * 0.5 * (mF[r[2]] - mF[r[3]]);
                        CR = gC + c1 * 0.5 * (mC[r[0]] - mC[r[1]]) + c2 * 0.5 * (mC[r[2]] - mC[r[3]]); break;
                    case 5: a1 = normal01(&rs); a2 = normal01(&rs); c1 = normal01(&rs); c2 = normal01(&rs);
                        F = mF[r[4]] + a1 * 0.5 * (mF[r[0]] - mF[r[1]]) + a2 * 0.5 * (mF[r[2]] - mF[r[3]]);
                        CR = mC[r[4]] + c1 * 0.5 * (mC[r[0]] - mC[r[1]]) + c2 * 0.5 * (mC[r[2]] - mC[r[3]]); break;
                    case 6: a1 = normal01(&rs); a2 = normal01(&rs); a3 = normal01(&rs); c1 = normal01(&rs);
                        F = mF[r[0]] + a1 * 0.5 * (mF[r[1]] - mF[r[2]]) + a2 * 0.5 * (mF[r[3]] - mF[r[4]]) + a3 * 0.5 * (mF[r[5]] - mF[r[6]]);
                        CR = mC[r[4]] + c1 * 0.5 * (mC[r[0]] + mC[r[1]] - mC[r[2]] - mC[r[3]]); break;
                    case 7: a1 = normal01(&rs); a2 = normal01(&rs); a3 = normal01(&rs); c1 = normal01(&rs);
                        F = gF + a1 * 0.5 * (mF[r[1]] - mF[r[2]]) + a2 * 0.5 * (mF[r[3]] - mF[r[4]]) + a3 * 0.5 * (mF[r[5]] - mF[r[6]]);
                        CR = gC + c1 * 0.5 * (mC[r[0]] + mC[r[1]] - mC[r[2]] - mC[r[3]]); break;
                    case 8: a1 = normal01(&rs); a2 = normal01(&rs); c1 = normal01(&rs); c2 = normal01(&rs);
                        F = mF[r[0]] + a1 * 0.5 * (mF[r[1]] - mF[i]) + a2 * 0.5 * (mF[r[3]] - mF[r[4]]);
                        CR = mC[r[0]] + c1 * 0.5 * (mC[r[1]] - mC[i]) + c2 * 0.5 * (mC[r[3]] - mC[r[4]]); break;
                    default: a1 = normal01(&rs); a2 = normal01(&rs); c1 = normal01(&rs); c2 = normal01(&rs);
                        F = mF[r[0]] + a1 * 0.5 * (mF[r[1]] - mF[i]) - a2 * 0.5 * (mF[r[2]] - gF);
                        CR = mC[r[0]] + c1 * 0.5 * (mC[r[1]] - mC[i]) - c2 * 0.5 * (mC[r[3]] - gC);
                }
            }
            const double *xi = x + i * dim;
            double *tmp = trial + i * dim;
            memcpy(tmp, xi, dim * sizeof(double));
            size_t n = uint_below(&rs, (unsigned)dim);
#define GENE(nn)                                                                                                       \
    do {                                                                                                               \
        double p[7];                                                                                                   \
        for (int k = 0; k < 7; ++k) p[k] = x[r[k] * dim + (nn)];                                                       \
        tmp[nn] = mutate(algo, base, tmp[nn], gbIter[nn], p, xi[nn], F);                                               \
    } while (0)
            if (expo) {
                size_t L = 0;
                do {
                    GENE(n);
                    n = (n + 1u) % dim;
                    ++L;
                } while ((oracle_next(&rs) < CR) && (L < dim));
            } else {
                for (size_t L = 0; L < dim; ++L) {
                    if ((oracle_next(&rs) < CR) || L + 1u == dim) GENE(n);
                    n = (n + 1u) % dim;
                }
            }
            for (size_t j = 0; j < dim; ++j)
                if ((tmp[j] < lb[j]) || (tmp[j] > ub[j])) tmp[j] = (lb[j] == ub[j]) ? lb[j] : (ub[j] - lb[j]) * oracle_next(&rs) + lb[j];
            Ftry[i] = F;
            Ctry[i] = CR;
            Vtry[i] = VARIANT;
            if (sequential) { /* the reference: evaluate and select before the next individual is built (de.cpp:281-297) */
                if ((rc = oracle_problem_eval(prob, tmp, 1, ftrial + i))) break;
                SELECT(i, xnew);
            }
        }
        if (!sequential) {
            if ((rc = oracle_problem_eval(prob, trial, NP, ftrial))) break;
            for (size_t i = 0; i < NP; ++i) SELECT(i, x);
        } else {
            if (rc) break;
            memcpy(x, xnew, NP * dim * sizeof(double)); /* std::swap(popold, popnew) */
        }
        memcpy(gbIter, x + gbidx * dim, dim * sizeof(double));
        ++done;
        size_t best = 0, worst = 0;
        for (size_t i = 1; i < NP; ++i) {
            if (f[i] < f[best]) best = i;
            if (f[i] > f[worst]) worst = i;
        }
        double dx = 0.;
        for (size_t d = 0; d < dim; ++d) dx += fabs(x[worst * dim + d] - x[best * dim + d]);
        if (dx < xtol) break;
        if (fabs(f[worst] - f[best]) < ftol) break;
    }
    if (gens_done) *gens_done = done;
    if (F_state && algo) memcpy(F_state, mF, NP * sizeof(double));
    if (CR_state && algo) memcpy(CR_state, mC, NP * sizeof(double));
    if (variant_state && algo == 2) memcpy(variant_state, mV, NP * sizeof(unsigned));
    free(trial); free(ftrial); free(gbIter); free(mF); free(mC); free(Ftry); free(Ctry); free(mV); free(Vtry); free(xnew);
    return rc;
}

int oracle_de_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t dim,
                     unsigned gens, unsigned algo, unsigned variant_in, unsigned variant_adptv, double F0, double CR0,
                     const unsigned *allowed, unsigned n_allowed, double ftol, double xtol, uint64_t seed, uint32_t first_generation,
                     unsigned *gens_done, double *F_state, double *CR_state, unsigned *variant_state)
{
    return de_evolve_impl(prob, lb, ub, x, f, NP, dim, gens, algo, variant_in, variant_adptv, F0, CR0, allowed, n_allowed, ftol, xtol, seed,
                          first_generation, gens_done, F_state, CR_state, variant_state, 0);
}

/* Philox draws, but the reference's one-individual-at-a-time order (for the generational == sequential equivalence test) */
int oracle_de_evolve_sequential(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t dim,
                                unsigned gens, unsigned algo, unsigned variant_in, unsigned variant_adptv, double F0, double CR0,
                                const unsigned *allowed, unsigned n_allowed, double ftol, double xtol, uint64_t seed,
                                uint32_t first_generation, unsigned *gens_done, double *F_state, double *CR_state, unsigned *variant_state)
{
    return de_evolve_impl(prob, lb, ub, x, f, NP, dim, gens, algo, variant_in, variant_adptv, F0, CR0, allowed, n_allowed, ftol, xtol, seed,
                          first_generation, gens_done, F_state, CR_state, variant_state, 1);
}

/* de::evolve / sade::evolve / de1220::evolve on the reference's own stream and in its own order */
int oracle_de_evolve_mt(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t dim,
                        unsigned gens, unsigned algo, unsigned variant_in, unsigned variant_adptv, double F0, double CR0,
                        const unsigned *allowed, unsigned n_allowed, double ftol, double xtol, uint32_t seed, unsigned *gens_done)
{
    ORACLE_MT_BEGIN(seed);
    const int rc = de_evolve_impl(prob, lb, ub, x, f, NP, dim, gens, algo, variant_in, variant_adptv, F0, CR0, allowed, n_allowed, ftol, xtol,
                                  0, 0, gens_done, NULL, NULL, NULL, 1);
    ORACLE_MT_END();
    return rc;
}
