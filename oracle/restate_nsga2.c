/* oracle/restate_nsga2.c - plain-C restatement of the NSGA-II generation operators and generation loop.
 * TEST INFRASTRUCTURE ONLY: the checker the CUDA path is compared against; never linked into the product.
 * Follows reference src/utils/genetic_operators.cpp:49-57 (sbx_betaq), :71-144 (sbx_crossover_impl), :148-197
 * (polynomial_mutation_impl), :200-211 (mo_tournament_selection_impl) and src/algorithms/nsga2.cpp:176-304.
 * PINNED (tests/test_oracle_pin.py): with the draw source in sequential-mt19937 mode (philox.h) these functions reproduce the
 * compiled reference - sbx_crossover_impl, polynomial_mutation_impl, mo_tournament_selection_impl and whole nsga2::evolve runs -
 * bit for bit.  In the default Philox mode the SAME statements take the draw the device consumes at the same
 * (generation, group, slot), and a shuffle is the stable argsort of Philox keys (the reference re-shuffles one persistent
 * index vector with std::shuffle, nsga2.cpp:138-139,180-181).  Continuous decision variables only (nix = 0).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "philox.h"

static double sbx_betaq(double beta, double eta_c, double rand01)
{
    double alpha = 2. - pow(beta, -(eta_c + 1.));
    if (rand01 < 1. / alpha) return pow(rand01 * alpha, 1. / (eta_c + 1.));
    return pow(1. / (2. - rand01 * alpha), 1. / (eta_c + 1.));
}

static size_t tournament(size_t i1, size_t i2, const size_t *rank, const double *cd, oracle_stream *rs)
{
    if (rank[i1] < rank[i2]) return i1;
    if (rank[i1] > rank[i2]) return i2;
    if (cd[i1] > cd[i2]) return i1;
    if (cd[i1] < cd[i2]) return i2;
    return (oracle_next(rs) < 0.5) ? i1 : i2;
}

/* integer alleles sit at the end of the chromosome (problem::get_nix()); set per thread before calling the operators below */
static _Thread_local size_t nsga2_nix = 0;
void oracle_nsga2_set_nix(size_t nix) { nsga2_nix = nix; }

/* uniform_integral_from_range(lb, ub) (generic.hpp:142-164): std::uniform_int_distribution<long long>; Philox mode: lb + floor(u * range) */
static double next_integral(oracle_stream *rs, double lb, double ub)
{
    const long long l = (long long)lb, u = (long long)ub;
    return (double)(l + (long long)oracle_next_below(rs, (size_t)(u - l + 1)));
}

static void sbx(const double *p1, const double *p2, double *c1, double *c2, size_t nx, const double *lb, const double *ub, double p_cr,
                double eta_c, oracle_stream *rs)
{
    const size_t nix = nsga2_nix <= nx ? nsga2_nix : nx, ncx = nx - nix;
    memcpy(c1, p1, nx * sizeof(double));
    memcpy(c2, p2, nx * sizeof(double));
    if (oracle_next(rs) < p_cr) {
        for (size_t i = 0; i < ncx; i++) {
            if ((oracle_next(rs) < 0.5) && (fabs(p1[i] - p2[i])) > 1e-14 && lb[i] != ub[i]) {
                double y1, y2, yl, yu, beta, betaq, v1, v2, rand01;
                if (p1[i] < p2[i]) { y1 = p1[i]; y2 = p2[i]; } else { y1 = p2[i]; y2 = p1[i]; }
                yl = lb[i];
                yu = ub[i];
                rand01 = oracle_next(rs);
                beta = 1. + (2. * (y1 - yl) / (y2 - y1));
                betaq = sbx_betaq(beta, eta_c, rand01);
                v1 = 0.5 * ((y1 + y2) - betaq * (y2 - y1));
                beta = 1. + (2. * (yu - y2) / (y2 - y1));
                betaq = sbx_betaq(beta, eta_c, rand01);
                v2 = 0.5 * ((y1 + y2) + betaq * (y2 - y1));
                if (v1 < lb[i]) v1 = lb[i];
                if (v2 < lb[i]) v2 = lb[i];
                if (v1 > ub[i]) v1 = ub[i];
                if (v2 > ub[i]) v2 = ub[i];
                if (oracle_next(rs) < .5) { c1[i] = v1; c2[i] = v2; } else { c1[i] = v2; c2[i] = v1; }
            }
        }
        if (nix > 0) { /* two-point crossover of the integer part, genetic_operators.cpp:125-137 */
            size_t site1 = ncx + oracle_next_below(rs, nix), site2 = ncx + oracle_next_below(rs, nix);
            if (site1 > site2) { const size_t t = site1; site1 = site2; site2 = t; }
            for (size_t j = site1; j <= site2; ++j) { c1[j] = p2[j]; c2[j] = p1[j]; }
        }
    }
}

static void polymut(double *child, size_t nx, const double *lb, const double *ub, double p_m, double eta_m, oracle_stream *rs)
{
    const size_t nix = nsga2_nix <= nx ? nsga2_nix : nx, ncx = nx - nix;
    for (size_t j = 0; j < ncx; ++j) {
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
    for (size_t j = ncx; j < nx; ++j) /* integer mutation, genetic_operators.cpp:187-195 */
        if (oracle_next(rs) < p_m) child[j] = next_integral(rs, lb[j], ub[j]);
}

struct kv { uint64_t k; size_t v; };
static int kv_cmp(const void *a, const void *b)
{
    const struct kv *x = (const struct kv *)a, *y = (const struct kv *)b;
    if (x->k != y->k) return x->k < y->k ? -1 : 1;
    return x->v < y->v ? -1 : (x->v > y->v);
}

/* permutation = stable argsort of the Philox keys (stands in for std::shuffle, nsga2.cpp:180-181) */
int oracle_philox_perm(size_t n, uint64_t seed, uint32_t tag, uint32_t generation, size_t *perm)
{
    struct kv *a = (struct kv *)malloc((n ? n : 1) * sizeof(struct kv));
    for (size_t i = 0; i < n; ++i) { a[i].k = oracle_philox_u64(seed, tag, generation, (uint32_t)i, 0); a[i].v = i; }
    qsort(a, n, sizeof(struct kv), kv_cmp);
    for (size_t i = 0; i < n; ++i) perm[i] = a[i].v;
    free(a);
    return 0;
}

double oracle_philox_u01_at(uint64_t seed, uint32_t tag, uint32_t generation, uint32_t index, uint32_t slot)
{
    return oracle_philox_u01(seed, tag, generation, index, slot);
}

void oracle_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { oracle_philox4x32_10(ctr, key, out); }

/* nsga2.cpp:215-239 */
int oracle_nsga2_variation(const double *x, const size_t *rank, const double *cd, size_t NP, size_t nx, const double *lb,
                           const double *ub, const size_t *sh1, const size_t *sh2, double cr, double eta_c, double m, double eta_m,
                           uint64_t seed, uint32_t generation, double *children)
{
    if (NP < 5 || NP % 4) return -1;
    for (size_t i = 0; i < NP; i += 4) {
        oracle_stream rs = {seed, ORACLE_TAG_NSGA2_VAR, generation, (uint32_t)(i / 4), 0};
        for (int half = 0; half < 2; ++half) {
            const size_t *sh = half ? sh2 : sh1;
            const size_t a = tournament(sh[i], sh[i + 1], rank, cd, &rs), b = tournament(sh[i + 2], sh[i + 3], rank, cd, &rs);
            double *c1 = children + (i + 2 * half) * nx, *c2 = c1 + nx;
            sbx(x + a * nx, x + b * nx, c1, c2, nx, lb, ub, cr, eta_c, &rs);
            polymut(c1, nx, lb, ub, m, eta_m, &rs);
            polymut(c2, nx, lb, ub, m, eta_m, &rs);
        }
    }
    return 0;
}

/* crowding distance of the whole population with nsga2's small-front rule, nsga2.cpp:184-206 */
int oracle_nsga2_rank_crowding(const double *f, size_t NP, size_t nobj, size_t *rank, double *cd)
{
    size_t *fi = (size_t *)malloc(NP * sizeof(size_t)), *fo = (size_t *)malloc((NP + 1) * sizeof(size_t)), nf = 0;
    if (oracle_fnds(f, NP, nobj, rank, NULL, fi, fo, &nf)) return -1;
    for (size_t k = 0; k < nf; ++k) {
        const size_t b = fo[k], sz = fo[k + 1] - b;
        if (sz <= 2) {
            for (size_t i = 0; i < sz; ++i) cd[fi[b + i]] = INFINITY;
        } else {
            double *sub = (double *)malloc(sz * nobj * sizeof(double)), *c2 = (double *)malloc(sz * sizeof(double));
            for (size_t i = 0; i < sz; ++i) memcpy(sub + i * nobj, f + fi[b + i] * nobj, nobj * sizeof(double));
            if (oracle_crowding_distance(sub, sz, nobj, c2)) return -1;
            for (size_t i = 0; i < sz; ++i) cd[fi[b + i]] = c2[i];
            free(sub); free(c2);
        }
    }
    free(fi); free(fo);
    return 0;
}

/* nsga2::evolve, :176-304, for the restated ZDT (family 8) / DTLZ (family 9) problems */
int oracle_nsga2_evolve(int family, unsigned prob_id, size_t nx, size_t nobj, unsigned alpha, const double *lb, const double *ub,
                        double *x, double *f, size_t NP, unsigned gens, double cr, double eta_c, double m, double eta_m, uint64_t seed,
                        uint32_t first_generation)
{
    if (NP < 5 || NP % 4 || nobj < 2) return -1;
    double *x2 = (double *)malloc(2 * NP * nx * sizeof(double)), *f2 = (double *)malloc(2 * NP * nobj * sizeof(double));
    double *cd = (double *)malloc(NP * sizeof(double));
    size_t *rank = (size_t *)malloc(NP * sizeof(size_t)), *sh1 = (size_t *)malloc(NP * sizeof(size_t)),
           *sh2 = (size_t *)malloc(NP * sizeof(size_t)), *sel = (size_t *)malloc(2 * NP * sizeof(size_t)), nsel = 0;
    int rc = 0;
    for (size_t i = 0; i < NP; ++i) sh1[i] = sh2[i] = i; /* nsga2.cpp:138-139 */
    for (unsigned g = 0; g < gens && !rc; ++g) {
        const uint32_t generation = first_generation + g;
        memcpy(x2, x, NP * nx * sizeof(double));
        memcpy(f2, f, NP * nobj * sizeof(double));
        if (oracle_mt_active) { /* nsga2.cpp:180-181: the index vectors persist across generations */
            oracle_mt_shuffle(oracle_mt_active, sh1, NP);
            oracle_mt_shuffle(oracle_mt_active, sh2, NP);
        } else {
            oracle_philox_perm(NP, seed, ORACLE_TAG_SHUFFLE1, generation, sh1);
            oracle_philox_perm(NP, seed, ORACLE_TAG_SHUFFLE2, generation, sh2);
        }
        if ((rc = oracle_nsga2_rank_crowding(f, NP, nobj, rank, cd))) break;
        if ((rc = oracle_nsga2_variation(x, rank, cd, NP, nx, lb, ub, sh1, sh2, cr, eta_c, m, eta_m, seed, generation, x2 + NP * nx))) break;
        rc = (family == 8) ? oracle_zdt_batch(prob_id, x2 + NP * nx, NP, nx, f2 + NP * nobj)
                           : oracle_dtlz_batch(prob_id, x2 + NP * nx, NP, nx, nobj, alpha, f2 + NP * nobj);
        if (rc) break;
        if ((rc = oracle_select_best_N_mo(f2, 2 * NP, nobj, NP, sel, &nsel))) break;
        for (size_t i = 0; i < NP; ++i) {
            memcpy(x + i * nx, x2 + sel[i] * nx, nx * sizeof(double));
            memcpy(f + i * nobj, f2 + sel[i] * nobj, nobj * sizeof(double));
        }
    }
    free(x2); free(f2); free(cd); free(rank); free(sh1); free(sh2); free(sel);
    return rc;
}

/* ---- sequential-mt19937 entry points: the draws of a std::mt19937 seeded with `seed`, as the reference consumes them ---- */
_Thread_local oracle_mt *oracle_mt_active = NULL;

int oracle_nsga2_evolve_mt(int family, unsigned prob_id, size_t nx, size_t nobj, unsigned alpha, const double *lb, const double *ub,
                           double *x, double *f, size_t NP, unsigned gens, double cr, double eta_c, double m, double eta_m, uint32_t seed)
{
    ORACLE_MT_BEGIN(seed);
    const int rc = oracle_nsga2_evolve(family, prob_id, nx, nobj, alpha, lb, ub, x, f, NP, gens, cr, eta_c, m, eta_m, 0, 0);
    ORACLE_MT_END();
    return rc;
}

/* one sbx_crossover_impl call followed by polynomial_mutation_impl on both children, then `n_tournaments` tournaments of
 * (i, i+1) pairs - all on ONE engine, so the draw position carries from one operator into the next */
int oracle_genetic_operators_mt(const double *p1, const double *p2, size_t nx, const double *lb, const double *ub, double p_cr, double eta_c,
                                double p_m, double eta_m, const size_t *rank, const double *cd, size_t n_pairs, uint32_t seed, double *c1,
                                double *c2, size_t *winners)
{
    ORACLE_MT_BEGIN(seed);
    oracle_stream rs = {0, 0, 0, 0, 0};
    sbx(p1, p2, c1, c2, nx, lb, ub, p_cr, eta_c, &rs);
    polymut(c1, nx, lb, ub, p_m, eta_m, &rs);
    polymut(c2, nx, lb, ub, p_m, eta_m, &rs);
    for (size_t i = 0; i < n_pairs; ++i) winners[i] = tournament(2 * i, 2 * i + 1, rank, cd, &rs);
    ORACLE_MT_END();
    return 0;
}

/* the helpers of mt19937.h, exported so that they can be compared with the real std:: classes (ref_std_*, ref_capi.h) */
int oracle_mt_sequence(uint32_t seed, int kind, uint64_t a, uint64_t b, size_t n, double *out_real, uint64_t *out_int)
{
    oracle_mt mt;
    oracle_mt_seed(&mt, seed);
    for (size_t i = 0; i < n; ++i) switch (kind) {
            case 0: out_int[i] = oracle_mt_u32(&mt); break;
            case 1: out_real[i] = oracle_mt_u01(&mt); break;
            case 2: out_int[i] = oracle_mt_int(&mt, a, b); break;
            case 3: out_real[i] = oracle_mt_normal(&mt, 0., 1.); break;
            case 4: out_real[i] = oracle_mt_real(&mt, -(double)a, (double)b); break;
            default: return -1;
        }
    return 0;
}

int oracle_mt_shuffles(uint32_t seed, size_t n, size_t rounds, size_t *perm)
{
    oracle_mt mt;
    oracle_mt_seed(&mt, seed);
    for (size_t i = 0; i < n; ++i) perm[i] = i;
    for (size_t r = 0; r < rounds; ++r) oracle_mt_shuffle(&mt, perm, n);
    return 0;
}
