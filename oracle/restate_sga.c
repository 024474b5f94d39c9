/* oracle/restate_sga.c - plain-C restatement of pagmo::sga::evolve.  TEST INFRASTRUCTURE ONLY.
 * Follows reference src/algorithms/sga.cpp: evolve :184-292, perform_selection :341-383, perform_crossover :385-457 (sbx:
 * src/utils/genetic_operators.cpp:49-57,71-144), perform_mutation :459-547, force_bounds_stick src/utils/generic.cpp:171-183,
 * reinsertion (best NP of children followed by parents) :275-289.
 * Two drivers over the SAME operator arithmetic (sbx pair, exponential / binomial / single-point copy, the three per-gene
 * mutations, bound sticking, truncated / tournament winners, reinsertion):
 *  - oracle_sga_evolve_mt: the reference's statement order on its sequential mt19937 stream - one index array threaded through
 *    all tournaments (:359-368), all crossovers before all mutations, std::shuffle + binomial count choosing the genes to mutate
 *    (:489-491), std::sort's tie order.  PINNED: reproduces the compiled sga::evolve bit for bit (tests/test_oracle_pin.py).
 *  - oracle_sga_evolve (what the device is compared with): the generational per-individual form described in
 *    pagmo2_b200/csrc/sga.cu - every draw from the Philox substream (seed, TAG_SGA, generation, index), index j for the
 *    tournament of offspring j (a fresh index array each), NP + i for the variation of offspring (or sbx pair) i, each gene
 *    mutated with probability m (the same law as "binomial count of a shuffled index list"), stable sorts.
 * The mating partner rule :411-412 is the same in both.  Continuous decision vectors only.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "philox.h"

enum { X_EXP = 0, X_BIN = 1, X_SINGLE = 2, X_SBX = 3 };
enum { M_GAUSS = 0, M_UNIFORM = 1, M_POLY = 2 };

static int less_f(double a, double b)
{
    if (!isnan(a)) return !isnan(b) ? a < b : 1;
    return 0;
}

static double betaq_of(double beta, double eta_c, double rand01)
{
    const double alpha = 2. - pow(beta, -(eta_c + 1.));
    if (rand01 < (1. / alpha)) return pow(rand01 * alpha, 1. / (eta_c + 1.));
    return pow(1. / (2. - rand01 * alpha), 1. / (eta_c + 1.));
}

/* the arithmetic of one mutated gene, sga.cpp:497-540 */
static void mutate_gene(double *c, size_t g, const double *lb, const double *ub, double param_m, unsigned mutation, oracle_stream *rs)
{
    if (mutation == M_UNIFORM) {
        c[g] = (lb[g] == ub[g]) ? lb[g] : oracle_next(rs) * (ub[g] - lb[g]) + lb[g];
    } else if (mutation == M_GAUSS) {
        const double sd = (ub[g] - lb[g]) * param_m;
        c[g] += oracle_next_normal(rs) * sd;
    } else {
        const double u = oracle_next(rs);
        if (u <= 0.5) {
            const double delta_l = pow(2. * u, 1. / (1. + param_m)) - 1.;
            c[g] += delta_l * (c[g] - lb[g]);
        } else {
            const double delta_r = 1 - pow(2. * (1. - u), 1. / (1. + param_m));
            c[g] += delta_r * (ub[g] - c[g]);
        }
    }
}

static void stick(double *c, size_t nx, const double *lb, const double *ub) /* force_bounds_stick, generic.cpp:171-183 */
{
    for (size_t g = 0; g < nx; ++g) {
        if (c[g] < lb[g]) c[g] = lb[g];
        if (c[g] > ub[g]) c[g] = ub[g];
    }
}

/* per-individual form: every gene mutates with probability m */
static void mutate(double *c, size_t nx, const double *lb, const double *ub, double m, double param_m, unsigned mutation, oracle_stream *rs)
{
    for (size_t g = 0; g < nx; ++g) {
        if (oracle_next(rs) < m) mutate_gene(c, g, lb, ub, param_m, mutation, rs);
        if (c[g] < lb[g]) c[g] = lb[g];
        if (c[g] > ub[g]) c[g] = ub[g];
    }
}

/* sbx_crossover_impl on one pair, genetic_operators.cpp:71-144 */
static void sbx_pair(const double *p1, const double *p2, double *c1, double *c2, size_t nx, const double *lb, const double *ub, double cr,
                     double eta_c, oracle_stream *rs)
{
    memcpy(c1, p1, nx * sizeof(double));
    memcpy(c2, p2, nx * sizeof(double));
    if (oracle_next(rs) < cr) {
        for (size_t i = 0; i < nx; ++i) {
            const double a = p1[i], b = p2[i], yl = lb[i], yu = ub[i];
            if ((oracle_next(rs) < 0.5) && (fabs(a - b)) > 1e-14 && yl != yu) {
                const double y1 = (a < b) ? a : b, y2 = (a < b) ? b : a;
                const double rand01 = oracle_next(rs);
                double beta = 1. + (2. * (y1 - yl) / (y2 - y1));
                double betaq = betaq_of(beta, eta_c, rand01);
                double v1 = 0.5 * ((y1 + y2) - betaq * (y2 - y1));
                beta = 1. + (2. * (yu - y2) / (y2 - y1));
                betaq = betaq_of(beta, eta_c, rand01);
                double v2 = 0.5 * ((y1 + y2) + betaq * (y2 - y1));
                if (v1 < yl) v1 = yl;
                if (v2 < yl) v2 = yl;
                if (v1 > yu) v1 = yu;
                if (v2 > yu) v2 = yu;
                if (oracle_next(rs) < .5) { c1[i] = v1; c2[i] = v2; }
                else { c1[i] = v2; c2[i] = v1; }
            }
        }
    }
}

/* exponential / binomial / single-point copy of parent2 genes into child after the partner draw, sga.cpp:411-447 */
static void cross_child(double *child, const double *x, const size_t *sel, size_t t, size_t NP, size_t nx, double cr, unsigned crossover,
                        oracle_stream *rs)
{
    const size_t pidx = 1 + oracle_next_below(rs, NP - 1);
    const size_t partner = (pidx <= t) ? pidx - 1 : pidx; /* all_idx after swap(all_idx[0], all_idx[t]), :411-412 */
    const double *parent2 = x + sel[partner] * nx;
    if (crossover == X_EXP) {
        size_t n = oracle_next_below(rs, nx), L = 0;
        do {
            child[n] = parent2[n];
            n = (n + 1u) % nx;
            ++L;
        } while ((oracle_next(rs) < cr) && (L < nx));
    } else if (crossover == X_BIN) {
        size_t n = oracle_next_below(rs, nx);
        for (size_t L = 0; L < nx; ++L) {
            if ((oracle_next(rs) < cr) || L + 1 == nx) child[n] = parent2[n];
            n = (n + 1) % nx;
        }
    } else if (oracle_next(rs) < cr) {
        const size_t n = oracle_next_below(rs, nx);
        for (size_t k = n; k < nx; ++k) child[k] = parent2[k];
    }
}

static int before_fit(size_t a, size_t b, const void *ctx) { return less_f(((const double *)ctx)[a], ((const double *)ctx)[b]); }
static void order_by_fitness(const double *f, size_t n, size_t *idx, size_t *tmp)
{
    for (size_t i = 0; i < n; ++i) idx[i] = i;
    oracle_sort_indices(idx, tmp, n, before_fit, f);
}

static int sga_evolve_impl(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t nx,
                           unsigned gens, double cr, double eta_c, double m, double param_m, unsigned param_s, unsigned crossover,
                           unsigned mutation, unsigned selection, uint64_t seed, uint32_t first_generation, int reference_order)
{
    if (NP < 2 || param_s < 1 || param_s > NP || crossover > 3 || mutation > 2 || selection > 1) return -1;
    if (crossover == X_SBX && NP % 2) return -1;
    size_t *sel = (size_t *)malloc(NP * sizeof(size_t)), *order = (size_t *)malloc(2 * NP * sizeof(size_t));
    size_t *perm = (size_t *)malloc(NP * sizeof(size_t)), *virt = (size_t *)malloc(NP * sizeof(size_t));
    size_t *tmp = (size_t *)malloc(2 * NP * sizeof(size_t)), *tbm = (size_t *)malloc(nx * sizeof(size_t));
    double *xnew = (double *)malloc(NP * nx * sizeof(double)), *fboth = (double *)malloc(2 * NP * sizeof(double));
    double *xo = (double *)malloc(NP * nx * sizeof(double)), *fo = (double *)malloc(NP * sizeof(double));
    int rc = 0;
    for (unsigned g = 0; g < gens && !rc; ++g) {
        const uint32_t generation = first_generation + g;
        /* selection, sga.cpp:341-383 */
        if (selection == 1) {
            order_by_fitness(f, NP, order, tmp);
            for (size_t j = 0; j < NP; ++j) sel[j] = order[j % param_s];
        } else {
            if (reference_order)
                for (size_t k = 0; k < NP; ++k) virt[k] = k; /* ONE index array for all tournaments of the generation, :346-347 */
            for (size_t j = 0; j < NP; ++j) {
                oracle_stream rs = {seed, ORACLE_TAG_SGA, generation, (uint32_t)j, 0};
                if (!reference_order)
                    for (size_t k = 0; k < NP; ++k) virt[k] = k;
                for (unsigned i = 0; i < param_s; ++i) {
                    const size_t index = i + oracle_next_below(&rs, NP - i);
                    const size_t t = virt[index];
                    virt[index] = virt[i];
                    virt[i] = t;
                }
                size_t winner = virt[0];
                for (unsigned i = 1; i < param_s; ++i)
                    if (f[virt[i]] < f[winner]) winner = virt[i];
                sel[j] = winner;
            }
        }
        if (reference_order) {
            /* perform_crossover on all of XNEW, then perform_mutation on all of XNEW (:248-251) */
            oracle_stream rs = {seed, ORACLE_TAG_SGA, generation, 0, 0};
            if (crossover == X_SBX) {
                for (size_t k = 0; k < NP; ++k) perm[k] = k;
                oracle_mt_shuffle(oracle_mt_active, perm, NP); /* std::shuffle(X.begin(), X.end(), m_e), :398 */
                for (size_t t = 0; t < NP / 2; ++t)
                    sbx_pair(x + sel[perm[2 * t]] * nx, x + sel[perm[2 * t + 1]] * nx, xnew + 2 * t * nx, xnew + (2 * t + 1) * nx, nx, lb, ub, cr,
                             eta_c, &rs);
            } else {
                for (size_t t = 0; t < NP; ++t) {
                    double *child = xnew + t * nx;
                    memcpy(child, x + sel[t] * nx, nx * sizeof(double));
                    cross_child(child, x, sel, t, NP, nx, cr, crossover, &rs);
                }
            }
            for (size_t k = 0; k < nx; ++k) tbm[k] = k;        /* :483-484: one index list for all individuals */
            oracle_mt_active->saved_available = 0;            /* `normal` is a fresh distribution object per call, :479 */
            for (size_t t = 0; t < NP; ++t) {
                oracle_mt_shuffle(oracle_mt_active, tbm, nx);
                const uint64_t N = oracle_mt_binomial(oracle_mt_active, nx, m);
                if (N == (uint64_t)-1) { rc = -2; break; }
                for (uint64_t j = 0; j < N; ++j) mutate_gene(xnew + t * nx, tbm[j], lb, ub, param_m, mutation, &rs);
                stick(xnew + t * nx, nx, lb, ub);
            }
            if (rc) break;
        } else if (crossover == X_SBX) {
            if (oracle_philox_perm(NP, seed, ORACLE_TAG_SHUFFLE1, generation, perm)) { rc = -1; break; }
            for (size_t t = 0; t < NP / 2; ++t) {
                oracle_stream rs = {seed, ORACLE_TAG_SGA, generation, (uint32_t)(NP + t), 0};
                double *c1 = xnew + 2 * t * nx, *c2 = c1 + nx;
                sbx_pair(x + sel[perm[2 * t]] * nx, x + sel[perm[2 * t + 1]] * nx, c1, c2, nx, lb, ub, cr, eta_c, &rs);
                mutate(c1, nx, lb, ub, m, param_m, mutation, &rs);
                mutate(c2, nx, lb, ub, m, param_m, mutation, &rs);
            }
        } else {
            for (size_t t = 0; t < NP; ++t) {
                oracle_stream rs = {seed, ORACLE_TAG_SGA, generation, (uint32_t)(NP + t), 0};
                double *child = xnew + t * nx;
                memcpy(child, x + sel[t] * nx, nx * sizeof(double));
                cross_child(child, x, sel, t, NP, nx, cr, crossover, &rs);
                mutate(child, nx, lb, ub, m, param_m, mutation, &rs);
            }
        }
        /* evaluation and reinsertion, :253-289 */
        if ((rc = oracle_problem_eval(prob, xnew, NP, fboth))) break;
        memcpy(fboth + NP, f, NP * sizeof(double));
        order_by_fitness(fboth, 2 * NP, order, tmp);
        for (size_t j = 0; j < NP; ++j) {
            const size_t s = order[j];
            memcpy(xo + j * nx, (s < NP ? xnew + s * nx : x + (s - NP) * nx), nx * sizeof(double));
            fo[j] = fboth[s];
        }
        memcpy(x, xo, NP * nx * sizeof(double));
        memcpy(f, fo, NP * sizeof(double));
    }
    free(sel); free(order); free(perm); free(virt); free(tmp); free(tbm); free(xnew); free(fboth); free(xo); free(fo);
    return rc;
}

int oracle_sga_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t nx, unsigned gens,
                      double cr, double eta_c, double m, double param_m, unsigned param_s, unsigned crossover, unsigned mutation,
                      unsigned selection, uint64_t seed, uint32_t first_generation)
{
    return sga_evolve_impl(prob, lb, ub, x, f, NP, nx, gens, cr, eta_c, m, param_m, param_s, crossover, mutation, selection, seed,
                           first_generation, 0);
}

/* sga::evolve on the reference's own stream, in its own statement order and with its std::sort tie order */
int oracle_sga_evolve_mt(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t nx,
                         unsigned gens, double cr, double eta_c, double m, double param_m, unsigned param_s, unsigned crossover,
                         unsigned mutation, unsigned selection, uint32_t seed)
{
    ORACLE_MT_BEGIN(seed);
    const int rc = sga_evolve_impl(prob, lb, ub, x, f, NP, nx, gens, cr, eta_c, m, param_m, param_s, crossover, mutation, selection, 0, 0, 1);
    ORACLE_MT_END();
    return rc;
}

int oracle_mt_binomial_sequence(uint32_t seed, uint64_t t, double p, size_t n, uint64_t *out)
{
    oracle_mt mt;
    oracle_mt_seed(&mt, seed);
    for (size_t i = 0; i < n; ++i)
        if ((out[i] = oracle_mt_binomial(&mt, t, p)) == (uint64_t)-1) return -1;
    return 0;
}
