/* restate_gaco.c - TEST INFRASTRUCTURE ONLY (see oracle.h): plain-C restatement of pagmo::gaco::evolve (extended ant colony
 * optimisation), reference src/algorithms/gaco.cpp:104-445, penalty_computation :506-549, update_sol_archive :563-675,
 * pheromone_computation :690-796, generate_new_ants :812-875 - for unconstrained single-objective problems (the ones with a device
 * evaluator), memory = false.
 *
 * gaco is generational in the reference: all ants of a generation are sampled from the solution archive the previous generation
 * left and evaluated as one batch (the bfe branch, :288-320).  Draws, in the reference's order, per ant: one uniform (which kernel),
 * then per variable one normal deviate, redrawn up to ten times while the sample falls outside the box.  The normal distribution
 * object is passed to generate_new_ants BY VALUE (:812), so its spare deviate is forgotten between generations.  The draw source is
 * dispatched (philox.h): one Philox substream per (generation, ant) for the device comparison; std::mt19937 + libstdc++'s
 * distributions and std::sort tie order for the bit-exact pin against the compiled reference (oracle_gaco_evolve_mt).
 *
 * The algorithm's scalar members survive between evolve() calls even with memory = false (m_oracle, m_q, m_n_evalstop, m_n_impstop,
 * m_gen_mark, m_fevals): they travel in oracle_gaco_state. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "philox.h"
#include "std_sort.h"

static int less_f(double a, double b) { return !isnan(a) && (isnan(b) || a < b); } /* detail::less_than_f */

typedef struct {
    const double *key;
} key_ctx;
static int key_less(size_t a, size_t b, const void *c) { return less_f(((const key_ctx *)c)->key[a], ((const key_ctx *)c)->key[b]); }

/* penalty_computation with no constraints (m_res = 0), :506-549; also the archive's re-evaluation after an oracle update, :365-397 */
static double gaco_penalty(double fitness, double oracle, double res)
{
    double alpha = 0.0;
    const double diff = fabs(fitness - oracle);
    double penalty = 0.0;
    if (fitness > oracle && res < diff / 3.0) {
        alpha = (diff * (6.0 * sqrt(3.0) - 2.0) / (6.0 * sqrt(3)) - res) / (diff - res);
    } else if (fitness > oracle && res >= diff / 3.0 && res <= diff) {
        alpha = 1.0 - 1.0 / (2.0 * sqrt(diff / res));
    } else if (fitness > oracle && res > diff) {
        alpha = 1.0 / 2.0 * sqrt(diff / res);
    }
    if (fitness > oracle || res > 0.) {
        penalty = alpha * diff + (1 - alpha) * res;
    } else if (fitness <= oracle && res == 0.) {
        penalty = -diff;
    }
    return penalty;
}

void oracle_gaco_state_init(oracle_gaco_state *s, double q, double oracle_par)
{
    s->oracle = oracle_par;
    s->q = q;
    s->n_evalstop = 1;
    s->n_impstop = 1;
    s->gen_mark = 1;
    s->fevals = 0;
    s->counter = 0;
    s->memory = 0;
    s->archive = NULL;
    s->has_champion = 0;
    s->champion = 0.;
}

/* archive rows: [penalty | x (nx) | f (1)] */
int oracle_gaco_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t n, size_t nx, size_t nix,
                       unsigned gens, unsigned ker, double acc, unsigned threshold, unsigned n_gen_mark, unsigned impstop, unsigned evalstop,
                       double focus, uint64_t seed, uint32_t first_generation, oracle_gaco_state *st, unsigned *gens_done)
{
    if (gens_done) *gens_done = 0;
    if (n == 0 || gens == 0) return 0;
    if (st->memory) ++st->counter; /* :106-108 */
    if (n < 2 || ker < 2 || ker > n || acc < 0. || focus < 0. || threshold < 1 || (!st->memory && threshold > gens) || st->q < 0.) return -1;
    const int memory = st->memory;
    const unsigned counter = st->counter;
    const size_t row = 1 + nx + 1, ncx = nx - nix;
    double *arch = (double *)malloc(ker * row * sizeof(double)), *tmp_arch = (double *)malloc(ker * row * sizeof(double)),
           *pen = (double *)malloc(n * sizeof(double)), *sorted_pen = (double *)malloc(n * sizeof(double)),
           *tp = (double *)malloc(2 * ker * sizeof(double)), *omega = (double *)malloc(ker * sizeof(double)),
           *pc = (double *)malloc(ker * sizeof(double)), *sigma = (double *)malloc(nx * sizeof(double)),
           *ants = (double *)malloc(n * nx * sizeof(double)), *fnew = (double *)malloc(n * sizeof(double));
    size_t *sl = (size_t *)malloc(n * sizeof(size_t)), *slp = (size_t *)malloc(2 * ker * sizeof(size_t)),
           *nsl = (size_t *)malloc(2 * ker * sizeof(size_t)), *stmp = (size_t *)malloc((n > 2 * ker ? n : 2 * ker) * sizeof(size_t));
    /* the population's champion: the first best individual (population::update_champion replaces on strict improvement only) */
    double champ = f[0];
    for (size_t i = 1; i < n; ++i)
        if (less_f(f[i], champ)) champ = f[i];
    if (st->has_champion && less_f(st->champion, champ)) champ = st->champion; /* the population remembers better ants than it holds */
    if (memory && counter > 1) memcpy(arch, st->archive, ker * row * sizeof(double)); /* sol_archive = m_sol_archive, :223-225 */
    int rc = 0, stopped = 0;
    unsigned gen;
    for (gen = 1; gen <= gens && !rc; ++gen) {
        const uint32_t generation = first_generation + (gen - 1);
        const double champ_old = champ; /* popold's champion, :184 */
        if ((impstop != 0 && st->n_impstop >= impstop) || (evalstop != 0 && st->n_evalstop >= evalstop)) { /* :193-205: `return pop` */
            stopped = 1;
            break;
        }
        /* 1 - penalties, :209-212 */
        for (size_t i = 0; i < n; ++i) pen[i] = gaco_penalty(f[i], st->oracle, 0.0);
        /* 2 - the archive, :216-250 */
        for (size_t i = 0; i < n; ++i) sl[i] = i;
        key_ctx kc = {pen};
        oracle_sort_indices(sl, stmp, n, key_less, &kc); /* std::sort on the mt19937 pin, stable for the device comparison */
        if (gen == 1 && counter < 2) {
            for (size_t i = 0; i < ker; ++i) {
                arch[i * row] = pen[sl[i]];
                memcpy(arch + i * row + 1, x + sl[i] * nx, nx * sizeof(double));
                arch[i * row + 1 + nx] = f[sl[i]];
            }
        } else {
            for (size_t i = 0; i < n; ++i) sorted_pen[i] = pen[sl[i]]; /* std::sort of the values: the same multiset in order */
            /* update_sol_archive, :563-675 */
            if (sorted_pen[0] < arch[(ker - 1) * row]) {
                st->n_impstop = 1;
                for (size_t i = 0; i < ker; ++i) {
                    tp[i] = sorted_pen[i];
                    tp[ker + i] = arch[i * row];
                }
                for (size_t i = 0; i < 2 * ker; ++i) slp[i] = i;
                key_ctx kt = {tp};
                oracle_sort_indices(slp, stmp, 2 * ker, key_less, &kt);
                memcpy(tmp_arch, arch, ker * row * sizeof(double));
#define FROM_POP(dst, idx)                                                                                                                   \
    do {                                                                                                                                         \
        tmp_arch[(dst)*row] = tp[(idx)];                                                                                                         \
        memcpy(tmp_arch + (dst)*row + 1, x + sl[(idx)] * nx, nx * sizeof(double));                                                               \
        tmp_arch[(dst)*row + 1 + nx] = f[sl[(idx)]];                                                                                             \
    } while (0)
                size_t count = 0, n_new = 0;
                if (slp[0] < ker) FROM_POP(0, slp[0]);
                else ++count;
                nsl[n_new++] = 0;
                for (size_t j = 1; j < 2 * ker; ++j) {
                    if (fabs(tp[slp[j]] - tp[slp[count]]) < acc) {
                    } else {
                        ++count;
                        nsl[n_new++] = j;
                    }
                }
                for (size_t ii = 0; ii < ker && ii < n_new; ++ii) {
                    const size_t idx = slp[nsl[ii]];
                    if (idx < ker) FROM_POP(ii, idx);
                    else memcpy(tmp_arch + ii * row, arch + (idx - ker) * row, row * sizeof(double));
                }
#undef FROM_POP
                memcpy(arch, tmp_arch, ker * row * sizeof(double));
            } else {
                ++st->n_impstop;
            }
            if (st->n_evalstop == 1 || st->n_evalstop > 2) ++st->gen_mark;
            if (st->gen_mark > n_gen_mark) st->gen_mark = 1;
        }
        /* 4 - pheromone_computation, :690-796 */
        if (memory ? 1 : (gen == 1 || gen == threshold)) { /* with memory the weights are recomputed every generation, :732-752 */
            if (memory ? counter == threshold : gen == threshold) st->q = 0.01;
            double sum_omega = 0;
            for (unsigned l = 1; l <= ker; ++l) {
                const double omega_new = 1.0 / (st->q * ker * sqrt(2 * 3.141592653589793238462643383279502884))
                                         * exp(-pow(l - 1.0, 2) / (2.0 * pow(st->q, 2) * pow(ker, 2)));
                omega[l - 1] = omega_new;
                sum_omega += omega_new;
            }
            for (unsigned k = 0; k < ker; ++k) {
                double cumulative = 0;
                for (unsigned j = 0; j <= k; ++j) cumulative += omega[j] / sum_omega;
                pc[k] = cumulative;
            }
        }
        for (size_t h = 1; h <= nx; ++h) {
            double d_min = fabs(arch[h] - arch[row + h]), d_max = d_min;
            for (unsigned c = 0; c + 1 < ker; ++c)
                for (unsigned k = c + 1; k < ker; ++k) {
                    const double d = fabs(arch[c * row + h] - arch[k * row + h]);
                    if (d < d_min) d_min = d;
                    if (d > d_max) d_max = d;
                }
            if (focus != 0. && ((d_max - d_min) / (memory ? counter : gen) > (ub[h - 1] - lb[h - 1]) / focus)) { /* :778-784 */
                sigma[h - 1] = (ub[h - 1] - lb[h - 1]) / focus;
            } else if (h <= ncx) {
                sigma[h - 1] = (d_max - d_min) / st->gen_mark;
            } else {
                const double a = fmax((d_max - d_min) / st->gen_mark, 1.0 / st->gen_mark);
                sigma[h - 1] = fmax(a, (1.0 - 1.0 / (sqrt((double)(nx - ncx)))));
            }
        }
        /* 5 - generate_new_ants, :812-875 */
        if (oracle_mt_active) oracle_mt_active->saved_available = 0; /* a fresh copy of the normal distribution every generation */
        for (size_t j = 0; j < n; ++j) {
            oracle_stream rs = {seed, ORACLE_TAG_GACO, generation, (uint32_t)j, 0};
            const double number = oracle_next(&rs);
            size_t k_omega = 0;
            if (number <= pc[0]) k_omega = 0;
            else if (number > pc[ker - 2]) k_omega = ker - 1;
            else
                for (unsigned k = 1; k + 1 < ker; ++k)
                    if (number > pc[k - 1] && number <= pc[k]) k_omega = k;
            for (size_t h = 0; h < nx; ++h) {
                double g_h = arch[k_omega * row + 1 + h] + sigma[h] * oracle_next_normal(&rs);
                if (g_h < lb[h] || g_h > ub[h]) {
                    int iter_while = 0;
                    while ((g_h < lb[h] || g_h > ub[h]) && iter_while < 10) {
                        g_h = arch[k_omega * row + 1 + h] + sigma[h] * oracle_next_normal(&rs);
                        ++iter_while;
                    }
                    if (g_h < lb[h]) g_h = lb[h];
                    if (g_h > ub[h]) g_h = ub[h];
                }
                ants[j * nx + h] = (h >= ncx) ? round(g_h) : g_h;
            }
        }
        rc = oracle_problem_eval(prob, ants, n, fnew);
        if (rc) break;
        st->fevals += n;
        memcpy(x, ants, n * nx * sizeof(double));
        memcpy(f, fnew, n * sizeof(double));
        for (size_t i = 0; i < n; ++i)
            if (less_f(f[i], champ)) champ = f[i];
        /* :338-347 */
        if (!less_f(champ, champ_old)) ++st->n_evalstop;
        else st->n_evalstop = 1;
        /* the oracle parameter, :349-402 */
        if (arch[1 + nx] < st->oracle) {
            for (unsigned r = 0; r < ker; ++r) {
                if (r == 0) st->oracle = arch[1 + nx];
                arch[r * row] = gaco_penalty(arch[r * row + 1 + nx], st->oracle, 0.0);
            }
        }
    }
    if (gens_done) *gens_done = gen - 1;
    st->champion = champ;
    st->has_champion = 1;
    if (memory) memcpy(st->archive, arch, ker * row * sizeof(double)); /* m_sol_archive */
    /* the archive goes back into the population, :408-421 (memory = false only; not when a stopping criterion returned early) */
    if (!rc && !stopped && !memory)
        for (size_t i = 0; i < ker; ++i) {
            memcpy(x + i * nx, arch + i * row + 1, nx * sizeof(double));
            f[i] = arch[i * row + 1 + nx];
        }
    free(arch); free(tmp_arch); free(pen); free(sorted_pen); free(tp); free(omega); free(pc); free(sigma); free(ants); free(fnew);
    free(sl); free(slp); free(nsl); free(stmp);
    return rc;
}

/* `calls` evolve() calls of ONE algorithm object on the mt19937 stream (memory != 0: constructed with memory = true) */
int oracle_gaco_evolve_mt(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t n, size_t nx, size_t nix,
                          unsigned gens, unsigned ker, double q, double oracle_par, double acc, unsigned threshold, unsigned n_gen_mark,
                          unsigned impstop, unsigned evalstop, double focus, uint32_t seed, int memory, unsigned calls)
{
    oracle_gaco_state st;
    oracle_gaco_state_init(&st, q, oracle_par);
    st.memory = memory;
    st.archive = (double *)calloc((size_t)ker * (nx + 2), sizeof(double));
    ORACLE_MT_BEGIN(seed);
    int rc = 0;
    for (unsigned c = 0; c < calls && !rc; ++c)
        rc = oracle_gaco_evolve(prob, lb, ub, x, f, n, nx, nix, gens, ker, acc, threshold, n_gen_mark, impstop, evalstop, focus, 0, 0, &st, NULL);
    ORACLE_MT_END();
    free(st.archive);
    return rc;
}
