/* oracle/restate_pso.c - plain-C restatement of pagmo::pso_gen::evolve (generational PSO).  TEST INFRASTRUCTURE ONLY.
 * Follows reference src/algorithms/pso_gen.cpp: velocity update :231-327 (variants 1-6), clamp/move/box correction :329-363,
 * evaluation :417-440, memory update :445-459, best neighbour :593-623, lbest ring :679-698, gbest :644-664, von Neumann
 * lattice :719-744, adaptive random graph :772-796 (re-drawn after every generation without a new best, :462).
 * PINNED (tests/test_oracle_pin.py): oracle_pso_evolve_mt - these statements on the reference's sequential mt19937 stream -
 * reproduces the compiled pso_gen::evolve bit for bit (variants 1-6, all four topologies).  In the default Philox mode every draw
 * is the value the device consumes at the same (generation, particle, slot) - see oracle/philox.h and pagmo2_b200/csrc/pso.cu.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "philox.h"

static int less_f(double a, double b) { return !isnan(a) && (isnan(b) || a < b); }
static int equal_f(double a, double b) { return (isnan(a) && isnan(b)) || a == b; }
static int leq_f(double a, double b) { return less_f(a, b) || equal_f(a, b); }

/* generic problem evaluation for the restated algorithms */
int oracle_problem_eval(const oracle_problem *p, const double *xs, size_t n, double *fs)
{
    switch (p->family) {
        case 1: case 2: case 3: case 4: case 5: return oracle_simple_batch(p->family, p->dim, xs, n, fs);
        case 6: return oracle_cec2014_batch(p->prob_id, p->dim, p->rotation, p->shift, p->shuffle, xs, n, fs, 1);
        case 7: return oracle_cec2013_batch(p->prob_id, p->dim, p->rotation, p->shift, xs, n, fs);
        case 8: return oracle_zdt_batch(p->prob_id, xs, n, p->dim, fs);
        case 9: return oracle_dtlz_batch(p->prob_id, xs, n, p->dim, p->nobj, p->param, fs);
        case 11: return oracle_lj_batch(p->dim, xs, n, fs);
        default: return -1;
    }
}

int oracle_pso_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x /* in: X, out: lbX */,
                      double *f /* in: fit, out: lbfit */, double *v /* in/out or NULL */, double *xcur /* out or NULL */, size_t n,
                      size_t dim, unsigned gens, double omega, double eta1, double eta2, double max_vel, unsigned variant,
                      unsigned neighb_type, unsigned neighb_param, uint64_t seed, uint32_t first_generation)
{
    if (variant < 1 || variant > 6 || neighb_type < 1 || neighb_type > 4 || n == 0 || neighb_param < 1) return -1;
    double *X = (double *)malloc(n * dim * sizeof(double)), *V = (double *)malloc(n * dim * sizeof(double)),
           *fit = (double *)malloc(n * sizeof(double));
    double *lbX = x, *lbfit = f;
    memcpy(X, x, n * dim * sizeof(double));
    memcpy(fit, f, n * sizeof(double));
    if (v) memcpy(V, v, n * dim * sizeof(double));
    else
        for (size_t p = 0; p < n; ++p)
            for (size_t d = 0; d < dim; ++d) {
                const double vwidth = (ub[d] - lb[d]) * max_vel, minv = -1. * vwidth, maxv = vwidth;
                /* uniform_real_from_range (generic.hpp:98-104): no draw when the range is empty */
                V[p * dim + d] = (minv == maxv) ? minv
                                                : oracle_u01_at(seed, ORACLE_TAG_INIT, first_generation, (uint32_t)p, (uint32_t)d) * (maxv - minv) + minv;
            }
    size_t gbest = 0;
    double gbest_fit = 0;
    if (neighb_type == 1 || neighb_type == 4) { /* pop.best_idx(): first minimum (:205-209) */
        gbest = 0;
        for (size_t p = 1; p < n; ++p)
            if (less_f(lbfit[p], lbfit[gbest])) gbest = p;
        gbest_fit = lbfit[gbest];
    }
    const size_t radius = neighb_param / 2u;
    /* explicit neighbour lists for the lattice and the random graph, as the reference keeps them */
    size_t *nb = NULL, *nb_len = NULL;
    if (neighb_type == 3) { /* initialize_topology__von, :719-744 */
        nb = (size_t *)malloc(n * 4 * sizeof(size_t));
        nb_len = (size_t *)malloc(n * sizeof(size_t));
        static const int diff[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
        int swarm = (int)n, rows = (int)sqrt((double)swarm);
        while (swarm % rows != 0) rows -= 1;
        const int cols = swarm / rows;
        for (int pidx = 0; pidx < swarm; ++pidx) {
            const int p_x = pidx % cols, p_y = pidx / cols;
            for (int k = 0; k < 4; ++k) {
                int n_x = (p_x + diff[k][0]) % cols, n_y = (p_y + diff[k][1]) % rows;
                if (n_x < 0) n_x = cols + n_x;
                if (n_y < 0) n_y = rows + n_y;
                nb[(size_t)pidx * 4 + (size_t)k] = (size_t)(n_y * cols + n_x);
            }
            nb_len[pidx] = 4;
        }
    } else if (neighb_type == 4) { /* n * neighb_param entries in total, whatever the draws */
        nb = (size_t *)malloc(n * neighb_param * sizeof(size_t));
        nb_len = (size_t *)malloc(n * sizeof(size_t));
    }
    size_t *targets = neighb_type == 4 ? (size_t *)malloc(n * neighb_param * sizeof(size_t)) : NULL;
    size_t *nb_off = neighb_type == 4 ? (size_t *)malloc((n + 1) * sizeof(size_t)) : NULL;
/* initialize_topology__adaptive_random, :772-796: particle p informs itself and neighb_param - 1 random particles; the list of q
 * receives its informants in ascending p (p's own entry at p's turn).  Stored as CSR (nb_off / nb) rebuilt from `targets`. */
#define REWIRE(gen_key)                                                                                                \
    do {                                                                                                               \
        for (size_t q = 0; q <= n; ++q) nb_off[q] = 0;                                                                 \
        for (size_t p = 0; p < n; ++p) {                                                                               \
            oracle_stream rs = {seed, ORACLE_TAG_PSO_TOPOLOGY, (gen_key), (uint32_t)p, 1};                             \
            targets[p * neighb_param] = p;                                                                             \
            for (unsigned j = 1; j < neighb_param; ++j) targets[p * neighb_param + j] = oracle_next_below(&rs, n);     \
            for (unsigned j = 0; j < neighb_param; ++j) ++nb_off[targets[p * neighb_param + j] + 1];                   \
        }                                                                                                              \
        for (size_t q = 0; q < n; ++q) nb_off[q + 1] += nb_off[q];                                                     \
        for (size_t q = 0; q < n; ++q) nb_len[q] = 0;                                                                  \
        for (size_t p = 0; p < n; ++p)                                                                                 \
            for (unsigned j = 0; j < neighb_param; ++j) {                                                              \
                const size_t q = targets[p * neighb_param + j];                                                        \
                nb[nb_off[q] + nb_len[q]++] = p;                                                                       \
            }                                                                                                          \
    } while (0)
    if (neighb_type == 4) REWIRE(first_generation);
    int rc = 0;
    for (unsigned g = 0; g < gens && !rc; ++g) {
        const uint32_t generation = first_generation + g;
        for (size_t p = 0; p < n; ++p) {
            if (variant == 6) { /* fully informed particle swarm, :318-326: every neighbour pulls, one draw per (gene, neighbour) */
                const double acceleration_coefficient = eta1 + eta2; /* :220 */
                const size_t K = neighb_type == 1 ? n : (neighb_type == 2 ? 2u * radius : nb_len[p]);
                for (size_t d = 0; d < dim; ++d) {
                    double sum_forces = 0.;
                    for (size_t k = 0; k < K; ++k) {
                        size_t q;
                        if (neighb_type == 1) q = k; /* gbest: neighb[p] = 0 .. n-1 (:656-663) */
                        else if (neighb_type == 2) { /* ring: p - radius .. p - 1, p + 1 .. p + radius (:679-698) */
                            if (k < radius) { const size_t j = radius - k; q = (p < j) ? p - j + n : p - j; }
                            else { const size_t j = k - radius + 1u; q = (p + j >= n) ? p + j - n : p + j; }
                        } else q = (neighb_type == 3 ? nb + p * 4 : nb + nb_off[p])[k];
                        const double u = oracle_u01_at(seed, ORACLE_TAG_PSO, generation, (uint32_t)p, (uint32_t)(d * K + k));
                        sum_forces += u * acceleration_coefficient * (lbX[q * dim + d] - X[p * dim + d]);
                    }
                    V[p * dim + d] = omega * (V[p * dim + d] + sum_forces / (double)K);
                }
                continue;
            }
            size_t b = gbest;
            if (neighb_type >= 3) { /* particle__get_best_neighbor over the explicit list, :608-621: a later entry wins ties */
                const size_t *list = neighb_type == 3 ? nb + p * 4 : nb + nb_off[p];
                b = list[0];
                for (size_t k = 1; k < nb_len[p]; ++k)
                    if (leq_f(lbfit[list[k]], lbfit[b])) b = list[k];
            }
            if (neighb_type == 2) { /* lbest ring + particle__get_best_neighbor */
                int first = 1;
                for (size_t j = radius; j > 0u; --j) {
                    const size_t q = (p < j) ? p - j + n : p - j;
                    if (first || leq_f(lbfit[q], lbfit[b])) b = q;
                    first = 0;
                }
                for (size_t j = 1u; j <= radius; ++j) {
                    const size_t q = (p + j >= n) ? p + j - n : p + j;
                    if (first || leq_f(lbfit[q], lbfit[b])) b = q;
                    first = 0;
                }
            }
            const double *best_neighb = lbX + b * dim;
            double r1 = 0, r2 = 0;
            if (variant == 3 || variant == 4) {
                r1 = oracle_u01_at(seed, ORACLE_TAG_PSO, generation, (uint32_t)p, 0);
                if (variant == 3) r2 = oracle_u01_at(seed, ORACLE_TAG_PSO, generation, (uint32_t)p, 1); /* variant 4 draws r1 only, :274 */
            }
            for (size_t d = 0; d < dim; ++d) {
                double *Vp = &V[p * dim + d];
                const double Xp = X[p * dim + d], lbXp = lbX[p * dim + d];
                if (variant == 1 || variant == 5) {
                    r1 = oracle_u01_at(seed, ORACLE_TAG_PSO, generation, (uint32_t)p, (uint32_t)(2 * d));
                    r2 = oracle_u01_at(seed, ORACLE_TAG_PSO, generation, (uint32_t)p, (uint32_t)(2 * d + 1));
                } else if (variant == 2) {
                    r1 = oracle_u01_at(seed, ORACLE_TAG_PSO, generation, (uint32_t)p, (uint32_t)d);
                }
                switch (variant) {
                    case 1: case 3: *Vp = omega * *Vp + eta1 * r1 * (lbXp - Xp) + eta2 * r2 * (best_neighb[d] - Xp); break;
                    case 2: case 4: *Vp = omega * *Vp + eta1 * r1 * (lbXp - Xp) + eta2 * r1 * (best_neighb[d] - Xp); break;
                    default: *Vp = omega * (*Vp + eta1 * r1 * (lbXp - Xp) + eta2 * r2 * (best_neighb[d] - Xp));
                }
            }
        }
        /* NOTE: the reference finishes ALL velocity updates before moving any particle (two loops); best_neighb only reads lbX,
         * which does not change here, so fusing is equivalent - kept as two loops anyway */
        for (size_t p = 0; p < n; ++p)
            for (size_t d = 0; d < dim; ++d) {
                const double vwidth = (ub[d] - lb[d]) * max_vel, minv = -1. * vwidth, maxv = vwidth;
                double *Vp = &V[p * dim + d];
                if (*Vp > maxv) *Vp = maxv;
                else if (*Vp < minv) *Vp = minv;
                double new_x = X[p * dim + d] + *Vp;
                if (new_x < lb[d]) { new_x = lb[d]; *Vp = 0.; }
                else if (new_x > ub[d]) { new_x = ub[d]; *Vp = 0.; }
                X[p * dim + d] = new_x;
            }
        rc = oracle_problem_eval(prob, X, n, fit);
        int best_fit_improved = 0;
        for (size_t p = 0; p < n && !rc; ++p) {
            if (leq_f(fit[p], lbfit[p])) {
                lbfit[p] = fit[p];
                memcpy(lbX + p * dim, X + p * dim, dim * sizeof(double));
                if ((neighb_type == 1 || neighb_type == 4) && leq_f(fit[p], gbest_fit)) { /* :452-457 */
                    gbest = p;
                    gbest_fit = fit[p];
                    best_fit_improved = 1;
                }
            }
        }
        if (neighb_type == 4 && !best_fit_improved && !rc) REWIRE(generation + 1); /* :462 */
    }
    free(nb); free(nb_len); free(targets); free(nb_off);
    if (v) memcpy(v, V, n * dim * sizeof(double));
    if (xcur) memcpy(xcur, X, n * dim * sizeof(double));
    free(X); free(V); free(fit);
    return rc;
}

/* pso_gen::evolve on the reference's own stream: std::mt19937(seed), velocities drawn inside (memory = false, :193-201) */
int oracle_pso_evolve_mt(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t n, size_t dim,
                         unsigned gens, double omega, double eta1, double eta2, double max_vel, unsigned variant, unsigned neighb_type,
                         unsigned neighb_param, uint32_t seed)
{
    ORACLE_MT_BEGIN(seed);
    const int rc = oracle_pso_evolve(prob, lb, ub, x, f, NULL, NULL, n, dim, gens, omega, eta1, eta2, max_vel, variant, neighb_type,
                                     neighb_param, 0, 0);
    ORACLE_MT_END();
    return rc;
}
