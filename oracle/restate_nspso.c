/* restate_nspso.c - TEST INFRASTRUCTURE ONLY (see oracle.h): plain-C restatement of pagmo::nspso::evolve,
 * reference src/algorithms/nspso.cpp:84-411 (compute_maxmin / minfit / compute_niche_count :444-520).
 *
 * nspso is generational in the reference already: every particle of a generation moves on the state the previous generation
 * left (positions, velocities, the archive m_best_dvs / m_best_fit, the leaders chosen at the top of the generation), the new
 * positions are evaluated as ONE batch (the bfe branch, :342-359) and the archive becomes the best N of (new positions + old
 * archive).  Draws, in the reference's order: the initial velocities (particle by particle, coordinate by coordinate), then per
 * generation and particle the leader index - uniform_int(0, ext), repeated while it points at the particle itself - and r1, r2.
 * The draw source is dispatched (philox.h): Philox substreams for the device comparison, std::mt19937 + libstdc++ distributions
 * and libstdc++'s std::sort tie order for the bit-exact pin against the compiled reference (oracle_nspso_evolve_mt). */
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

/* minfit, :444-462 */
static double minfit(size_t i, size_t j, const double *fit, size_t m)
{
    double mn = fit[i * m] - fit[j * m];
    for (size_t k = 0; k < m; ++k) {
        const double t = fit[i * m + k] - fit[j * m + k];
        if (t < mn) mn = t;
    }
    return mn;
}

/* compute_maxmin, :464-484 */
static void compute_maxmin(double *maxmin, const double *fit, size_t n, size_t m)
{
    for (size_t i = 0; i < n; ++i) {
        maxmin[i] = minfit(i, (i + 1) % n, fit, m);
        for (size_t j = 0; j < n; ++j)
            if (i != j) {
                const double t = minfit(i, j, fit, m);
                if (t > maxmin[i]) maxmin[i] = t;
            }
    }
}

int oracle_nspso_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t n, size_t dim, size_t m,
                        unsigned gens, double omega, double c1, double c2, double chi, double v_coeff, unsigned leader_selection_range,
                        unsigned diversity /* 0 crowding distance, 1 niche count, 2 max min */, uint64_t seed, uint32_t first_generation,
                        double *vel /* in/out, or NULL: drawn and dropped */, double *best_x /* in/out, or NULL: starts as x */,
                        double *best_f /* with best_x */)
{
    if (n < 2 || m < 2 || diversity > 2 || leader_selection_range > 100) return -1;
    if (gens == 0) return 0;
    int rc = 0;
    double *V = (double *)malloc(n * dim * sizeof(double)), *bx = (double *)malloc(n * dim * sizeof(double)),
           *bf = (double *)malloc(n * m * sizeof(double)), *minv = (double *)malloc(dim * sizeof(double)),
           *maxv = (double *)malloc(dim * sizeof(double)), *nx = (double *)malloc(2 * n * dim * sizeof(double)),
           *nf = (double *)malloc(2 * n * m * sizeof(double)), *key = (double *)malloc(2 * n * sizeof(double));
    size_t *rank = (size_t *)malloc(2 * n * sizeof(size_t)), *domc = (size_t *)malloc(2 * n * sizeof(size_t)),
           *fidx = (size_t *)malloc(2 * n * sizeof(size_t)), *foff = (size_t *)malloc((2 * n + 1) * sizeof(size_t)),
           *order = (size_t *)malloc(2 * n * sizeof(size_t)), *bnd = (size_t *)malloc(2 * n * sizeof(size_t)),
           *sl = (size_t *)malloc(2 * n * sizeof(size_t)), *tmp = (size_t *)malloc(2 * n * sizeof(size_t));
    memcpy(bx, best_x ? best_x : x, n * dim * sizeof(double));
    memcpy(bf, best_x ? best_f : f, n * m * sizeof(double));
    for (size_t j = 0; j < dim; ++j) { /* :139-143 */
        const double vwidth = (ub[j] - lb[j]) * v_coeff;
        minv[j] = -1. * vwidth;
        maxv[j] = vwidth;
    }
    if (vel) memcpy(V, vel, n * dim * sizeof(double));
    else /* :145-152, uniform_real_from_range (generic.hpp:98-104): no draw when the range is empty */
        for (size_t i = 0; i < n; ++i)
            for (size_t j = 0; j < dim; ++j)
                V[i * dim + j] = (minv[j] == maxv[j])
                                     ? minv[j]
                                     : oracle_u01_at(seed, ORACLE_TAG_INIT, first_generation, (uint32_t)i, (uint32_t)j) * (maxv[j] - minv[j]) + minv[j];

    for (unsigned g = 0; g < gens && !rc; ++g) {
        const uint32_t generation = first_generation + g;
        size_t nfronts = 0, nb = 0;
        if ((rc = oracle_fnds(f, n, m, rank, domc, fidx, foff, &nfronts))) break; /* :160 */
        const size_t n0 = foff[1] - foff[0];
        /* 1 - the leaders, :195-291 */
        if (diversity == 0) {
            if ((rc = oracle_sort_population_mo(f, n, m, order))) break;
            nb = n0 > 1 ? n0 : 2;
            memcpy(bnd, order, nb * sizeof(size_t));
        } else if (diversity == 1) {
            if ((rc = oracle_sort_population_mo(f, n, m, order))) break; /* computed by the reference, unused */
            double ideal[64], nadir[64];
            if (m > 64) { rc = -1; break; }
            for (size_t k = 0; k < m; ++k) { /* ideal: min over all points; nadir: max over the first front (multi_objective.cpp:480-541) */
                size_t bi = 0, wi = 0;
                for (size_t i = 1; i < n; ++i)
                    if (less_f(f[i * m + k], f[bi * m + k])) bi = i;
                for (size_t q = 1; q < n0; ++q)
                    if (less_f(f[fidx[wi] * m + k], f[fidx[q] * m + k])) wi = q;
                ideal[k] = f[bi * m + k];
                nadir[k] = f[fidx[wi] * m + k];
            }
            double delta = 1.0; /* Fonseca-Fleming, :222-252 */
            if (m == 2) {
                const size_t dd = n0 == 1 ? 2 : n0;
                delta = ((nadir[0] - ideal[0]) + (nadir[1] - ideal[1])) / ((double)dd - 1);
            } else if (m == 3) {
                const double d1 = nadir[0] - ideal[0], d2 = nadir[1] - ideal[1], d3 = nadir[2] - ideal[2];
                double ns = (double)n0;
                if (ns < 2.0) ns = 2.0;
                delta = sqrt(4 * d2 * d1 * ns + 4 * d3 * d1 * ns + 4 * d2 * d3 * ns + pow(d1, 2) + pow(d2, 2) + pow(d3, 2) - 2 * d2 * d1 - 2 * d3 * d1
                             - 2 * d2 * d3 + d1 + d2 + d3)
                        / (2 * (ns - 1));
            } else {
                for (size_t k = 0; k < m; ++k) delta *= nadir[k] - ideal[k];
                delta = pow(delta, 1.0 / (double)m) / (double)n0;
            }
            for (size_t a = 0; a < n0; ++a) { /* compute_niche_count, :500-518 (the point itself counts) */
                size_t cnt = 0;
                for (size_t b = 0; b < n0; ++b) {
                    double sum = 0.0;
                    for (size_t j = 0; j < dim; ++j) {
                        const double d = x[fidx[a] * dim + j] - x[fidx[b] * dim + j];
                        sum += d * d;
                    }
                    if (sqrt(sum) < delta) ++cnt;
                }
                key[a] = (double)cnt;
            }
            for (size_t a = 0; a < n0; ++a) sl[a] = a;
            key_ctx kc = {key};
            oracle_sort_indices(sl, tmp, n0, key_less, &kc);
            if (n0 > 1) {
                nb = n0;
                for (size_t a = 0; a < n0; ++a) bnd[a] = fidx[sl[a]];
            } else {
                nb = 2;
                bnd[0] = fidx[foff[0]];
                bnd[1] = fidx[foff[1]];
            }
        } else {
            compute_maxmin(key, f, n, m);
            for (size_t a = 0; a < n; ++a) sl[a] = a;
            key_ctx kc = {key};
            oracle_sort_indices(sl, tmp, n, key_less, &kc);
            size_t i = 1;
            for (; i < n && key[sl[i]] < 0; ++i) {
            }
            if (i < 2) i = 2;
            nb = i;
            memcpy(bnd, sl, nb * sizeof(size_t));
        }
        /* 2 - move, :293-340 */
        int ext = (int)(ceil((double)nb * (double)leader_selection_range / 100.0) - 1);
        if (ext < 1) ext = 1;
        for (size_t idx = 0; idx < n; ++idx) {
            oracle_stream rs = {seed, ORACLE_TAG_NSPSO, generation, (uint32_t)idx, 0};
            size_t leader_idx;
            do {
                leader_idx = oracle_next_below(&rs, (size_t)ext + 1);
            } while (bnd[leader_idx] == idx);
            const double *leader = bx + bnd[leader_idx] * dim;
            const double r1 = oracle_next(&rs);
            const double r2 = oracle_next(&rs);
            for (size_t i = 0; i < dim; ++i) {
                const double xi = x[idx * dim + i];
                double v = omega * V[idx * dim + i] + c1 * r1 * (bx[idx * dim + i] - xi) + c2 * r2 * (leader[i] - xi);
                if (v > maxv[i]) v = maxv[i];
                else if (v < minv[i]) v = minv[i];
                double xn = xi + chi * v;
                if (xn > ub[i]) {
                    xn = ub[i];
                    v = 0.0;
                } else if (xn < lb[i]) {
                    xn = lb[i];
                    v = 0.0;
                }
                V[idx * dim + i] = v;
                nx[idx * dim + i] = xn;
            }
        }
        if ((rc = oracle_problem_eval(prob, nx, n, nf))) break; /* the bfe branch, :342-359 */
        /* 3 - best N of (moved particles | archive), :361-390 */
        memcpy(nx + n * dim, bx, n * dim * sizeof(double));
        memcpy(nf + n * m, bf, n * m * sizeof(double));
        if (diversity != 2) {
            if ((rc = oracle_sort_population_mo(nf, 2 * n, m, order))) break;
        } else {
            compute_maxmin(key, nf, 2 * n, m);
            for (size_t a = 0; a < 2 * n; ++a) order[a] = a;
            key_ctx kc = {key};
            oracle_sort_indices(order, tmp, 2 * n, key_less, &kc);
        }
        for (size_t i = 0; i < n; ++i) {
            memcpy(bx + i * dim, nx + order[i] * dim, dim * sizeof(double));
            memcpy(bf + i * m, nf + order[i] * m, m * sizeof(double));
        }
        /* 4 - the population is the moved swarm, :392-395 */
        memcpy(x, nx, n * dim * sizeof(double));
        memcpy(f, nf, n * m * sizeof(double));
    }
    if (vel) memcpy(vel, V, n * dim * sizeof(double));
    if (best_x) {
        memcpy(best_x, bx, n * dim * sizeof(double));
        memcpy(best_f, bf, n * m * sizeof(double));
    }
    free(V); free(bx); free(bf); free(minv); free(maxv); free(nx); free(nf); free(key);
    free(rank); free(domc); free(fidx); free(foff); free(order); free(bnd); free(sl); free(tmp);
    return rc;
}

/* nspso::evolve on the reference's own stream: std::mt19937(seed), memory = false */
int oracle_nspso_evolve_mt(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t n, size_t dim, size_t m,
                           unsigned gens, double omega, double c1, double c2, double chi, double v_coeff, unsigned leader_selection_range,
                           unsigned diversity, uint32_t seed)
{
    ORACLE_MT_BEGIN(seed);
    const int rc = oracle_nspso_evolve(prob, lb, ub, x, f, n, dim, m, gens, omega, c1, c2, chi, v_coeff, leader_selection_range, diversity, 0, 0,
                                       NULL, NULL, NULL);
    ORACLE_MT_END();
    return rc;
}
