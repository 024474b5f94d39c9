/* restate_maco.c - TEST INFRASTRUCTURE ONLY (see oracle.h): plain-C restatement of pagmo::maco::evolve (multi-objective hypervolume-based
 * ant colony optimisation), reference src/algorithms/maco.cpp:88-533, pheromone_computation :584-678, generate_new_ants :680-750,
 * memory = false.
 *
 * Generational like gaco (restate_gaco.c), and the pheromone values and the ants are gaco's (archive rows are [x | f] here).  The archive
 * is rebuilt every generation from the non-dominated fronts of (archive + population): front by front, each ordered by DEcreasing
 * exclusive hypervolume contribution w.r.t. the front's own nadir + offset (0.1 in the first generation, 0.01 afterwards), and when the
 * first front alone overflows the archive its extreme points are forced into the last rows (:231-259, :379-407 - as written, including
 * the row the reference copies into sol_archive_fit).  Contributions come from oracle_hv_contributions (pinned to hv2d / HyCon3D /
 * hvwfg to ~1e-16 of the hypervolume): the ORDER they induce is what the algorithm consumes.
 * Draw source dispatched as everywhere (philox.h); std::sort tie order through std_sort.h on the mt19937 pin. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "philox.h"
#include "std_sort.h"

static int greater_f(double a, double b) /* detail::greater_than_f */
{
    if (!isnan(a)) return !isnan(b) ? a > b : 0;
    return !isnan(b);
}

typedef struct {
    const double *key;
} key_ctx;
static int key_greater(size_t a, size_t b, const void *c) { return greater_f(((const key_ctx *)c)->key[a], ((const key_ctx *)c)->key[b]); }

void oracle_maco_state_init(oracle_maco_state *s, double q)
{
    s->q = q;
    s->n_evalstop = 0;
    s->gen_mark = 1;
}

/* the archive from the fronts of (fit, dvs) [np rows]: :177-262 (offset 0.1, updates sol_archive_fit) and :318-410 (offset 0.01) */
static int build_archive(const double *dvs, const double *fit, size_t np, size_t nx, size_t m, size_t ker, double offset, int first,
                         double *arch, double *arch_fit)
{
    const size_t row = nx + m;
    size_t *rank = (size_t *)malloc(np * sizeof(size_t)), *domc = (size_t *)malloc(np * sizeof(size_t)),
           *fidx = (size_t *)malloc(np * sizeof(size_t)), *foff = (size_t *)malloc((np + 1) * sizeof(size_t)),
           *sl = (size_t *)malloc(np * sizeof(size_t)), *stmp = (size_t *)malloc(np * sizeof(size_t));
    double *lf = (double *)malloc(np * m * sizeof(double)), *contrib = (double *)malloc(np * sizeof(double)),
           *ref = (double *)malloc(m * sizeof(double)), *idp = (double *)malloc(m * sizeof(double));
    size_t nfronts = 0;
    int rc = oracle_fnds(fit, np, m, rank, domc, fidx, foff, &nfronts);
    size_t i_arch = 0;
    for (size_t fr = 0; fr < nfronts && !rc; ++fr) {
        if (!(i_arch < ker)) break;
        const size_t *idxs = fidx + foff[fr], k = foff[fr + 1] - foff[fr];
        for (size_t i = 0; i < k; ++i) memcpy(lf + i * m, fit + idxs[i] * m, m * sizeof(double));
        for (size_t c = 0; c < m; ++c) { /* hypervolume::refpoint(offset), hypervolume.cpp:160-181 */
            ref[c] = lf[c];
            for (size_t i = 1; i < k; ++i) ref[c] = fmax(ref[c], lf[i * m + c]);
            ref[c] += offset;
        }
        if (k == 1) { /* hypervolume::contributions' trivial case, hypervolume.cpp:292-297 */
            contrib[0] = 1.0;
            for (size_t c = 0; c < m; ++c) contrib[0] *= (lf[c] - ref[c]);
            contrib[0] = fabs(contrib[0]);
        } else if ((rc = oracle_hv_contributions(lf, k, m, ref, contrib))) break;
        for (size_t i = 0; i < k; ++i) sl[i] = i;
        key_ctx kc = {contrib};
        oracle_sort_indices(sl, stmp, k, key_greater, &kc);
        size_t i_hv = 0;
        for (size_t i = 0; i < k && i_arch < ker; ++i) {
            memcpy(arch + i_arch * row, dvs + idxs[sl[i_hv]] * nx, nx * sizeof(double));
            memcpy(arch + i_arch * row + nx, lf + sl[i_hv] * m, m * sizeof(double));
            if (first) memcpy(arch_fit + i_arch * m, lf + sl[i_hv] * m, m * sizeof(double));
            ++i_hv;
            ++i_arch;
        }
        if (i_arch >= ker && fr == 0) { /* the extremities of an overflowing first front, :231-259 */
            for (size_t c = 0; c < m; ++c) {
                idp[c] = lf[c];
                for (size_t i = 1; i < k; ++i) idp[c] = fmin(idp[c], lf[i * m + c]);
            }
            size_t *border = sl; /* reuse: indices into the front's list */
            size_t elem = 0;
            for (size_t c = 0; c < m; ++c)
                for (size_t i = 0; i < k; ++i)
                    if (lf[i * m + c] == idp[c]) {
                        border[elem++] = i;
                        break;
                    }
            for (size_t c = 0; c < m && c < ker; ++c) {
                memcpy(arch + (ker - 1 - c) * row, dvs + idxs[border[c]] * nx, nx * sizeof(double));
                memcpy(arch + (ker - 1 - c) * row + nx, lf + border[c] * m, m * sizeof(double));
                memcpy(arch_fit + (ker - 1 - c) * m, arch + c * row + nx, m * sizeof(double)); /* row c, as the reference writes it */
            }
        }
    }
    free(rank); free(domc); free(fidx); free(foff); free(sl); free(stmp); free(lf); free(contrib); free(ref); free(idp);
    return rc;
}

int oracle_maco_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t n, size_t nx, size_t nix,
                       size_t m, unsigned gens, unsigned ker, unsigned threshold, unsigned n_gen_mark, unsigned evalstop, double focus,
                       uint64_t seed, uint32_t first_generation, oracle_maco_state *st, unsigned *gens_done)
{
    if (gens_done) *gens_done = 0;
    if (n == 0 || m < 2 || ker < 2 || ker > n || focus < 0.) return -1;
    if (gens == 0) return 0;
    if (threshold < 1 || threshold > gens) return -1;
    const size_t row = nx + m, ncx = nx - nix, np = ker + n;
    double *arch = (double *)calloc(ker * row, sizeof(double)), *arch_fit = (double *)calloc(ker * m, sizeof(double)),
           *mdvs = (double *)malloc(np * nx * sizeof(double)), *mfit = (double *)malloc(np * m * sizeof(double)),
           *omega = (double *)malloc(ker * sizeof(double)), *pc = (double *)malloc(ker * sizeof(double)),
           *sigma = (double *)malloc(nx * sizeof(double)), *ants = (double *)malloc(n * nx * sizeof(double)),
           *fnew = (double *)malloc(n * m * sizeof(double)), *id_a = (double *)malloc(m * sizeof(double)),
           *id_b = (double *)malloc(m * sizeof(double));
    for (size_t i = 0; i < ker * row; ++i) arch[i] = 1.0; /* std::vector<vector_double>(m_ker, vector_double(n_x + n_f, 1)), :108 */
    for (size_t i = 0; i < ker * m; ++i) arch_fit[i] = 1.0;
    int rc = 0, stopped = 0;
    unsigned gen;
    for (gen = 1; gen <= gens && !rc; ++gen) {
        const uint32_t generation = first_generation + (gen - 1);
        if (gen == 1) {
            rc = build_archive(x, f, n, nx, m, ker, 0.1, 1, arch, arch_fit);
            if (rc) break;
        } else { /* :264-284 */
            for (size_t j = 0; j < ker; ++j) {
                memcpy(mfit + j * m, arch + j * row + nx, m * sizeof(double));
                memcpy(arch_fit + j * m, arch + j * row + nx, m * sizeof(double));
                memcpy(mdvs + j * nx, arch + j * row, nx * sizeof(double));
            }
            memcpy(mfit + ker * m, f, n * m * sizeof(double));
            memcpy(mdvs + ker * nx, x, n * nx * sizeof(double));
        }
        /* :289-318 */
        for (size_t c = 0; c < m; ++c) {
            id_a[c] = arch_fit[c];
            for (size_t j = 1; j < ker; ++j) id_a[c] = fmin(id_a[c], arch_fit[j * m + c]);
            if (gen == 1) id_b[c] = id_a[c];
            else {
                id_b[c] = mfit[c];
                for (size_t j = 1; j < np; ++j) id_b[c] = fmin(id_b[c], mfit[j * m + c]);
            }
        }
        int check = 0;
        for (size_t c = 0; c < m && !check; ++c)
            if (id_a[c] != id_b[c]) check = 1;
        if (check) ++st->n_evalstop;
        else st->n_evalstop = 0;
        if (st->n_evalstop == 0 || st->n_evalstop > 2) ++st->gen_mark;
        if (st->gen_mark > n_gen_mark) st->gen_mark = 1;
        if (evalstop != 0 && st->n_evalstop >= evalstop) {
            stopped = 1;
            break;
        }
        if (gen > 1) {
            rc = build_archive(mdvs, mfit, np, nx, m, ker, 0.01, 0, arch, arch_fit);
            if (rc) break;
        }
        /* 3 - pheromone_computation, :584-678 (gaco's, on rows [x | f]) */
        if (gen == 1 || gen == threshold) {
            if (gen == threshold) st->q = 0.01;
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
        for (size_t h = 0; h < nx; ++h) {
            double d_min = fabs(arch[h] - arch[row + h]), d_max = d_min;
            for (unsigned c = 0; c + 1 < ker; ++c)
                for (unsigned k = c + 1; k < ker; ++k) {
                    const double d = fabs(arch[c * row + h] - arch[k * row + h]);
                    if (d < d_min) d_min = d;
                    if (d > d_max) d_max = d;
                }
            if (focus != 0. && ((d_max - d_min) / gen > (ub[h] - lb[h]) / focus)) sigma[h] = (ub[h] - lb[h]) / focus;
            else if (h < ncx) sigma[h] = (d_max - d_min) / st->gen_mark;
            else sigma[h] = fmax(fmax((d_max - d_min) / st->gen_mark, 1.0 / st->gen_mark), (1.0 - 1.0 / (sqrt((double)(nx - ncx)))));
        }
        /* 4 - generate_new_ants, :680-750 */
        if (oracle_mt_active) oracle_mt_active->saved_available = 0; /* the normal distribution is passed by value */
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
                double g_h = arch[k_omega * row + h] + sigma[h] * oracle_next_normal(&rs);
                if (g_h < lb[h] || g_h > ub[h]) {
                    int iter_while = 0;
                    while ((g_h < lb[h] || g_h > ub[h]) && iter_while < 10) {
                        g_h = arch[k_omega * row + h] + sigma[h] * oracle_next_normal(&rs);
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
        memcpy(x, ants, n * nx * sizeof(double));
        memcpy(f, fnew, n * m * sizeof(double));
    }
    if (gens_done) *gens_done = gen - 1;
    if (!rc && !stopped) /* :535-545 */
        for (size_t i = 0; i < ker; ++i) {
            memcpy(x + i * nx, arch + i * row, nx * sizeof(double));
            memcpy(f + i * m, arch + i * row + nx, m * sizeof(double));
        }
    free(arch); free(arch_fit); free(mdvs); free(mfit); free(omega); free(pc); free(sigma); free(ants); free(fnew); free(id_a); free(id_b);
    return rc;
}

int oracle_maco_evolve_mt(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t n, size_t nx, size_t nix,
                          size_t m, unsigned gens, unsigned ker, double q, unsigned threshold, unsigned n_gen_mark, unsigned evalstop,
                          double focus, uint32_t seed)
{
    oracle_maco_state st;
    oracle_maco_state_init(&st, q);
    ORACLE_MT_BEGIN(seed);
    const int rc = oracle_maco_evolve(prob, lb, ub, x, f, n, nx, nix, m, gens, ker, threshold, n_gen_mark, evalstop, focus, 0, 0, &st, NULL);
    ORACLE_MT_END();
    return rc;
}
