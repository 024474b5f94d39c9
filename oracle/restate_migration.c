/* oracle/restate_migration.c - plain-C restatement of the migration step of pagmo's archipelago.  TEST INFRASTRUCTURE ONLY.
 * Follows reference
 *   select_best::select   src/s_policies/select_best.cpp:63-171   (all three branches)
 *   fair_replace::replace src/r_policies/fair_replace.cpp:63-221  (all three branches; n_migr :80-108)
 *   ring                  src/topologies/ring.cpp:74-116 + base_bgl_topology::get_connections base_bgl_topology.cpp:214-230
 *   fully_connected       src/topologies/fully_connected.cpp:86-115
 * Flat row-major groups (ids[n], x[n x nx], f[n x nobj]).  The reference orders by std::sort (ties unspecified); this restatement,
 * like the device code, uses a STABLE order: ties keep the original index order (residents before migrants).  Checked against the
 * compiled reference (oracle/_ref) on tie-free inputs by tests/test_oracle.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

static int less_f(double a, double b) /* detail::less_than_f: NaNs are greater than everything, custom_comparisons.hpp:54-74 */
{
    if (!isnan(a)) return !isnan(b) ? a < b : 1;
    return 0;
}

/* constrained single-objective groups (select_best.cpp:137-152, fair_replace.cpp:158-188): rows are [f | nec equality | nic inequality
 * constraints] and the order is sort_population_con's, i.e. compare_fc (src/utils/constrained.cpp:76-118) - restated as written,
 * including its two different norms: the left argument's violation is the SUM of the equality and inequality norms, the right
 * argument's their Euclidean combination.  The two coincide whenever an individual violates constraints of one kind only. */
static inline double max0(double a) { return a < 0. ? 0. : a; } /* std::max(a, 0.): a NaN stays a NaN (never satisfied) */

static _Thread_local struct {
    size_t nec, nic;
    const double *tol;
} g_con = {0, 0, NULL};

static void con_test(const double *f, size_t *nsat, double *leq, double *lineq) /* detail::test_eq_constraints / test_ineq_constraints */
{
    double l2 = 0.;
    size_t n = 0;
    for (size_t j = 0; j < g_con.nec; ++j) {
        const double err = max0(fabs(f[1 + j]) - g_con.tol[j]);
        l2 += err * err;
        if (err <= 0.) ++n;
    }
    *leq = sqrt(l2);
    l2 = 0.;
    for (size_t j = 0; j < g_con.nic; ++j) {
        const double err = max0(f[1 + g_con.nec + j] - g_con.tol[g_con.nec + j]);
        l2 += err * err;
        if (err <= 0.) ++n;
    }
    *lineq = sqrt(l2);
    *nsat = n;
}

static int compare_fc(const double *f1, const double *f2)
{
    size_t n1, n2;
    double e1, i1, e2, i2;
    con_test(f1, &n1, &e1, &i1);
    con_test(f2, &n2, &e2, &i2);
    const double l1 = e1 + i1, l2 = sqrt(e2 * e2 + i2 * i2);
    if (n1 == n2) return n1 == g_con.nec + g_con.nic ? less_f(f1[0], f2[0]) : less_f(l1, l2);
    return n1 > n2;
}

static int row_before(const double *f, size_t nf, size_t a, size_t b)
{
    return (g_con.nec || g_con.nic) ? compare_fc(f + a * nf, f + b * nf) : less_f(f[a * nf], f[b * nf]);
}

/* stable argsort of f[idx][0] (insertion into a merge would be faster; n is a population) */
static void stable_order(const double *f, size_t nobj, size_t n, size_t *idx)
{
    size_t *tmp = (size_t *)malloc(n * sizeof(size_t));
    for (size_t i = 0; i < n; ++i) idx[i] = i;
    for (size_t w = 1; w < n; w *= 2) {
        for (size_t lo = 0; lo < n; lo += 2 * w) {
            size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n, a = lo, b = mid, k = lo;
            while (a < mid && b < hi) tmp[k++] = row_before(f, nobj, idx[b], idx[a]) ? idx[b++] : idx[a++];
            while (a < mid) tmp[k++] = idx[a++];
            while (b < hi) tmp[k++] = idx[b++];
        }
        memcpy(idx, tmp, n * sizeof(size_t));
    }
    free(tmp);
}

/* indices of the best `k` of n individuals: single-objective by fitness, multi-objective by select_best_N_mo */
static int best_indices(const double *f, size_t n, size_t nobj, size_t k, size_t *out)
{
    if (nobj == 1 || g_con.nec || g_con.nic) { /* nobj = the row width: 1 + nec + nic for a constrained group */
        size_t *idx = (size_t *)malloc((n ? n : 1) * sizeof(size_t));
        stable_order(f, nobj, n, idx);
        memcpy(out, idx, k * sizeof(size_t));
        free(idx);
        return 0;
    }
    if (k == 0) return 0;
    size_t nout = 0;
    size_t *tmp = (size_t *)malloc(n * sizeof(size_t));
    int rc = oracle_select_best_N_mo(f, n, nobj, k, tmp, &nout);
    if (!rc && nout != k) rc = -1;
    if (!rc) memcpy(out, tmp, k * sizeof(size_t));
    free(tmp);
    return rc;
}

static long rate_count(int rate_is_frac, double rate, size_t n) /* select_best.cpp:80-100 / fair_replace.cpp:80-103; -1: throws */
{
    if (rate_is_frac) {
        if (!(rate >= 0.0 && rate <= 1.0)) return -1; /* base_sr_policy.cpp:46-55 */
        size_t c = (size_t)(rate * (double)n);
        return (long)(c < n ? c : n);
    }
    if (rate < 0 || (size_t)rate > n) return -1;
    return (long)(size_t)rate;
}

int oracle_select_best(const uint64_t *ids, const double *x, const double *f, size_t n, size_t nx, size_t nobj, int rate_is_frac, double rate,
                       uint64_t *ids_out, double *x_out, double *f_out, size_t *n_out)
{
    const long k = rate_count(rate_is_frac, rate, n);
    if (k < 0) return -1;
    size_t *sel = (size_t *)malloc((k ? k : 1) * sizeof(size_t));
    int rc = best_indices(f, n, nobj, (size_t)k, sel);
    for (long i = 0; i < k && !rc; ++i) {
        ids_out[i] = ids[sel[i]];
        memcpy(x_out + i * nx, x + sel[i] * nx, nx * sizeof(double));
        memcpy(f_out + i * nobj, f + sel[i] * nobj, nobj * sizeof(double));
    }
    *n_out = (size_t)k;
    free(sel);
    return rc;
}

int oracle_fair_replace(const uint64_t *ids, const double *x, const double *f, size_t n, size_t nx, size_t nobj, int rate_is_frac, double rate,
                        const uint64_t *mids, const double *mx, const double *mf, size_t nm, uint64_t *ids_out, double *x_out, double *f_out)
{
    long k = rate_count(rate_is_frac, rate, n);
    if (k < 0) return -1;
    if ((size_t)k > nm) k = (long)nm; /* :105-107 */
    /* the top k migrants, then the merged group residents + migrants */
    size_t *top = (size_t *)malloc((nm ? nm : 1) * sizeof(size_t));
    int rc = best_indices(mf, nm, nobj, (size_t)k, top);
    const size_t tot = n + (size_t)k;
    uint64_t *gid = (uint64_t *)malloc((tot ? tot : 1) * sizeof(uint64_t));
    double *gx = (double *)malloc((tot ? tot : 1) * nx * sizeof(double)), *gf = (double *)malloc((tot ? tot : 1) * nobj * sizeof(double));
    size_t *keep = (size_t *)malloc((tot ? tot : 1) * sizeof(size_t));
    if (!rc) {
        memcpy(gid, ids, n * sizeof(uint64_t));
        memcpy(gx, x, n * nx * sizeof(double));
        memcpy(gf, f, n * nobj * sizeof(double));
        for (long i = 0; i < k; ++i) {
            gid[n + i] = mids[top[i]];
            memcpy(gx + (n + i) * nx, mx + top[i] * nx, nx * sizeof(double));
            memcpy(gf + (n + i) * nobj, mf + top[i] * nobj, nobj * sizeof(double));
        }
        rc = best_indices(gf, tot, nobj, n, keep);
    }
    for (size_t i = 0; i < n && !rc; ++i) {
        ids_out[i] = gid[keep[i]];
        memcpy(x_out + i * nx, gx + keep[i] * nx, nx * sizeof(double));
        memcpy(f_out + i * nobj, gf + keep[i] * nobj, nobj * sizeof(double));
    }
    free(top); free(gid); free(gx); free(gf); free(keep);
    return rc;
}

/* the constrained branches: rows of nf = 1 + nec + nic doubles, tol [nec + nic] */
int oracle_sort_population_con(const double *f, size_t n, size_t nec, size_t nic, const double *tol, size_t *out)
{
    g_con.nec = nec, g_con.nic = nic, g_con.tol = tol;
    stable_order(f, 1 + nec + nic, n, out);
    g_con.nec = g_con.nic = 0;
    return 0;
}

int oracle_select_best_con(const uint64_t *ids, const double *x, const double *f, size_t n, size_t nx, size_t nec, size_t nic, const double *tol,
                           int rate_is_frac, double rate, uint64_t *ids_out, double *x_out, double *f_out, size_t *n_out)
{
    g_con.nec = nec, g_con.nic = nic, g_con.tol = tol;
    const int rc = oracle_select_best(ids, x, f, n, nx, 1 + nec + nic, rate_is_frac, rate, ids_out, x_out, f_out, n_out);
    g_con.nec = g_con.nic = 0;
    return rc;
}

int oracle_fair_replace_con(const uint64_t *ids, const double *x, const double *f, size_t n, size_t nx, size_t nec, size_t nic, const double *tol,
                            int rate_is_frac, double rate, const uint64_t *mids, const double *mx, const double *mf, size_t nm,
                            uint64_t *ids_out, double *x_out, double *f_out)
{
    g_con.nec = nec, g_con.nic = nic, g_con.tol = tol;
    const int rc = oracle_fair_replace(ids, x, f, n, nx, 1 + nec + nic, rate_is_frac, rate, mids, mx, mf, nm, ids_out, x_out, f_out);
    g_con.nec = g_con.nic = 0;
    return rc;
}

/* sources of the edges INTO vertex i, in the order base_bgl_topology::get_connections lists them: the in-edge list of a
 * vecS/bidirectionalS graph keeps insertion order, so the n push_back() calls are replayed (ring.cpp:83-110). */
int oracle_ring_connections(size_t n, size_t i, size_t *out, size_t *count)
{
    if (i >= n) return -1;
    size_t (*in)[4] = (size_t(*)[4])calloc(n ? n : 1, sizeof(*in)); /* in[v] = {count, src...} (at most 2 live + transients) */
#define ADD(u, v) (in[v][1 + in[v][0]++] = (u))
#define DEL(u, v)                                                   \
    do {                                                            \
        size_t c = in[v][0], w = 0;                                 \
        for (size_t q = 0; q < c; ++q)                              \
            if (in[v][1 + q] != (u)) in[v][1 + w++] = in[v][1 + q]; \
        in[v][0] = w;                                               \
    } while (0)
    for (size_t size = 1; size <= n; ++size) {
        if (size == 2) { ADD(0, 1); ADD(1, 0); }
        else if (size == 3) { ADD(1, 2); ADD(2, 1); ADD(2, 0); ADD(0, 2); }
        else if (size > 3) {
            DEL(size - 2, 0); DEL(0, size - 2);
            ADD(size - 2, size - 1); ADD(size - 1, size - 2); ADD(0, size - 1); ADD(size - 1, 0);
        }
    }
    *count = in[i][0];
    for (size_t q = 0; q < in[i][0]; ++q) out[q] = in[i][1 + q];
    free(in);
    return 0;
}

int oracle_fully_connected_connections(size_t n, size_t i, size_t *out, size_t *count)
{
    if (i >= n) return -1;
    size_t k = 0;
    for (size_t j = 0; j < n; ++j)
        if (j != i) out[k++] = j;
    *count = k;
    return 0;
}

/* population(prob, bfe, n, seed): batch_random_decision_vector (include/pagmo/utils/generic.hpp:326-389, continuous part :376-381;
 * libstdc++ uniform_real_distribution = (b - a) * canonical + a) and one random 64-bit ID per individual (population.cpp:155-160),
 * with the device's Philox addressing: gene j of individual i = draw (seed, TAG_POPULATION, 0, i, j), ID = (seed, TAG_POPULATION, 1, i, 0).
 * The Philox form below is what the device is compared with; oracle_population_init_mt restates the same constructor on the
 * reference's sequential mt19937 and is PINNED bit for bit to population(prob, n, seed) (tests/test_oracle_pin.py). */
#include "philox.h"
int oracle_population_init(const double *lb, const double *ub, size_t n, size_t nx, uint64_t seed, double *x, uint64_t *ids)
{
    for (size_t i = 0; i < n; ++i) {
        for (size_t j = 0; j < nx; ++j) {
            if (!isfinite(lb[j]) || !isfinite(ub[j])) return -1;
            x[i * nx + j] = (lb[j] == ub[j]) ? lb[j] : (ub[j] - lb[j]) * oracle_philox_u01(seed, ORACLE_TAG_POPULATION, 0, (uint32_t)i, (uint32_t)j) + lb[j];
        }
        if (ids) ids[i] = oracle_philox_u64(seed, ORACLE_TAG_POPULATION, 1, (uint32_t)i, 0);
    }
    return 0;
}

/* population::population(prob, n, seed) (population.cpp:62-80): n decision vectors drawn first - uniform_real_from_range per
 * gene, generic.hpp:98-104 - then one id per push_back, std::uniform_int_distribution<unsigned long long>() (population.cpp:592),
 * which for a 32-bit engine is two words, high word first */
int oracle_population_init_mt(const double *lb, const double *ub, size_t n, size_t nx, uint32_t seed, double *x, uint64_t *ids)
{
    oracle_mt mt;
    oracle_mt_seed(&mt, seed);
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < nx; ++j) {
            if (!isfinite(lb[j]) || !isfinite(ub[j])) return -1;
            x[i * nx + j] = (lb[j] == ub[j]) ? lb[j] : oracle_mt_real(&mt, lb[j], ub[j]);
        }
    for (size_t i = 0; i < n && ids; ++i) {
        const uint64_t hi = oracle_mt_u32(&mt);
        ids[i] = (hi << 32) + oracle_mt_u32(&mt);
    }
    return 0;
}
