/* oracle/restate_mo_utils.c - plain-C restatement of pagmo's multi-objective utilities.
 * TEST INFRASTRUCTURE ONLY: the checker the CUDA path is compared against; never linked into the product.
 * Follows reference src/utils/multi_objective.cpp: pareto_dominance :97-113, fast_non_dominated_sorting :200-257,
 * crowding_distance :280-315, select_best_N_mo :344-396, sort_population_mo :425-465, with the NaN-aware comparisons of
 * include/pagmo/detail/custom_comparisons.hpp:54-88.
 * Sorting: the reference uses std::sort (unstable).  By default this restatement uses a STABLE merge sort (what the device is
 * compared with): bit-identical to the reference whenever the sort keys are distinct, and on the reference's own known answers
 * (tests/multi_objective.cpp:81-209, inputs of <= 16 elements, where libstdc++'s std::sort is an insertion sort).  With
 * oracle_set_sort_mode(1) the sorts follow libstdc++'s introsort (std_sort.h) and the results equal the compiled reference's
 * also where keys tie.  Pinned by tests/test_oracle.py and tests/test_oracle_pin.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "std_sort.h"

static int less_f(double a, double b) /* less_than_f<double, true> */
{
    if (!isnan(a)) return !isnan(b) ? a < b : 1;
    return 0;
}
static int greater_f(double a, double b) /* greater_than_f<double, true> */
{
    if (!isnan(a)) return !isnan(b) ? a > b : 0;
    return !isnan(b) ? 1 : 0;
}

int oracle_pareto_dominance(const double *a, const double *b, size_t m)
{
    int strict = 0;
    for (size_t i = 0; i < m; ++i) {
        if (greater_f(a[i], b[i])) return 0;
        else if (less_f(a[i], b[i])) strict = 1;
    }
    return strict;
}

/* fast_non_dominated_sorting.  Outputs: rank[n], dom_count[n], fronts concatenated in front_idx[n] with
 * front_off[nfronts+1].  The dom_list is kept internally (O(n^2) worst case: small inputs only). */
int oracle_fnds(const double *f, size_t n, size_t m, size_t *rank, size_t *dom_count, size_t *front_idx, size_t *front_off,
                size_t *nfronts)
{
    if (n < 2) return -1;
    if (n > 8192) return oracle_fnds_nolist(f, n, m, rank, dom_count, front_idx, front_off, nfronts); /* same results, O(n) memory */
    size_t *dl_len = (size_t *)calloc(n, sizeof(size_t)), *dl_cap = (size_t *)calloc(n, sizeof(size_t));
    size_t **dl = (size_t **)calloc(n, sizeof(size_t *));
    size_t *cnt = (size_t *)calloc(n, sizeof(size_t));
#define PUSH(a, v)                                                                                                     \
    do {                                                                                                               \
        if (dl_len[a] == dl_cap[a]) {                                                                                  \
            dl_cap[a] = dl_cap[a] ? 2 * dl_cap[a] : 8;                                                                 \
            dl[a] = (size_t *)realloc(dl[a], dl_cap[a] * sizeof(size_t));                                              \
        }                                                                                                              \
        dl[a][dl_len[a]++] = (v);                                                                                      \
    } while (0)
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < i; ++j) {
            if (oracle_pareto_dominance(f + i * m, f + j * m, m)) {
                PUSH(i, j);
                ++cnt[j];
            } else if (oracle_pareto_dominance(f + j * m, f + i * m, m)) {
                PUSH(j, i);
                ++cnt[i];
            }
        }
    size_t nf = 0, filled = 0;
    front_off[0] = 0;
    for (size_t i = 0; i < n; ++i) {
        if (dom_count) dom_count[i] = cnt[i];
        if (cnt[i] == 0) {
            rank[i] = 0;
            front_idx[filled++] = i;
        }
    }
    front_off[1] = filled;
    nf = 1;
    size_t cur_b = 0, cur_e = filled;
    while (cur_e > cur_b) {
        size_t next_b = filled;
        for (size_t p = cur_b; p < cur_e; ++p) {
            const size_t pi = front_idx[p];
            for (size_t q = 0; q < dl_len[pi]; ++q) {
                const size_t qi = dl[pi][q];
                if (--cnt[qi] == 0) {
                    rank[qi] = nf;
                    front_idx[filled++] = qi;
                }
            }
        }
        cur_b = next_b;
        cur_e = filled;
        if (cur_e > cur_b) front_off[++nf] = filled;
    }
    *nfronts = nf;
    for (size_t i = 0; i < n; ++i) free(dl[i]);
    free(dl); free(dl_len); free(dl_cap); free(cnt);
    return 0;
}

/* fast_non_dominated_sorting without the O(n^2)-memory dom_list, for the full-size (65 536 / 131 072 point) parity tests.
 * The reference appends j to dom_list[i] in ascending j (multi_objective.cpp:215-227: entries below i during outer iteration
 * i, entries above i in later outer iterations), and while peeling walks front k in order and, per member p, its dom_list in
 * order (:235-254).  So "for p in front order: for q ascending: if p dominates q: if --count[q] == 0: append q" is the same
 * sequence of appends; the list lookup is replaced by the dominance test itself.  The ascending q scan of one p is split over
 * OpenMP threads in contiguous chunks whose appends are concatenated in chunk order - the order is unchanged.
 * Pinned against oracle_fnds and the compiled reference in tests/test_oracle.py. */
int oracle_fnds_nolist(const double *f, size_t n, size_t m, size_t *rank, size_t *dom_count, size_t *front_idx, size_t *front_off,
                       size_t *nfronts)
{
    if (n < 2) return -1;
    size_t *cnt = (size_t *)calloc(n, sizeof(size_t));
    size_t *alive = (size_t *)malloc(n * sizeof(size_t)); /* ascending indices not yet in a front */
    size_t *freed = (size_t *)malloc(n * sizeof(size_t));
#pragma omp parallel for schedule(dynamic, 64)
    for (size_t j = 0; j < n; ++j) {
        size_t c = 0;
        for (size_t i = 0; i < n; ++i)
            if (i != j && oracle_pareto_dominance(f + i * m, f + j * m, m)) ++c;
        cnt[j] = c;
    }
    size_t nf = 1, filled = 0, nalive = 0;
    front_off[0] = 0;
    for (size_t i = 0; i < n; ++i) {
        if (dom_count) dom_count[i] = cnt[i];
        if (cnt[i] == 0) {
            rank[i] = 0;
            front_idx[filled++] = i;
        } else
            alive[nalive++] = i;
    }
    front_off[1] = filled;
    size_t cur_b = 0, cur_e = filled;
    enum { MAXT = 64 };
    while (cur_e > cur_b && nalive) {
        const size_t next_b = filled;
        for (size_t p = cur_b; p < cur_e; ++p) {
            const double *fp = f + front_idx[p] * m;
            size_t napp[MAXT] = {0};
            int nt = 1;
#pragma omp parallel
            {
                int t = 0, T = 1;
#ifdef _OPENMP
                extern int omp_get_thread_num(void);
                extern int omp_get_num_threads(void);
                t = omp_get_thread_num();
                T = omp_get_num_threads();
                if (T > MAXT) T = MAXT;
#endif
                if (t < T) {
                    if (t == 0) nt = T;
                    const size_t lo = nalive * (size_t)t / (size_t)T, hi = nalive * (size_t)(t + 1) / (size_t)T;
                    size_t k = 0;
                    /* members whose count reaches zero go to freed[lo..lo+k) in order and leave a hole in alive[] */
                    for (size_t a = lo; a < hi; ++a) {
                        const size_t q = alive[a];
                        if (oracle_pareto_dominance(fp, f + q * m, m) && --cnt[q] == 0) {
                            freed[lo + k++] = q;
                            alive[a] = (size_t)-1;
                        }
                    }
                    napp[t] = k;
                }
            }
            /* append the chunks' newly freed points in chunk order, then compact `alive` */
            size_t w = 0;
            for (int t = 0; t < nt; ++t) {
                const size_t lo = nalive * (size_t)t / (size_t)nt, hi = nalive * (size_t)(t + 1) / (size_t)nt;
                for (size_t a = lo; a < lo + napp[t]; ++a) {
                    rank[freed[a]] = nf;
                    front_idx[filled++] = freed[a];
                }
                for (size_t a = lo; a < hi; ++a)
                    if (alive[a] != (size_t)-1) alive[w++] = alive[a];
            }
            nalive = w;
        }
        cur_b = next_b;
        cur_e = filled;
        if (cur_e > cur_b) front_off[++nf] = filled;
    }
    *nfronts = nf;
    free(cnt); free(alive); free(freed);
    return 0;
}

/* index sorts: stable merge sort by default, libstdc++'s std::sort order when oracle_sort_libstdcxx is set (std_sort.h) */
#define msort oracle_sort_indices

struct objkey { const double *f; size_t m, obj; };
static int before_obj(size_t a, size_t b, const void *c)
{
    const struct objkey *k = (const struct objkey *)c;
    return less_f(k->f[a * k->m + k->obj], k->f[b * k->m + k->obj]);
}

int oracle_crowding_distance(const double *f, size_t n, size_t m, double *out)
{
    if (n < 2 || m < 2) return -1;
    size_t *idx = (size_t *)malloc(n * sizeof(size_t)), *tmp = (size_t *)malloc(n * sizeof(size_t));
    for (size_t i = 0; i < n; ++i) { idx[i] = i; out[i] = 0.; }
    for (size_t i = 0; i < m; ++i) {
        struct objkey k = {f, m, i};
        msort(idx, tmp, n, before_obj, &k);
        out[idx[0]] = INFINITY;
        out[idx[n - 1]] = INFINITY;
        const double df = f[idx[n - 1] * m + i] - f[idx[0] * m + i];
        for (size_t j = 1; j + 1 < n; ++j) out[idx[j]] += (f[idx[j + 1] * m + i] - f[idx[j - 1] * m + i]) / df;
    }
    free(idx); free(tmp);
    return 0;
}

struct cdkey { const double *cd; const size_t *rank; };
static int before_cd_desc(size_t a, size_t b, const void *c)
{
    return greater_f(((const struct cdkey *)c)->cd[a], ((const struct cdkey *)c)->cd[b]);
}
static int before_rank_cd(size_t a, size_t b, const void *c)
{
    const struct cdkey *k = (const struct cdkey *)c;
    if (k->rank[a] == k->rank[b]) return greater_f(k->cd[a], k->cd[b]);
    return k->rank[a] < k->rank[b];
}

int oracle_select_best_N_mo(const double *f, size_t n, size_t m, size_t N, size_t *out, size_t *nout)
{
    *nout = 0;
    if (N == 0 || n == 0) return 0;
    if (n == 1) { out[0] = 0; *nout = 1; return 0; }
    if (N >= n) { for (size_t i = 0; i < n; ++i) out[i] = i; *nout = n; return 0; }
    size_t *rank = (size_t *)malloc(n * sizeof(size_t)), *fi = (size_t *)malloc(n * sizeof(size_t)),
           *fo = (size_t *)malloc((n + 1) * sizeof(size_t)), nf = 0, k = 0, front_id = 0;
    if (oracle_fnds(f, n, m, rank, NULL, fi, fo, &nf)) return -1;
    while (front_id < nf && k + (fo[front_id + 1] - fo[front_id]) <= N) {
        for (size_t p = fo[front_id]; p < fo[front_id + 1]; ++p) out[k++] = fi[p];
        ++front_id;
    }
    if (k < N) {
        const size_t b = fo[front_id], sz = fo[front_id + 1] - b;
        double *sub = (double *)malloc(sz * m * sizeof(double)), *cd = (double *)malloc(sz * sizeof(double));
        size_t *idx = (size_t *)malloc(sz * sizeof(size_t)), *tmp = (size_t *)malloc(sz * sizeof(size_t));
        for (size_t i = 0; i < sz; ++i) { memcpy(sub + i * m, f + fi[b + i] * m, m * sizeof(double)); idx[i] = i; }
        if (oracle_crowding_distance(sub, sz, m, cd)) return -1;
        struct cdkey ck = {cd, NULL};
        msort(idx, tmp, sz, before_cd_desc, &ck);
        for (size_t i = 0; k < N; ++i) out[k++] = fi[b + idx[i]];
        free(sub); free(cd); free(idx); free(tmp);
    }
    *nout = N;
    free(rank); free(fi); free(fo);
    return 0;
}

int oracle_sort_population_mo(const double *f, size_t n, size_t m, size_t *out)
{
    if (n == 0) return 0;
    if (n == 1) { out[0] = 0; return 0; }
    size_t *rank = (size_t *)malloc(n * sizeof(size_t)), *fi = (size_t *)malloc(n * sizeof(size_t)),
           *fo = (size_t *)malloc((n + 1) * sizeof(size_t)), *tmp = (size_t *)malloc(n * sizeof(size_t)), nf = 0;
    double *cd = (double *)calloc(n, sizeof(double));
    if (oracle_fnds(f, n, m, rank, NULL, fi, fo, &nf)) return -1;
    for (size_t k = 0; k < nf; ++k) {
        const size_t b = fo[k], sz = fo[k + 1] - b;
        if (sz == 1) { cd[fi[b]] = 0; continue; }
        double *sub = (double *)malloc(sz * m * sizeof(double)), *c2 = (double *)malloc(sz * sizeof(double));
        for (size_t i = 0; i < sz; ++i) memcpy(sub + i * m, f + fi[b + i] * m, m * sizeof(double));
        if (oracle_crowding_distance(sub, sz, m, c2)) return -1;
        for (size_t i = 0; i < sz; ++i) cd[fi[b + i]] = c2[i];
        free(sub); free(c2);
    }
    for (size_t i = 0; i < n; ++i) out[i] = i;
    struct cdkey ck = {cd, rank};
    msort(out, tmp, n, before_rank_cd, &ck);
    free(rank); free(fi); free(fo); free(tmp); free(cd);
    return 0;
}
