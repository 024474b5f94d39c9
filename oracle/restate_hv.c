/* oracle/restate_hv.c - plain-C restatement of pagmo's exact hypervolume for 2 and 3 objectives.  TEST INFRASTRUCTURE ONLY.
 * Follows reference
 *   hv2d::compute  src/utils/hv_algos/hv_hv2d.cpp:59-84   (sort by the second objective, sweep with a running width)
 *   hv3d::compute  src/utils/hv_algos/hv_hv3d.cpp:107-166 (sort by the third objective, sweep plane with an ordered front; the
 *                  reference keeps the front in a std::multiset ordered by DEcreasing first objective - here a sorted array)
 *   assert_minimisation src/utils/hv_algos/hv_algorithm.cpp:226-258 (reference point checks)
 * Exclusive contributions (hypervolume::contributions, hypervolume.cpp:286-330; hv2d::contributions hv_hv2d.cpp:133-148 and HyCon3D
 * hv_hv3d.cpp:170-343) are restated through the DEFINITION the reference documents - contribution(p) = HV(S) - HV(S \ {p}) - on top
 * of the restated compute(): O(n) hypervolumes, meant for the small cases of the tests, and accurate to ~1e-16 * HV(S) absolute.
 * Pinned against the reference's fixtures (tests/hypervolume_test_data, through tests/golden/hv_ref.npz) and against the compiled
 * reference (HyCon3D itself) by tests/test_oracle.py.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

int oracle_hv_check(const double *f, size_t n, size_t m, const double *r)
{
    for (size_t i = 0; i < n; ++i) {
        int outside = 0, all_equal = 1;
        for (size_t d = 0; d < m; ++d) {
            outside |= (r[d] < f[i * m + d]);
            all_equal &= (r[d] == f[i * m + d]);
        }
        if (all_equal || outside) return -1;
    }
    return 0;
}

static size_t g_key;
static const double *g_pts;
static size_t g_m;
static int by_key(const void *a, const void *b)
{
    const double x = g_pts[*(const size_t *)a * g_m + g_key], y = g_pts[*(const size_t *)b * g_m + g_key];
    if (x < y) return -1;
    if (x > y) return 1;
    return *(const size_t *)a < *(const size_t *)b ? -1 : 1; /* deterministic ties */
}

static double hv2d(const double *f, const size_t *idx, size_t n, const double *r) /* idx sorted by f[.,1] */
{
    if (n == 0) return 0.0;
#define P(i, d) f[idx[i] * 2 + (d)]
    double hv = 0.0, w = r[0] - P(0, 0);
    for (size_t i = 0; i + 1 < n; ++i) {
        hv += (P(i + 1, 1) - P(i, 1)) * w;
        w = fmax(w, r[0] - P(i + 1, 0));
    }
    hv += (r[1] - P(n - 1, 1)) * w;
#undef P
    return hv;
}

static double hv3d(const double *f, const size_t *idx, size_t n, const double *r) /* idx sorted by f[.,2] */
{
    if (n == 0) return 0.0;
    /* front T ordered by decreasing x; entries hold (x, y); sentinels (r0, -INF) and (-INF, r1) */
    double (*T)[2] = (double(*)[2])malloc((n + 2) * sizeof(*T));
    size_t tn = 0;
    const double INF = DBL_MAX;
    T[tn][0] = r[0]; T[tn][1] = -INF; ++tn;
    T[tn][0] = -INF; T[tn][1] = r[1]; ++tn;
#define X(i) f[idx[i] * 3]
#define Y(i) f[idx[i] * 3 + 1]
#define Z(i) f[idx[i] * 3 + 2]
    double V = 0.0, z3 = Z(0), A = fabs((X(0) - r[0]) * (Y(0) - r[1]));
    memmove(T + 2, T + 1, sizeof(*T)); /* insert the first point between the sentinels */
    T[1][0] = X(0); T[1][1] = Y(0); ++tn;
    for (size_t i = 1; i < n; ++i) {
        /* multiset::insert places the new element after the elements that compare equal (upper bound, decreasing x) */
        size_t pos = 0;
        while (pos < tn && !(X(i) > T[pos][0])) ++pos;
        const double qy = T[pos][1]; /* successor of the new element */
        if (qy <= Y(i)) continue;    /* dominated in the plane: ignored */
        V += A * fabs(z3 - Z(i));
        z3 = Z(i);
        size_t k = pos; /* walk back over the elements the new point dominates: T[k-1], T[k-2], ... */
        while (T[k - 1][1] >= Y(i)) {
            A -= fabs((T[k - 1][0] - T[k - 2][0]) * (T[k - 1][1] - qy));
            --k;
        }
        A += fabs((X(i) - T[k - 1][0]) * (Y(i) - qy));
        /* erase T[k .. pos), put the new point at k */
        memmove(T + k + 1, T + pos, (tn - pos) * sizeof(*T));
        T[k][0] = X(i); T[k][1] = Y(i);
        tn = tn - (pos - k) + 1;
    }
    V += A * fabs(z3 - r[2]);
#undef X
#undef Y
#undef Z
    free(T);
    return V;
}

static double compute_skipping(const double *f, size_t n, size_t m, const double *r, size_t skip)
{
    size_t *idx = (size_t *)malloc((n ? n : 1) * sizeof(size_t)), k = 0;
    for (size_t i = 0; i < n; ++i)
        if (i != skip) idx[k++] = i;
    g_pts = f; g_m = m; g_key = m - 1;
    qsort(idx, k, sizeof(size_t), by_key);
    const double v = (m == 2) ? hv2d(f, idx, k, r) : hv3d(f, idx, k, r);
    free(idx);
    return v;
}

int oracle_hv_compute(const double *f, size_t n, size_t m, const double *r, double *out)
{
    if ((m != 2 && m != 3) || oracle_hv_check(f, n, m, r)) return -1;
    *out = compute_skipping(f, n, m, r, (size_t)-1);
    return 0;
}

int oracle_hv_contributions(const double *f, size_t n, size_t m, const double *r, double *out)
{
    if ((m != 2 && m != 3) || oracle_hv_check(f, n, m, r)) return -1;
    const double all = compute_skipping(f, n, m, r, (size_t)-1);
    for (size_t i = 0; i < n; ++i) out[i] = all - compute_skipping(f, n, m, r, i);
    return 0;
}
