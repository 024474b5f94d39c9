/* oracle/restate_hv.c - plain-C restatement of pagmo's exact hypervolume (hv2d, hv3d, and WFG for 4 and more objectives).  TEST INFRASTRUCTURE ONLY.
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

/* ---- four and more objectives: the WFG algorithm as the reference runs it (src/utils/hv_algos/hv_hvwfg.cpp) ------------------
 * compute_hv (:227-296): one and two points by inclusion-exclusion; otherwise sort the frame by the current last objective
 * (cmp_points, :303-313: larger first), and with that objective dropped, H = sum_p |p[last] - r[last]| * exclusive_hv(p) where
 * exclusive_hv (:212-224) = volume_between(p, r) - hv(limitset) and limitset (:153-209) = the non-dominated subset of
 * { max(p, q) : q after p }.  At two objectives the reference switches to hv2d (:245-248). */
static int wfg_dom_cmp(const double *a, const double *b, size_t d) /* hv_algorithm::dom_cmp, hv_algorithm.cpp:176-209: 1 a dominates, 2 b dominates, 3 equal, 4 incomparable */
{
    for (size_t i = 0; i < d; ++i) {
        if (a[i] > b[i]) {
            for (size_t j = i + 1; j < d; ++j)
                if (a[j] < b[j]) return 4;
            return 2;
        } else if (a[i] < b[i]) {
            for (size_t j = i + 1; j < d; ++j)
                if (a[j] > b[j]) return 4;
            return 1;
        }
    }
    return 3;
}

static double wfg_volume_between(const double *a, const double *b, size_t d)
{
    double v = 1.0;
    for (size_t i = 0; i < d; ++i) v *= (a[i] - b[i]);
    return fabs(v);
}

/* limitset of frame[p] against frame[begin ..] into `out` (row stride m); returns its size */
static size_t wfg_limitset(const double *frame, size_t k, size_t m, size_t d, size_t begin, size_t p, double *out, int *cmp)
{
    size_t no = 0;
    for (size_t idx = begin; idx < k; ++idx) {
        if (idx == p) continue;
        double *s = out + no * m;
        for (size_t c = 0; c < d; ++c) s[c] = fmax(frame[idx * m + c], frame[p * m + c]);
        int keep = 1;
        for (size_t q = 0; q < no; ++q) {
            cmp[q] = wfg_dom_cmp(s, out + q * m, d);
            if (cmp[q] == 2) { keep = 0; break; }
        }
        if (keep) {
            size_t prev = 0;
            for (size_t next = 0; next < no; ++next)
                if (cmp[next] != 1 && cmp[next] != 3) {
                    if (prev < next) memcpy(out + prev * m, out + next * m, d * sizeof(double));
                    ++prev;
                }
            if (prev < no) memmove(out + prev * m, s, d * sizeof(double));
            no = prev + 1;
        }
    }
    return no;
}

static size_t g_wfg_d, g_wfg_m;
static int wfg_by_last_desc(const void *a, const void *b) /* cmp_points :303-313: the larger last objective first, then the earlier ones */
{
    const double *x = (const double *)a, *y = (const double *)b;
    for (size_t i = g_wfg_d; i-- > 0;) {
        if (x[i] > y[i]) return -1;
        if (x[i] < y[i]) return 1;
    }
    return 0;
}

static double wfg_hv(double *frame, size_t k, size_t m, size_t d, const double *r)
{
    if (k == 0) return 0.0;
    if (k == 1) return wfg_volume_between(frame, r, d);
    if (k == 2) {
        double isect = 1.0;
        for (size_t i = 0; i < d; ++i) isect *= (r[i] - fmax(frame[i], frame[m + i]));
        return wfg_volume_between(frame, r, d) + wfg_volume_between(frame + m, r, d) - isect;
    }
    if (d == 2) { /* hv2d on (x, y) rows */
        double *xy = (double *)malloc(k * 2 * sizeof(double));
        size_t *idx = (size_t *)malloc(k * sizeof(size_t));
        for (size_t i = 0; i < k; ++i) { xy[2 * i] = frame[i * m]; xy[2 * i + 1] = frame[i * m + 1]; idx[i] = i; }
        g_pts = xy; g_m = 2; g_key = 1;
        qsort(idx, k, sizeof(size_t), by_key);
        const double v = hv2d(xy, idx, k, r);
        free(xy); free(idx);
        return v;
    }
    g_wfg_d = d; g_wfg_m = m;
    qsort(frame, k, m * sizeof(double), wfg_by_last_desc);
    double *child = (double *)malloc(k * m * sizeof(double));
    int *cmp = (int *)malloc(k * sizeof(int));
    double H = 0.0;
    for (size_t p = 0; p < k; ++p) {
        const size_t no = wfg_limitset(frame, k, m, d - 1, p + 1, p, child, cmp);
        double e = wfg_volume_between(frame + p * m, r, d - 1);
        if (no == 1) e -= wfg_volume_between(child, r, d - 1);
        else if (no > 1) e -= wfg_hv(child, no, m, d - 1, r);
        H += fabs((frame[p * m + d - 1] - r[d - 1]) * e);
    }
    free(child); free(cmp);
    return H;
}

int oracle_hv_compute(const double *f, size_t n, size_t m, const double *r, double *out)
{
    if (m < 2 || oracle_hv_check(f, n, m, r)) return -1;
    if (m >= 4) {
        double *frame = (double *)malloc((n ? n : 1) * m * sizeof(double));
        memcpy(frame, f, n * m * sizeof(double));
        *out = wfg_hv(frame, n, m, m, r);
        free(frame);
        return 0;
    }
    *out = compute_skipping(f, n, m, r, (size_t)-1);
    return 0;
}

int oracle_hv_contributions(const double *f, size_t n, size_t m, const double *r, double *out)
{
    if (m < 2 || oracle_hv_check(f, n, m, r)) return -1;
    if (m >= 4) { /* hvwfg::contributions :92-117: limitset(0, p) in the full dimension, then exclusive_hv */
        double *child = (double *)malloc((n ? n : 1) * m * sizeof(double));
        int *cmp = (int *)malloc((n ? n : 1) * sizeof(int));
        for (size_t p = 0; p < n; ++p) {
            const size_t no = wfg_limitset(f, n, m, m, 0, p, child, cmp);
            double e = wfg_volume_between(f + p * m, r, m);
            if (no == 1) e -= wfg_volume_between(child, r, m);
            else if (no > 1) e -= wfg_hv(child, no, m, m, r);
            out[p] = e;
        }
        free(child); free(cmp);
        return 0;
    }
    const double all = compute_skipping(f, n, m, r, (size_t)-1);
    for (size_t i = 0; i < n; ++i) out[i] = all - compute_skipping(f, n, m, r, i);
    return 0;
}
