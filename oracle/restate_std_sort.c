/* oracle/restate_std_sort.c - see std_sort.h.  TEST INFRASTRUCTURE ONLY.
 * Function by function after libstdc++ 13 bits/stl_algo.h (__sort, __introsort_loop, __unguarded_partition_pivot,
 * __move_median_to_first, __unguarded_partition, __final_insertion_sort, __insertion_sort, __unguarded_linear_insert,
 * __partial_sort/__heap_select) and bits/stl_heap.h (__make_heap, __adjust_heap, __push_heap, __pop_heap, __sort_heap).
 */
#include <string.h>

#include "std_sort.h"

_Thread_local int oracle_sort_libstdcxx = 0;

typedef struct { oracle_before_fn before; const void *ctx; } cmp_t;
#define LESS(a, b) (c->before((a), (b), c->ctx))
enum { THRESHOLD = 16 };

static void swap_(size_t *a, size_t *b) { const size_t t = *a; *a = *b; *b = t; }

static void push_heap_(size_t *first, ptrdiff_t hole, ptrdiff_t top, size_t value, const cmp_t *c)
{
    ptrdiff_t parent = (hole - 1) / 2;
    while (hole > top && LESS(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

static void adjust_heap_(size_t *first, ptrdiff_t hole, ptrdiff_t len, size_t value, const cmp_t *c)
{
    const ptrdiff_t top = hole;
    ptrdiff_t child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (LESS(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap_(first, hole, top, value, c);
}

static void heap_sort_(size_t *first, size_t *last, const cmp_t *c) /* __partial_sort(first, last, last) */
{
    const ptrdiff_t len = last - first;
    if (len >= 2) { /* __make_heap */
        ptrdiff_t parent = (len - 2) / 2;
        for (;;) {
            const size_t value = first[parent];
            adjust_heap_(first, parent, len, value, c);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) { /* __sort_heap: __pop_heap(first, last, last) */
        --last;
        const size_t value = *last;
        *last = *first;
        adjust_heap_(first, 0, last - first, value, c);
    }
}

static void move_median_to_first_(size_t *result, size_t *a, size_t *b, size_t *d, const cmp_t *c)
{
    if (LESS(*a, *b)) {
        if (LESS(*b, *d)) swap_(result, b);
        else if (LESS(*a, *d)) swap_(result, d);
        else swap_(result, a);
    } else if (LESS(*a, *d)) swap_(result, a);
    else if (LESS(*b, *d)) swap_(result, d);
    else swap_(result, b);
}

static size_t *unguarded_partition_(size_t *first, size_t *last, size_t *pivot, const cmp_t *c)
{
    for (;;) {
        while (LESS(*first, *pivot)) ++first;
        --last;
        while (LESS(*pivot, *last)) --last;
        if (!(first < last)) return first;
        swap_(first, last);
        ++first;
    }
}

static void introsort_loop_(size_t *first, size_t *last, long depth_limit, const cmp_t *c)
{
    while (last - first > THRESHOLD) {
        if (depth_limit == 0) {
            heap_sort_(first, last, c);
            return;
        }
        --depth_limit;
        size_t *mid = first + (last - first) / 2;
        move_median_to_first_(first, first + 1, mid, last - 1, c);
        size_t *cut = unguarded_partition_(first + 1, last, first, c);
        introsort_loop_(cut, last, depth_limit, c);
        last = cut;
    }
}

static void unguarded_linear_insert_(size_t *last, const cmp_t *c)
{
    const size_t val = *last;
    size_t *next = last;
    --next;
    while (LESS(val, *next)) {
        *last = *next;
        last = next;
        --next;
    }
    *last = val;
}

static void insertion_sort_(size_t *first, size_t *last, const cmp_t *c)
{
    if (first == last) return;
    for (size_t *i = first + 1; i != last; ++i) {
        if (LESS(*i, *first)) {
            const size_t val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(size_t));
            *first = val;
        } else
            unguarded_linear_insert_(i, c);
    }
}

void oracle_std_sort(size_t *idx, size_t n, oracle_before_fn before, const void *ctx)
{
    if (n == 0) return;
    const cmp_t cc = {before, ctx}, *c = &cc;
    size_t *first = idx, *last = idx + n;
    long lg = 0;
    for (size_t k = n; k > 1; k >>= 1) ++lg; /* std::__lg(n) */
    introsort_loop_(first, last, lg * 2, c);
    if (last - first > THRESHOLD) { /* __final_insertion_sort */
        insertion_sort_(first, first + THRESHOLD, c);
        for (size_t *i = first + THRESHOLD; i != last; ++i) unguarded_linear_insert_(i, c);
    } else
        insertion_sort_(first, last, c);
}

void oracle_stable_sort(size_t *idx, size_t *tmp, size_t n, oracle_before_fn before, const void *ctx)
{
    if (n < 2) return;
    const size_t h = n / 2;
    oracle_stable_sort(idx, tmp, h, before, ctx);
    oracle_stable_sort(idx + h, tmp, n - h, before, ctx);
    size_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = before(idx[j], idx[i], ctx) ? idx[j++] : idx[i++];
    while (i < h) tmp[k++] = idx[i++];
    while (j < n) tmp[k++] = idx[j++];
    memcpy(idx, tmp, n * sizeof(size_t));
}

void oracle_sort_indices(size_t *idx, size_t *tmp, size_t n, oracle_before_fn before, const void *ctx)
{
    if (oracle_sort_libstdcxx) oracle_std_sort(idx, n, before, ctx);
    else oracle_stable_sort(idx, tmp, n, before, ctx);
}

void oracle_set_sort_mode(int libstdcxx) { oracle_sort_libstdcxx = libstdcxx; }

/* std::sort of iota(n) by keys[] ascending (descending if desc) - for the pin test of this file against the real std::sort */
static int before_key_asc(size_t a, size_t b, const void *ctx) { return ((const double *)ctx)[a] < ((const double *)ctx)[b]; }
static int before_key_desc(size_t a, size_t b, const void *ctx) { return ((const double *)ctx)[a] > ((const double *)ctx)[b]; }
int oracle_std_argsort(const double *keys, size_t n, int desc, size_t *out)
{
    for (size_t i = 0; i < n; ++i) out[i] = i;
    oracle_std_sort(out, n, desc ? before_key_desc : before_key_asc, keys);
    return 0;
}
