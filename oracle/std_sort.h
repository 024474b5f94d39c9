/* oracle/std_sort.h - GNU libstdc++ (GCC 13) std::sort restated on an index array.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference orders indices with std::sort (multi_objective.cpp:153,304,388,457; sga.cpp:284,349; fair_replace.cpp:123,141;
 * select_best.cpp:117), which is not stable: where keys tie, the result depends on the library's algorithm.  The reference's own
 * tests leave that order open (tests/multi_objective.cpp:239 asserts is_permutation), and the device uses stable sorts.  To compare
 * WHOLE reference runs bit for bit (tests/test_oracle_pin.py) the oracle can switch to this restatement of the published
 * bits/stl_algo.h / bits/stl_heap.h algorithm: introsort (median-of-three to the front, unguarded Hoare partition, depth limit
 * 2*floor(log2 n) then heapsort) down to runs of 16, then one insertion pass.  oracle_sort_libstdcxx selects it (set by the *_mt
 * entry points and oracle_set_sort_mode); the default is the stable merge sort the device is compared with.  The two differ only
 * in the order of elements whose keys compare equal.
 */
#ifndef ORACLE_STD_SORT_H
#define ORACLE_STD_SORT_H
#include <stddef.h>

typedef int (*oracle_before_fn)(size_t a, size_t b, const void *ctx);
extern _Thread_local int oracle_sort_libstdcxx;

/* sorts idx[0..n) like std::sort(idx, idx + n, [&](size_t a, size_t b) { return before(a, b, ctx); }) */
void oracle_std_sort(size_t *idx, size_t n, oracle_before_fn before, const void *ctx);
/* stable merge sort (tmp: n scratch entries) */
void oracle_stable_sort(size_t *idx, size_t *tmp, size_t n, oracle_before_fn before, const void *ctx);
/* the mode switch: libstdc++ order if oracle_sort_libstdcxx, else stable */
void oracle_sort_indices(size_t *idx, size_t *tmp, size_t n, oracle_before_fn before, const void *ctx);

#endif
