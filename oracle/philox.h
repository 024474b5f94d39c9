/* oracle/philox.h - Philox4x32-10 and the draw addressing used by the restated generation operators.
 * TEST INFRASTRUCTURE ONLY.  Restated from the published algorithm (Salmon, Moraes, Dror, Shaw, SC'11; Random123
 * known-answer vectors are checked in tests/test_oracle.py); it must produce the same draws as the device-side
 * pagmo2_b200/csrc/philox.cuh for the "parity on injected draws" tests.
 *   counter = {slot / 2, index, generation, tag}, key = {seed lo, seed hi};  u64 = word1:word0 (even slot) or word3:word2 (odd slot);
 *   u01 = (u64 >> 11) * 2^-53
 */
#ifndef ORACLE_PHILOX_H
#define ORACLE_PHILOX_H
#include <math.h>
#include <stddef.h>
#include <stdint.h>

enum { ORACLE_TAG_SHUFFLE1 = 1, ORACLE_TAG_SHUFFLE2 = 2, ORACLE_TAG_NSGA2_VAR = 3, ORACLE_TAG_DE = 4, ORACLE_TAG_PSO = 5,
       ORACLE_TAG_SGA = 6, ORACLE_TAG_INIT = 7, ORACLE_TAG_CMAES = 8, ORACLE_TAG_MIGRATE = 9, ORACLE_TAG_POPULATION = 10,
       ORACLE_TAG_PSO_TOPOLOGY = 11, ORACLE_TAG_NSPSO = 12, ORACLE_TAG_MOEAD = 13, ORACLE_TAG_MOEAD_ORDER = 14,
       ORACLE_TAG_MOEAD_INSERT = 15, ORACLE_TAG_GACO = 16 };

static inline void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline uint64_t oracle_philox_u64(uint64_t seed, uint32_t tag, uint32_t generation, uint32_t index, uint32_t slot)
{
    /* one Philox call serves two consecutive slots: words 1:0 for the even slot, 3:2 for the odd one (csrc/philox.cuh) */
    const uint32_t ctr[4] = {slot >> 1, index, generation, tag}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    oracle_philox4x32_10(ctr, key, o);
    return (slot & 1u) ? (((uint64_t)o[3] << 32) | o[2]) : (((uint64_t)o[1] << 32) | o[0]);
}

static inline double oracle_philox_u01(uint64_t seed, uint32_t tag, uint32_t generation, uint32_t index, uint32_t slot)
{
    return (double)(oracle_philox_u64(seed, tag, generation, index, slot) >> 11) * (1.0 / 9007199254740992.0);
}

/* ---- draw source ------------------------------------------------------------------------------------------------------------
 * Two modes.  Philox (default): every draw is addressed by (seed, tag, generation, index, slot) - what the device consumes.
 * Sequential mt19937 (oracle_mt_active != NULL, set by the *_mt entry points): the address is ignored and the draw is the NEXT
 * value of the reference's own generator through the libstdc++ distribution the reference uses at that statement (mt19937.h),
 * so a restated loop written in the reference's statement order reproduces the compiled reference bit for bit.  The arithmetic
 * around the draws is shared by both modes - that is what the pin tests (tests/test_oracle_pin.py) certify. */
#include "mt19937.h"
#include "std_sort.h"
extern _Thread_local oracle_mt *oracle_mt_active;
/* "be the reference": its random stream and its std::sort tie order, for the duration of one *_mt entry point */
#define ORACLE_MT_BEGIN(seed32)          \
    oracle_mt oracle_mt_local_;          \
    oracle_mt_seed(&oracle_mt_local_, (seed32)); \
    oracle_mt_active = &oracle_mt_local_; \
    const int oracle_sort_saved_ = oracle_sort_libstdcxx; \
    oracle_sort_libstdcxx = 1
#define ORACLE_MT_END()       \
    oracle_mt_active = NULL;  \
    oracle_sort_libstdcxx = oracle_sort_saved_

typedef struct { uint64_t seed; uint32_t tag, generation, index, slot; } oracle_stream;
/* uniform_real_distribution<double>(0,1) */
static inline double oracle_u01_at(uint64_t seed, uint32_t tag, uint32_t generation, uint32_t index, uint32_t slot)
{
    return oracle_mt_active ? oracle_mt_u01(oracle_mt_active) : oracle_philox_u01(seed, tag, generation, index, slot);
}
static inline double oracle_next(oracle_stream *s) { return oracle_u01_at(s->seed, s->tag, s->generation, s->index, s->slot++); }
/* uniform_int_distribution(0, n-1): Philox mode = floor(u * n) */
static inline size_t oracle_next_below(oracle_stream *s, size_t n)
{
    if (oracle_mt_active) return (size_t)oracle_mt_below(oracle_mt_active, n);
    const size_t v = (size_t)(oracle_next(s) * (double)n);
    return v < n ? v : n - 1;
}
/* normal_distribution(0,1): Philox mode = Box-Muller on two draws */
static inline double oracle_next_normal(oracle_stream *s)
{
    if (oracle_mt_active) return oracle_mt_normal(oracle_mt_active, 0., 1.);
    const double u1 = 1.0 - oracle_next(s);
    const double u2 = oracle_next(s);
    return sqrt(-2.0 * log(u1)) * cos(2.0 * 3.141592653589793238462643383279502884 * u2);
}

#endif
