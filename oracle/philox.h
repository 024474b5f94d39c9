/* oracle/philox.h - Philox4x32-10 and the draw addressing used by the restated generation operators.
 * TEST INFRASTRUCTURE ONLY.  Restated from the published algorithm (Salmon, Moraes, Dror, Shaw, SC'11; Random123
 * known-answer vectors are checked in tests/test_oracle.py); it must produce the same draws as the device-side
 * pagmo2_b200/csrc/philox.cuh for the "parity on injected draws" tests.
 *   counter = {slot / 2, index, generation, tag}, key = {seed lo, seed hi};  u64 = word1:word0 (even slot) or word3:word2 (odd slot);
 *   u01 = (u64 >> 11) * 2^-53
 */
#ifndef ORACLE_PHILOX_H
#define ORACLE_PHILOX_H
#include <stdint.h>

enum { ORACLE_TAG_SHUFFLE1 = 1, ORACLE_TAG_SHUFFLE2 = 2, ORACLE_TAG_NSGA2_VAR = 3, ORACLE_TAG_DE = 4, ORACLE_TAG_PSO = 5,
       ORACLE_TAG_SGA = 6, ORACLE_TAG_INIT = 7, ORACLE_TAG_CMAES = 8, ORACLE_TAG_MIGRATE = 9, ORACLE_TAG_POPULATION = 10 };

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

typedef struct { uint64_t seed; uint32_t tag, generation, index, slot; } oracle_stream;
static inline double oracle_next(oracle_stream *s) { return oracle_philox_u01(s->seed, s->tag, s->generation, s->index, s->slot++); }

#endif
