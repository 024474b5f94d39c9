// philox.cuh - counter-based random draws (Philox4x32-10, Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
// SC'11) for the device-side generation operators.  The reference draws from a sequential std::mt19937 owned by each
// algorithm (e.g. nsga2.cpp:180-233); a counter-based stream gives every (generation, individual, slot) its own draw, so
// the operators parallelise without changing their per-individual logic, and a CPU restatement consuming the same
// (seed, tag, generation, index, slot) values reproduces the device results ("parity on injected draws").
//   counter = {slot / 2, index, generation, stream tag}, key = {seed lo, seed hi}: one Philox call serves TWO consecutive slots
//   u64     = word1:word0 of the output for an even slot, word3:word2 for the odd slot that follows
//   u01     = (u64 >> 11) * 2^-53 in [0, 1)
#pragma once
#include <cstdint>

namespace pgc
{

enum PhiloxTag : uint32_t { // stream tags: one per consumer so that streams never overlap
    kTagShuffle1 = 1,
    kTagShuffle2 = 2,
    kTagNsga2Var = 3,
    kTagDe = 4,
    kTagPso = 5,
    kTagSga = 6,
    kTagInit = 7,
    kTagCmaes = 8,
    kTagMigrate = 9,
    kTagPopulation = 10,
    kTagPsoTopology = 11, // adaptive-random swarm topology: informant draws (pso_gen.cpp:772-796)
    kTagNspso = 12,       // nspso: leader index (repeated while it is the particle itself), r1, r2 (nspso.cpp:304-316)
    kTagMoead = 13,       // moead_gen: an individual's candidate (diversity draw, parents, crossover, mutation; moead_gen.cpp:227-269)
    kTagMoeadOrder = 14,  // moead_gen: the order of a generation (stands in for std::shuffle, :213)
    kTagMoeadInsert = 15, // moead_gen: the shuffle of a candidate's neighbourhood at insertion (:322-324)
    kTagGaco = 16,        // gaco: an ant's kernel choice and its normal deviates (gaco.cpp:826-868)
    kTagHvApprox = 17     // bf_fpras / bf_approx: the Monte-Carlo samples (hv_bf_fpras.cpp:117-137, hv_bf_approx.cpp:292-320)
};

struct Philox4 {
    uint32_t v[4];
};

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0, p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0, n1 = static_cast<uint32_t>(p1);
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1, n3 = static_cast<uint32_t>(p0);
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{{c0, c1, c2, c3}};
}

__host__ __device__ inline uint64_t philox_u64(uint64_t seed, uint32_t tag, uint32_t generation, uint32_t index, uint32_t slot)
{
    const Philox4 r = philox4x32_10(slot >> 1, index, generation, tag, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    return (slot & 1u) ? ((static_cast<uint64_t>(r.v[3]) << 32) | r.v[2]) : ((static_cast<uint64_t>(r.v[1]) << 32) | r.v[0]);
}

__host__ __device__ inline double philox_u01(uint64_t seed, uint32_t tag, uint32_t generation, uint32_t index, uint32_t slot)
{
    return static_cast<double>(philox_u64(seed, tag, generation, index, slot) >> 11) * (1.0 / 9007199254740992.0);
}

// sequential view of one (tag, generation, index) substream: every second draw comes out of the previous call's upper half
struct PhiloxStream {
    uint64_t seed, spare;
    uint32_t tag, generation, index, slot;
    __host__ __device__ PhiloxStream(uint64_t s, uint32_t t, uint32_t g, uint32_t i) : seed(s), spare(0), tag(t), generation(g), index(i), slot(0) {}
    __host__ __device__ uint64_t next_u64()
    {
        uint64_t out;
        if (slot & 1u) {
            out = spare;
        } else {
            const Philox4 r = philox4x32_10(slot >> 1, index, generation, tag, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
            out = (static_cast<uint64_t>(r.v[1]) << 32) | r.v[0];
            spare = (static_cast<uint64_t>(r.v[3]) << 32) | r.v[2];
        }
        ++slot;
        return out;
    }
    __host__ __device__ double next() { return static_cast<double>(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
};

} // namespace pgc
