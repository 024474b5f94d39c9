// mo_utils.cu - multi-objective utilities on sm_100a: fast_non_dominated_sorting, crowding_distance,
// select_best_N_mo, sort_population_mo (reference src/utils/multi_objective.cpp:97-113,200-257,280-315,344-396,425-465).
//
// Integer results (ranks, fronts and the ORDER inside the fronts, dominator counts) are bit-exact with the reference.
// How the reference's front order is reproduced without its O(N^2) dom_list: front 0 is in ascending
// index order; a point q joins front k+1 when its dominator counter reaches zero, i.e. while the reference processes the
// LAST of q's dominators in front k's order, and dom_list[p] is ascending in q.  Hence front k+1 is ordered by
// (position in front k of q's last dominator there, q).  The algorithm is level synchronous:
//   count[q] = #dominators                                   one all-pairs pass  (fnds_count_kernel)
//   repeat: for every unassigned q, c = #dominators in front k and mp = max position among them;
//           count[q] -= c; if it hits 0: rank = k+1, key = mp, q becomes a candidate (fnds_peel_kernel)
//           front k+1 = candidates sorted by (key, q)          (fnds_order_kernel in shared memory, CUB radix sort when big)
// Every (dominator, dominated) pair is tested once in the count pass and once in the peel passes: 2*N^2 dominance tests
// of M FP64 compares in total, no N^2-bit matrix in memory.  The roofline that binds is compare throughput
// (FP64 DSETP issue), not HBM: bytes are 8*N*M in + O(N) out.
// That describes the kernels used for small inputs (n < 4096).  Large inputs run on dense integer ranks, in sorted space
// (SortedView): position cut-offs halve the pair tests, most of the rest shrink to M-1 integer compares, and the level loop is
// ONE resident cooperative kernel (fnds_persistent_kernel: point state in registers, levels separated by a ticket counter and a
// published level word, big levels closed by all blocks together) with a launch-per-level loop as the fallback for inputs
// above one thread per position (profiles/r1x_fnds_variants.txt).
// Crowding distances are exact IEEE (same subtraction / division per element, objectives applied in order); the sort
// inside a front is a stable segmented radix sort, which matches the reference's std::sort whenever the objective values
// inside a front are distinct (the reference's order of ties is unspecified: SURVEY.md F5).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <cub/device/device_segmented_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <limits>
#include <vector>

#include "pgc_internal.cuh"

namespace pgc
{

namespace
{

constexpr unsigned kUnassigned = 0xffffffffu;
constexpr int kTP = 256;          // dominators per shared-memory tile / threads per block in the pair kernels
constexpr int kMaxM = 8;          // objectives held in registers
constexpr int kOrderCap = 4096;   // candidates the single-CTA order kernel sorts in shared memory
constexpr int kBatch = 8;         // levels launched between two host polls
constexpr unsigned kFusedCap = 1024; // largest level the last thread block of a fused peel pass orders itself
#ifndef PGC_PEEL_THREADS
#define PGC_PEEL_THREADS 256
#endif
constexpr int kPT = PGC_PEEL_THREADS; // threads (= list entries) per thread block of the sorted-space peel pass; tiles stay kTP wide

struct Meta {            // device-side bookkeeping of the level loop
    unsigned ncand;      // candidates collected by the current peel
    unsigned front_size; // size of the front being peeled (front `level`)
    unsigned front_off;  // its offset in `order`
    unsigned assigned;   // points placed in fronts so far
    unsigned level;      // index of the front being peeled
    unsigned overflow;   // 1: ncand > kOrderCap, the host must order this level with the big path; 2: a fused pass met a level of
                         //    more than kFusedCap candidates, the host closes it with the 1024-thread order kernel
    unsigned nfronts;
    unsigned stop_after; // 0: peel everything; else stop once this many points sit in closed fronts (select_best_N_mo needs no more)
    unsigned done;
    unsigned nact;       // sorted-space loop: live length of the active list
    unsigned tickets;    // fused levels: thread blocks of the current peel pass that have finished (the last one closes the level)
    unsigned stuck;      // persistent loop: a thread block gave up waiting (watchdog); the host reports an internal error
    // the two words the waiting thread blocks poll, adjacent and 8-byte aligned so that one load reads both
    unsigned level_pub;  // persistent loop: `level`, stored LAST when a level is closed
    unsigned big_pub;    // persistent loop: index of a level that ALL thread blocks close together (see distributed_close)
};
static_assert(offsetof(Meta, level_pub) % 8 == 0 && offsetof(Meta, big_pub) == offsetof(Meta, level_pub) + 4, "polled pair");

// pareto_dominance, multi_objective.cpp:97-113 with the NaN-aware comparisons of detail/custom_comparisons.hpp:54-88
// (NaN is placed after +inf).  a dominates b.
template <int M, bool NANAWARE> __device__ __forceinline__ bool dominates(const double *a, const double *b)
{
    bool strict = false, worse = false;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        const double x = a[i], y = b[i];
        if (NANAWARE) {
            const bool nx = isnan(x), ny = isnan(y);
            worse |= (nx && !ny) || (x > y);  // greater_than_f(x, y)
            strict |= (!nx) && (ny || x < y); // less_than_f(x, y)
        } else {
            worse |= (x > y);
            strict |= (x < y);
        }
    }
    return strict && !worse;
}

__global__ void has_nan_kernel(const double *f, size_t len, unsigned *flag)
{
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    bool any = false;
    for (; i < len; i += static_cast<size_t>(gridDim.x) * blockDim.x) any |= isnan(f[i]);
    if (__syncthreads_or(any) && threadIdx.x == 0) atomicOr(flag, 1u);
}

// count[q] = number of points dominating q (multi_objective.cpp:215-227, both branches of the pair loop)
template <int M, bool NANAWARE>
__global__ void __launch_bounds__(kTP) fnds_count_kernel(const double *__restrict__ f, unsigned n, unsigned *count,
                                                         unsigned *dom_count)
{
    constexpr int m = M;
    __shared__ double tile[kTP * (M ? M : 1)];
    const unsigned q = blockIdx.x * kTP + threadIdx.x;
    double fq[M ? M : 1];
#pragma unroll
    for (int i = 0; i < M; ++i) fq[i] = (q < n) ? f[static_cast<size_t>(q) * m + i] : 0.0;
    unsigned c = 0;
    for (unsigned base = 0; base < n; base += kTP) {
        const unsigned np = min(static_cast<unsigned>(kTP), n - base);
        for (unsigned e = threadIdx.x; e < np * m; e += kTP) tile[e] = f[static_cast<size_t>(base) * m + e];
        __syncthreads();
        if (q < n) {
#pragma unroll 4
            for (unsigned t = 0; t < np; ++t) c += dominates<M, NANAWARE>(tile + t * M, fq) ? 1u : 0u;
        }
        __syncthreads();
    }
    if (q < n) {
        count[q] = c;
        if (dom_count) dom_count[q] = c;
    }
}

// one level: subtract the dominators found in front `level` from every unassigned point's counter
template <int M, bool NANAWARE>
__global__ void __launch_bounds__(kTP) fnds_peel_kernel(const double *__restrict__ f, unsigned n,
                                                        const unsigned *__restrict__ order, unsigned *count, unsigned *rank,
                                                        unsigned *key, unsigned *cand, Meta *meta)
{
    constexpr int m = M;
    __shared__ double tile[kTP * (M ? M : 1)];
    if (meta->overflow || meta->done) return;
    const unsigned fs = meta->front_size, fo = meta->front_off, level = meta->level;
    if (fs == 0) return;
    const unsigned q = blockIdx.x * kTP + threadIdx.x;
    const bool active = q < n && rank[q] == kUnassigned;
    if (!__syncthreads_or(active)) return;
    double fq[M ? M : 1];
#pragma unroll
    for (int i = 0; i < M; ++i) fq[i] = active ? f[static_cast<size_t>(q) * m + i] : 0.0;
    unsigned c = 0, mp = 0;
    for (unsigned base = 0; base < fs; base += kTP) {
        const unsigned np = min(static_cast<unsigned>(kTP), fs - base);
        for (unsigned e = threadIdx.x; e < np * m; e += kTP) {
            const unsigned t = e / m, i = e % m;
            tile[e] = f[static_cast<size_t>(order[fo + base + t]) * m + i];
        }
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (unsigned t = 0; t < np; ++t)
                if (dominates<M, NANAWARE>(tile + t * M, fq)) {
                    ++c;
                    mp = base + t; // positions ascend: the last hit is the maximum
                }
        }
        __syncthreads();
    }
    if (active && c) {
        const unsigned left = count[q] - c;
        count[q] = left;
        if (left == 0) {
            rank[q] = level + 1;
            key[q] = mp;
            cand[atomicAdd(&meta->ncand, 1u)] = q;
        }
    }
}

// ---- the same two kernels on integer RANKS -------------------------------------------------------------------------------
// For large inputs every objective value is first replaced by its dense rank among the n values of that objective (equal
// values share a rank; NaNs, which the reference orders after everything - detail::less_than_f - get the last rank).  Pareto
// dominance only compares values of the same objective, so it is unchanged, but a pair test becomes M 32-bit integer compares
// on 4-byte operands instead of 2*M FP64 compares on 8-byte operands: the FP64 pipe (64 lanes/clk/SM) is no longer the limit.
template <int M> __device__ __forceinline__ bool dominates_rank(const unsigned *a, const unsigned *b)
{
    bool strict = false, worse = false;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        worse |= a[i] > b[i];
        strict |= a[i] < b[i];
    }
    return strict && !worse;
}

// ---- the level loop in SORTED space ----------------------------------------------------------------------------------------
// The points are sorted by the dense rank of their first objective (stable: ties in index order) and the whole loop runs on
// sorted positions; results go back to original indices at the end.  Two things follow:
//   * a point at position p can only be dominated from positions up to the end of its own run of equal first ranks, so
//     both the count pass and every peel pass stop there (half of the pair tests);
//   * for a dominator candidate that lies BEFORE the run that holds the first point of the thread block, the first
//     objective is already known to be strictly smaller: the pair test shrinks to "the other M-1 ranks are <=" - one
//     compare for two objectives - and only the tiles that overlap the block's own runs pay for the full test.
// To make whole tiles classifiable, every front is kept twice: in the reference's order (`order`, the output) and sorted
// by position with each member's index in the front beside it (`pm_pos`, `pm_fpos`; the key of a point that joins the next
// front is the largest such index among its dominators).
struct StillActive { // predicate of the active-list compaction
    const unsigned *count;
    __device__ __forceinline__ bool operator()(const unsigned &pos) const { return count[pos] != 0u; }
};

struct SortedView {
    const unsigned *rs;        // [n x m] dense ranks, rows in sorted order
    const unsigned *src;       // sorted position -> original index
    const unsigned *inv;       // original index -> sorted position
    const unsigned *run_lo;    // per position: start of its run of equal first ranks
    const unsigned *run_end;   //               end (exclusive) of that run
    const unsigned *act;       // positions not yet in a front, ascending: the list the peel passes walk.  Re-compacted every few
    unsigned nact_cap;         //   levels (compact_active below); nact_cap = its length at the last compaction (grid size), the
                               //   live length is meta->nact
    unsigned *count, *rank, *key; // per sorted position
    unsigned *cand;            // candidates of the level under construction (sorted positions) ...
    unsigned long long *cand_key; // ... their ordering keys (key << 32 | original index) and rank rows, written by the thread that
    unsigned *cand_rows;       //     found them, so that closing a level needs no dependent gathers
    unsigned *order;           // fronts in the reference's order, as sorted positions
    unsigned *pm_pos, *pm_fpos; // the same fronts, each sorted by position: position | index in the front
    unsigned *pm_rows;         //   and the members' rank rows in that sequence: a peel tile is one coalesced load
    unsigned *front_off;       // [n + 1] (output)
    Meta *meta;
    unsigned n;
    int m;
};
// A level is a chain of dependent global-memory round trips (~1 us each) around very little arithmetic - with 300-400
// fronts per sort that chain, not the pair tests, was the run time (20 us per peel launch, 10 us per order launch in
// profiles/r1s_fnds_launches.txt).  Hence the redundant arrays above: every phase reads what it needs with independent,
// coalesced loads, and `count != 0` doubles as "not yet in a front".

template <int M> __device__ __forceinline__ bool dominates_tail(const unsigned *a, const unsigned *b) // a[0] < b[0] is known
{
    bool ok = true;
#pragma unroll
    for (int i = 1; i < M; ++i) ok &= a[i] <= b[i];
    return ok;
}

template <int M> __global__ void run_bounds_kernel(const unsigned *__restrict__ rs, unsigned n, unsigned *run_lo, unsigned *run_end)
{
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const unsigned mine = rs[static_cast<size_t>(p) * M];
    unsigned lo = 0, hi = p; // first position whose first rank is >= mine
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        if (rs[static_cast<size_t>(mid) * M] < mine) lo = mid + 1;
        else hi = mid;
    }
    run_lo[p] = lo;
    lo = p + 1, hi = n; // first position with a larger first rank
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        if (rs[static_cast<size_t>(mid) * M] <= mine) lo = mid + 1;
        else hi = mid;
    }
    run_end[p] = lo;
}

// count pass: count[p] = dominators of the point at position p; CTAs issued largest first
template <int M>
__global__ void __launch_bounds__(kTP) fnds_count_sorted_kernel(SortedView V, unsigned *dom_count)
{
    constexpr int m = M;
    __shared__ unsigned tile[kTP * M];
    const unsigned b = gridDim.x - 1 - blockIdx.x;
    const unsigned q = b * kTP + threadIdx.x;
    const unsigned lo = V.run_lo[b * kTP], limit = V.run_end[min(V.n, (b + 1) * kTP) - 1];
    unsigned rq[M];
#pragma unroll
    for (int i = 0; i < M; ++i) rq[i] = (q < V.n) ? V.rs[static_cast<size_t>(q) * m + i] : 0u;
    unsigned c = 0;
    for (unsigned base = 0; base < limit; base += kTP) {
        const unsigned np = min(static_cast<unsigned>(kTP), limit - base);
        for (unsigned e = threadIdx.x; e < np * m; e += kTP) tile[e] = V.rs[static_cast<size_t>(base) * m + e];
        __syncthreads();
        if (q < V.n) {
            if (base + np <= lo) {
#pragma unroll 8
                for (unsigned t = 0; t < np; ++t) c += dominates_tail<M>(tile + t * M, rq) ? 1u : 0u;
            } else {
#pragma unroll 8
                for (unsigned t = 0; t < np; ++t) c += dominates_rank<M>(tile + t * M, rq) ? 1u : 0u;
            }
        }
        __syncthreads();
    }
    if (q < V.n) {
        V.count[q] = c;
        if (dom_count) dom_count[V.src[q]] = c;
    }
}

// ---- two objectives: the count pass in O(N (N / B) log B) instead of O(N^2) ---------------------------------------------------
// In sorted space a point p at a position before the run of q (first rank strictly smaller) dominates q iff its second rank
// is <= q's.  Blocks of kCB positions get their second ranks sorted once (count2_sort_blocks_kernel); the dominators of q in all
// blocks that end before q's run are then one upper_bound per block, and only the positions from the start of the block that
// holds the run's first position to the end of the run are tested pair by pair (at most kCB + the run).
constexpr int kCB = 1024;

__global__ void __launch_bounds__(kCB / 2) count2_sort_blocks_kernel(const unsigned *__restrict__ rs, unsigned n, unsigned *sorted)
{
    __shared__ unsigned s[kCB];
    const unsigned base = blockIdx.x * kCB;
    for (unsigned i = threadIdx.x; i < kCB; i += blockDim.x) s[i] = base + i < n ? rs[static_cast<size_t>(base + i) * 2 + 1] : 0xffffffffu;
    __syncthreads();
    for (unsigned k = 2; k <= kCB; k <<= 1)
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = threadIdx.x; i < kCB; i += blockDim.x) {
                const unsigned l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const unsigned a = s[i], b = s[l];
                    if ((a > b) == up) {
                        s[i] = b;
                        s[l] = a;
                    }
                }
            }
            __syncthreads();
        }
    for (unsigned i = threadIdx.x; i < kCB; i += blockDim.x) sorted[base + i] = s[i];
}

__global__ void __launch_bounds__(256) fnds_count2_kernel(SortedView V, const unsigned *__restrict__ sorted, unsigned *dom_count)
{
    const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= V.n) return;
    const unsigned q0 = V.rs[static_cast<size_t>(q) * 2], q1 = V.rs[static_cast<size_t>(q) * 2 + 1];
    const unsigned lo = V.run_lo[q], end = V.run_end[q];
    const unsigned full_blocks = lo / kCB; // blocks that end at or before the start of q's run: first rank strictly smaller
    unsigned c = 0;
    for (unsigned b = 0; b < full_blocks; ++b) {
        const unsigned *blk = sorted + static_cast<size_t>(b) * kCB;
        unsigned l = 0, h = kCB; // upper_bound(q1): number of second ranks <= q1
        while (l < h) {
            const unsigned mid = (l + h) >> 1;
            if (blk[mid] <= q1) l = mid + 1;
            else h = mid;
        }
        c += l;
    }
    for (unsigned p = full_blocks * kCB; p < end; ++p) { // the rest pair by pair (pareto_dominance on ranks)
        const unsigned p0 = V.rs[static_cast<size_t>(p) * 2], p1 = V.rs[static_cast<size_t>(p) * 2 + 1];
        c += (p0 <= q0 && p1 <= q1 && (p0 < q0 || p1 < q1)) ? 1u : 0u;
    }
    V.count[q] = c;
    if (dom_count) dom_count[V.src[q]] = c;
}

// first level: the points without dominators, key 0 (front 0 is in index order, :228-233)
__global__ void fnds_front0_sorted_kernel(SortedView V)
{
    const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= V.n) return;
    if (V.count[q] == 0) {
        V.rank[q] = 0;
        V.key[q] = 0;
        const unsigned slot = atomicAdd(&V.meta->ncand, 1u);
        V.cand[slot] = q;
        V.cand_key[slot] = V.src[q];
        for (int i = 0; i < V.m; ++i) V.cand_rows[static_cast<size_t>(slot) * V.m + i] = V.rs[static_cast<size_t>(q) * V.m + i];
    } else {
        V.rank[q] = kUnassigned;
    }
}

// one block of kTP entries of the active list against front `level` (fs members at offset fo).  The list is ascending, so
// the block's dominators end at the run of its last entry and tiles before the run of its first entry need the tail test only.
// Entries whose point joined a front since the last compaction are skipped (count == 0); compaction keeps warps densely
// active - without it every warp kept a few live lanes until late and a pass cost N x |front| whatever was left.
// close a level (one CTA): the candidates ordered by (key, original index) become the next front, stored in that order and,
// beside it, sorted by position.  s_sort: kOrderCap entries, s_pos: 1024 entries of shared memory.
__device__ void order_level(const SortedView &V, int first, unsigned long long *s_sort, unsigned *s_pos)
{
    Meta *meta = V.meta;
    // speculative loads for the common small level (thread i owns candidate i): issued before the bookkeeping is known, so the
    // two round trips overlap; entries beyond ncand are stale and ignored (the arrays have n >= 4096 entries)
    const int mm = V.m;
    const unsigned spec_pos = __ldcg(V.cand + threadIdx.x);
    const unsigned long long spec_key = __ldcg(V.cand_key + threadIdx.x);
    unsigned spec_row[kMaxM];
#pragma unroll
    for (int k = 0; k < kMaxM; ++k) spec_row[k] = k < mm ? __ldcg(V.cand_rows + static_cast<size_t>(threadIdx.x) * mm + k) : 0u;
    const unsigned C = __ldcg(&meta->ncand);
    if (C > kOrderCap) {
        if (threadIdx.x == 0) meta->overflow = 1;
        return;
    }
    const unsigned fo = __ldcg(&meta->front_off), fs = __ldcg(&meta->front_size), lvl = __ldcg(&meta->level);
    const unsigned off = first ? 0u : fo + fs;
    if (C <= 1024u) {
        // small levels (the common case): rank sorts - every candidate counts the candidates that precede it in either order
        const int m = V.m;
        const bool spec = blockDim.x >= 1024u; // thread i == candidate i
        for (unsigned i = threadIdx.x; i < C; i += blockDim.x) {
            s_pos[i] = spec ? spec_pos : __ldcg(V.cand + i);
            s_sort[i] = spec ? spec_key : __ldcg(V.cand_key + i);
        }
        // With 1024 threads and C candidates the scan of a candidate is split over `parts` threads (thread t: candidate t % Cr,
        // part t / Cr); the partial counts meet in shared-memory counters kept in the unused upper part of s_sort.
        unsigned *cnt = reinterpret_cast<unsigned *>(s_sort + 1024); // [0, 1024): before | [1024, 2048): bp
        const unsigned Cr = (C + 31u) & ~31u;
        const unsigned parts = (spec && Cr) ? blockDim.x / Cr : 1u;
        if (parts > 1u)
            for (unsigned i = threadIdx.x; i < 2048u; i += blockDim.x) cnt[i] = 0u;
        __syncthreads();
        if (parts > 1u) {
            const unsigned i = threadIdx.x % Cr, part = threadIdx.x / Cr;
            if (i < C && part < parts) {
                const unsigned long long mine = s_sort[i];
                const unsigned myp = s_pos[i];
                const unsigned per = (C + parts - 1u) / parts, j0 = part * per, j1 = min(C, j0 + per);
                unsigned before = 0, bp = 0;
                for (unsigned j = j0; j < j1; ++j) {
                    before += s_sort[j] < mine ? 1u : 0u; // (key, index) pairs are distinct
                    bp += s_pos[j] < myp ? 1u : 0u;
                }
                atomicAdd(cnt + i, before);
                atomicAdd(cnt + 1024 + i, bp);
            }
            __syncthreads();
        }
        for (unsigned i = threadIdx.x; i < C; i += blockDim.x) {
            const unsigned long long mine = s_sort[i];
            const unsigned myp = s_pos[i];
            unsigned row[kMaxM];
            for (int k = 0; k < m; ++k) row[k] = spec ? spec_row[k] : __ldcg(V.cand_rows + static_cast<size_t>(i) * m + k);
            unsigned before = 0, bp = 0;
            if (parts > 1u) {
                before = cnt[i];
                bp = cnt[1024 + i];
            } else {
                for (unsigned j = 0; j < C; ++j) {
                    before += s_sort[j] < mine ? 1u : 0u;
                    bp += s_pos[j] < myp ? 1u : 0u;
                }
            }
            V.order[off + before] = myp;
            V.pm_pos[off + bp] = myp;
            V.pm_fpos[off + bp] = before;
            for (int k = 0; k < m; ++k) V.pm_rows[static_cast<size_t>(off + bp) * m + k] = row[k];
        }
    } else {
        unsigned P = 1;
        while (P < C) P <<= 1;
        auto bitonic = [&]() {
            for (unsigned k = 2; k <= P; k <<= 1)
                for (unsigned j = k >> 1; j > 0; j >>= 1) {
                    for (unsigned i = threadIdx.x; i < P; i += blockDim.x) {
                        const unsigned l = i ^ j;
                        if (l > i) {
                            const bool up = (i & k) == 0;
                            const unsigned long long a = s_sort[i], bb = s_sort[l];
                            if ((a > bb) == up) {
                                s_sort[i] = bb;
                                s_sort[l] = a;
                            }
                        }
                    }
                    __syncthreads();
                }
        };
        for (unsigned i = threadIdx.x; i < P; i += blockDim.x) {
            s_sort[i] = i < C ? __ldcg(V.cand_key + i) : ~0ull;
        }
        __syncthreads();
        bitonic();
        for (unsigned i = threadIdx.x; i < C; i += blockDim.x) {
            const unsigned sp = V.inv[static_cast<unsigned>(s_sort[i] & 0xffffffffu)];
            V.order[off + i] = sp;
            s_sort[i] = (static_cast<unsigned long long>(sp) << 32) | i;
        }
        __syncthreads();
        bitonic(); // the padding (~0) stays at the end
        for (unsigned i = threadIdx.x; i < C; i += blockDim.x) {
            const unsigned pos = static_cast<unsigned>(s_sort[i] >> 32);
            V.pm_pos[off + i] = pos;
            V.pm_fpos[off + i] = static_cast<unsigned>(s_sort[i] & 0xffffffffu);
            for (int k = 0; k < V.m; ++k) V.pm_rows[static_cast<size_t>(off + i) * V.m + k] = V.rs[static_cast<size_t>(pos) * V.m + k];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned level = first ? 0u : lvl + 1;
        if (C) {
            V.front_off[level] = off;
            V.front_off[level + 1] = off + C;
            meta->nfronts = level + 1;
        }
        const unsigned assigned = __ldcg(&meta->assigned) + C;
        meta->level = level;
        meta->front_off = off;
        meta->front_size = C;
        meta->assigned = assigned;
        meta->ncand = 0;
        const unsigned stop = __ldcg(&meta->stop_after);
        if (stop && assigned >= stop) meta->done = 1;
        __threadfence();
        *reinterpret_cast<volatile unsigned *>(&meta->level_pub) = level; // after everything else of the new level is visible
    }
}

// FUSED: the last thread block to finish a pass closes the level itself (order_level), so a level is ONE launch instead of two -
// at 300-400 levels per sort the launches, not the work inside them, were most of a level (ncu: ~4 us active of a ~10 us peel
// launch, ~5 of ~8.5 us for the order launch; profiles/r1x_fnds_variants.txt).
template <int M, bool FUSED> __global__ void __launch_bounds__(kPT) fnds_peel_sorted_kernel(SortedView V)
{
    constexpr int m = M;
    __shared__ unsigned tileR[kTP * M], tileF[kTP];
    __shared__ unsigned long long s_sort[FUSED ? kOrderCap : 1];
    __shared__ unsigned s_pos[FUSED ? 1024 : 1];
    __shared__ bool s_last;
    // A level is latency: every dependent global round trip is ~1 us against ~2 us of arithmetic.  So the loads are issued in
    // two independent chains before anything waits: (list entry -> its count / ranks / run bounds) and (bookkeeping -> tile 0).
    const unsigned nact = V.nact_cap; // == meta->nact: the list only changes in compact_active, after which the host re-reads it
    const unsigned first = blockIdx.x * kPT, idx = first + threadIdx.x;
    const bool in = idx < nact;
    const unsigned q = V.act[in ? idx : first];
    const unsigned q_first = V.act[first], q_last = V.act[min(nact, first + kPT) - 1];
    const Meta *meta = V.meta;
    const unsigned m_overflow = meta->overflow, m_done = meta->done, fs = meta->front_size, fo = meta->front_off, level = meta->level;
    const unsigned left0 = in ? V.count[q] : 0u; // dominators not yet in a closed front; 0 = the point sits in a front
    const unsigned lo = V.run_lo[q_first], limit = V.run_end[q_last];
    const unsigned srcq = V.src[q];
    unsigned rq[M];
#pragma unroll
    for (int i = 0; i < M; ++i) rq[i] = V.rs[static_cast<size_t>(q) * m + i];
    if (m_overflow || m_done || fs == 0) return; // nobody takes a ticket: the level loop is over or waits for the host
    const bool active = left0 != 0;
    const bool work = __syncthreads_or(active);
    unsigned c = 0, mp = 0;
    for (unsigned base = 0; work && base < fs; base += kTP) {
        const unsigned np = min(static_cast<unsigned>(kTP), fs - base);
        const unsigned first_pos = __ldcg(V.pm_pos + fo + base), last_pos = __ldcg(V.pm_pos + fo + base + np - 1);
        for (unsigned e = threadIdx.x; e < np * m; e += kPT) tileR[e] = __ldcg(V.pm_rows + static_cast<size_t>(fo + base) * m + e);
        for (unsigned t = threadIdx.x; t < np; t += kPT) tileF[t] = __ldcg(V.pm_fpos + fo + base + t);
        if (first_pos >= limit) break; // this member and all later ones lie beyond the run of the block's last entry
        __syncthreads();
        if (active) {
            if (last_pos < lo) {
#pragma unroll 4
                for (unsigned t = 0; t < np; ++t)
                    if (dominates_tail<M>(tileR + t * M, rq)) {
                        ++c;
                        mp = max(mp, tileF[t]);
                    }
            } else {
#pragma unroll 4
                for (unsigned t = 0; t < np; ++t)
                    if (dominates_rank<M>(tileR + t * M, rq)) {
                        ++c;
                        mp = max(mp, tileF[t]);
                    }
            }
        }
        __syncthreads();
    }
    if (active && c) {
        const unsigned left = left0 - c;
        V.count[q] = left;
        if (left == 0) {
            V.rank[q] = level + 1;
            V.key[q] = mp;
            const unsigned slot = atomicAdd(&V.meta->ncand, 1u);
            V.cand[slot] = q;
            V.cand_key[slot] = (static_cast<unsigned long long>(mp) << 32) | srcq;
#pragma unroll
            for (int i = 0; i < M; ++i) V.cand_rows[static_cast<size_t>(slot) * m + i] = rq[i];
        }
    }
    if (FUSED) {
        __threadfence(); // this block's candidates are visible before its ticket is
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(&V.meta->tickets, 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            if (threadIdx.x == 0) V.meta->tickets = 0;
            if (__ldcg(&V.meta->ncand) > kFusedCap) {
                if (threadIdx.x == 0) V.meta->overflow = 2; // too big for 256 threads: the host launches the order kernel
            } else {
                order_level(V, 0, s_sort, s_pos);
            }
        }
    }
}

// ---- the level loop as ONE resident kernel -----------------------------------------------------------------------------------
// On this system a dependent kernel launch costs ~10 us end to end whatever it does, and a sort has 300-400 levels.  Here the grid
// stays resident for the whole loop (cooperative launch: one 1024-thread block per SM, all co-resident): thread t of block b owns
// sorted position b * slice + t for the entire sort and keeps its point - ranks, remaining dominator count, original index - in
// REGISTERS, so a level reads nothing on the point side.  A level: every block peels the current front (tiles of its members
// through shared memory, cut at the end of the block's last run), appends its new candidates, fences and takes a ticket; the block
// that takes the last ticket closes the level (order_level, 1024 threads) and publishes meta->level_pub; the others spin on that
// word.  The loop leaves the kernel for the host only when a level has more than kOrderCap candidates (CUB path), writing the
// counts back first so that any path can resume.  A watchdog on the spin (never seen firing) turns a lost wake-up into an error
// instead of a hung device.
constexpr int kPersistThreads = 1024;
constexpr unsigned kSpinLimit = 1u << 26;
constexpr unsigned kBigInKernel = 16384; // largest level the resident grid orders together (rank sort, O(C^2 / threads)); above, CUB

// A level of 1024 < C <= kBigInKernel candidates is closed by ALL blocks of the resident grid: block b ranks its share of the
// candidates (C / gridDim of them) against all C in both orders, the scan of one candidate split over 1024 / share threads, partial
// counts meeting in shared-memory counters.  ~40 us at C = 9 000 against ~150 us for leaving the kernel, two CUB sorts and a
// cooperative relaunch.  Returns with everything stored; the caller fences, takes a ticket and the last block publishes the level.
__device__ void distributed_close(const SortedView &V, unsigned C, unsigned off, unsigned *cnt)
{
    const unsigned per = (C + gridDim.x - 1) / gridDim.x, i0 = min(C, blockIdx.x * per), nloc = min(C, i0 + per) - i0;
    const unsigned Cr = max(32u, (nloc + 31u) & ~31u), parts = blockDim.x / Cr;
    for (unsigned i = threadIdx.x; i < 2048u; i += blockDim.x) cnt[i] = 0u;
    __syncthreads();
    const unsigned li = threadIdx.x % Cr, part = threadIdx.x / Cr;
    if (li < nloc && part < parts) {
        const unsigned long long mine = __ldcg(V.cand_key + i0 + li);
        const unsigned myp = __ldcg(V.cand + i0 + li);
        const unsigned perj = (C + parts - 1) / parts, j0 = part * perj, j1 = min(C, j0 + perj);
        unsigned before = 0, bp = 0;
#pragma unroll 4
        for (unsigned j = j0; j < j1; ++j) {
            before += __ldcg(V.cand_key + j) < mine ? 1u : 0u; // (key, index) pairs are distinct
            bp += __ldcg(V.cand + j) < myp ? 1u : 0u;
        }
        atomicAdd(cnt + li, before);
        atomicAdd(cnt + 1024 + li, bp);
    }
    __syncthreads();
    if (threadIdx.x < nloc) {
        const unsigned before = cnt[threadIdx.x], bp = cnt[1024 + threadIdx.x], myp = __ldcg(V.cand + i0 + threadIdx.x);
        V.order[off + before] = myp;
        V.pm_pos[off + bp] = myp;
        V.pm_fpos[off + bp] = before;
        for (int k = 0; k < V.m; ++k)
            V.pm_rows[static_cast<size_t>(off + bp) * V.m + k] = __ldcg(V.cand_rows + static_cast<size_t>(i0 + threadIdx.x) * V.m + k);
    }
}

template <int M> __global__ void __launch_bounds__(kPersistThreads) fnds_persistent_kernel(SortedView V, unsigned big_inkernel)
{
    constexpr int m = M;
    __shared__ unsigned tileR[kTP * M], tileF[kTP], tileP[kTP];
    __shared__ unsigned long long s_sort[kOrderCap];
    __shared__ unsigned s_pos[1024];
    __shared__ unsigned s_flag[4]; // [0]: code seen by thread 0 (0 peel, 1 leave, 2 close together), [1]: this block took the last
                                   // ticket, [2], [3]: front_size, front_off of the level
    Meta *meta = V.meta;
    // Warp w of block b owns the 32 consecutive positions of chunk b + w * gridDim.x: a point at a high position has more of a
    // front before it than one at a low position, so contiguous slices per block left the last blocks with 4-5x the work of the
    // first (measured with clock64: 21 k against 4.6 k cycles per level) and a level lasts as long as its slowest block.
    // Chunks dealt round-robin give every block the same mix; the cut-offs are per warp.
    const unsigned chunk = blockIdx.x + (threadIdx.x >> 5) * gridDim.x;
    const unsigned q = chunk * 32u + (threadIdx.x & 31u);
    const bool in = q < V.n;
    unsigned left = in ? V.count[q] : 0u;
    const unsigned srcq = in ? V.src[q] : 0u;
    unsigned rq[M];
#pragma unroll
    for (int i = 0; i < M; ++i) rq[i] = in ? V.rs[static_cast<size_t>(q) * m + i] : 0u;
    const bool nonempty = chunk * 32u < V.n;
    const unsigned lo = nonempty ? V.run_lo[chunk * 32u] : 0u, limit = nonempty ? V.run_end[min(V.n, chunk * 32u + 32u) - 1] : 0u;
    unsigned expect = __ldcg(&meta->level); // the front to peel next; nothing changes it before the first publication below
    unsigned did_big = 0xffffffffu;         // last level this block helped to close in distributed_close
#ifdef PGC_FNDS_TIMING
    long long tw = 0, tp = 0, tt = 0, to = 0, nl = 0, norder = 0, t0 = clock64(), t1;
#define PGC_TICK(acc) t1 = clock64(); acc += t1 - t0; t0 = t1;
#else
#define PGC_TICK(acc)
#endif
    for (;;) {
        // ---- wait until front `expect` is published, then read the bookkeeping of this level
        if (threadIdx.x == 0) {
            unsigned spins = 0, code = 0;
            // fault injection (bit 1 of big_inkernel, tests only): block 0 drops out as if its wake-up had been lost; the other
            // blocks then wait for a ticket that never comes until THEIR watchdog fires
            if ((big_inkernel & 2u) && blockIdx.x == 0) meta->stuck = 1;
            for (;;) {
                const unsigned long long both = *reinterpret_cast<volatile unsigned long long *>(&meta->level_pub);
                if (static_cast<unsigned>(both) == expect) break;
                if (did_big != expect && static_cast<unsigned>(both >> 32) == expect) {
                    code = 2; // level `expect` is being closed by all blocks together (this block has not done its share yet)
                    break;
                }
                if (++spins >= kSpinLimit) {
                    meta->stuck = 1;
                    break;
                }
            }
            __threadfence();
            const unsigned m_fs = __ldcg(&meta->front_size), m_fo = __ldcg(&meta->front_off);
            if (__ldcg(&meta->overflow) || __ldcg(&meta->done) || __ldcg(&meta->stuck) || (code == 0 && __ldcg(&meta->assigned) >= V.n)
                || m_fs == 0)
                code = 1;
            s_flag[0] = code;
            s_flag[2] = m_fs; // the level's bookkeeping reaches the other threads through shared memory: one round trip less
            s_flag[3] = m_fo;
        }
        __syncthreads();
        PGC_TICK(tw)
        if (s_flag[0] == 1) break;
        if (s_flag[0] == 2) { // ---- close level `expect` together; meta still describes the front that was just peeled
            const unsigned C = __ldcg(&meta->ncand), off = __ldcg(&meta->front_off) + __ldcg(&meta->front_size);
            distributed_close(V, C, off, reinterpret_cast<unsigned *>(s_sort + 1024));
            did_big = expect;
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) s_flag[1] = atomicAdd(&meta->tickets, 1u) == gridDim.x - 1 ? 1u : 0u;
            __syncthreads();
            if (s_flag[1] && threadIdx.x == 0) { // the last block publishes the level (as the tail of order_level does)
                meta->tickets = 0;
                V.front_off[expect] = off;
                V.front_off[expect + 1] = off + C;
                meta->nfronts = expect + 1;
                const unsigned assigned = __ldcg(&meta->assigned) + C;
                meta->level = expect;
                meta->front_off = off;
                meta->front_size = C;
                meta->assigned = assigned;
                meta->ncand = 0;
                const unsigned stop = __ldcg(&meta->stop_after);
                if (stop && assigned >= stop) meta->done = 1;
                __threadfence();
                *reinterpret_cast<volatile unsigned *>(&meta->level_pub) = expect;
            }
            continue; // back to the wait: level_pub == expect now (or soon)
        }
        const unsigned fs = s_flag[2], fo = s_flag[3], level = expect;
        // ---- peel
        const bool active = left != 0;
        const bool warp_work = __any_sync(0xffffffffu, active);
        const bool work = __syncthreads_or(active);
        unsigned c = 0, mp = 0;
        for (unsigned base = 0; work && base < fs; base += kTP) {
            const unsigned np = min(static_cast<unsigned>(kTP), fs - base);
            for (unsigned e = threadIdx.x; e < np * m; e += kPersistThreads) tileR[e] = __ldcg(V.pm_rows + static_cast<size_t>(fo + base) * m + e);
            for (unsigned t = threadIdx.x; t < np; t += kPersistThreads) {
                tileF[t] = __ldcg(V.pm_fpos + fo + base + t);
                tileP[t] = __ldcg(V.pm_pos + fo + base + t);
            }
            __syncthreads();
            const unsigned first_pos = tileP[0], last_pos = tileP[np - 1]; // one round trip for the whole tile
            if (warp_work && first_pos < limit) {
                // the tile is sorted by position: members [0, ta) lie before the run of the warp's first position (first objective
                // strictly smaller: tail test), [ta, tb) inside the warp's runs (full test), [tb, np) after them (cannot dominate)
                unsigned ta = 0, tb = np;
                if (last_pos >= lo) {
                    unsigned l = 0, h2 = np;
                    while (l < h2) {
                        const unsigned mid = (l + h2) >> 1;
                        if (tileP[mid] < lo) l = mid + 1;
                        else h2 = mid;
                    }
                    ta = l;
                    h2 = np;
                    while (l < h2) {
                        const unsigned mid = (l + h2) >> 1;
                        if (tileP[mid] < limit) l = mid + 1;
                        else h2 = mid;
                    }
                    tb = l;
                } else {
                    ta = np;
                }
                if (active) {
#pragma unroll 8
                    for (unsigned t = 0; t < ta; ++t) { // branch-free: a data-dependent branch per pair costs issue slots
                        const bool hit = dominates_tail<M>(tileR + t * M, rq);
                        c += hit ? 1u : 0u;
                        mp = max(mp, hit ? tileF[t] : 0u);
                    }
#pragma unroll 4
                    for (unsigned t = ta; t < tb; ++t) {
                        const bool hit = dominates_rank<M>(tileR + t * M, rq);
                        c += hit ? 1u : 0u;
                        mp = max(mp, hit ? tileF[t] : 0u);
                    }
                }
            }
            __syncthreads();
        }
        if (active && c) {
            left -= c;
            if (left == 0) {
                V.rank[q] = level + 1;
                V.key[q] = mp;
                const unsigned slot = atomicAdd(&meta->ncand, 1u);
                V.cand[slot] = q;
                V.cand_key[slot] = (static_cast<unsigned long long>(mp) << 32) | srcq;
#pragma unroll
                for (int i = 0; i < M; ++i) V.cand_rows[static_cast<size_t>(slot) * m + i] = rq[i];
            }
        }
        // ---- the block that takes the last ticket closes the level and publishes the next one
        __threadfence();
        __syncthreads();
        PGC_TICK(tp)
        if (threadIdx.x == 0) s_flag[1] = atomicAdd(&meta->tickets, 1u) == gridDim.x - 1 ? 1u : 0u;
        __syncthreads();
        PGC_TICK(tt)
        if (s_flag[1]) {
            if (threadIdx.x == 0) meta->tickets = 0;
            __threadfence();
            const unsigned Cnow = __ldcg(&meta->ncand);
            if ((big_inkernel & 1u) && Cnow > 1024u && Cnow <= kBigInKernel) { // all blocks close this level together
                if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned *>(&meta->big_pub) = expect + 1;
            } else
            order_level(V, 0, s_sort, s_pos); // sets meta->overflow instead when the level is too big for shared memory ...
            if (threadIdx.x == 0 && __ldcg(&meta->overflow)) { // ... and then the waiting blocks are released to leave as well
                __threadfence();
                *reinterpret_cast<volatile unsigned *>(&meta->level_pub) = expect + 1;
            }
#ifdef PGC_FNDS_TIMING
            ++norder;
#endif
            PGC_TICK(to)
        }
        ++expect;
#ifdef PGC_FNDS_TIMING
        ++nl;
#endif
    }
#ifdef PGC_FNDS_TIMING
    if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 70 || blockIdx.x == 147) && nl > 20)
        printf("blk %3d: %lld levels  wait %lld  peel %lld  ticket %lld  order %lld (closed %lld) cycles per level\n", blockIdx.x, nl, tw / nl, tp / nl,
               tt / nl, norder ? to / norder : 0, norder);
#endif
    if (in) V.count[q] = left; // any path can resume from here
}

__global__ void __launch_bounds__(1024) fnds_order_sorted_kernel(SortedView V, int first)
{
    __shared__ unsigned long long s_sort[kOrderCap];
    __shared__ unsigned s_pos[1024];
    const Meta *meta = V.meta;
    if (meta->overflow || meta->done) return;
    if (!first && meta->front_size == 0) return; // finished earlier
    order_level(V, first, s_sort, s_pos);
}

// big levels (more than kOrderCap candidates): the host sorts packed keys with CUB
__global__ void pack_level_keys_kernel(SortedView V, unsigned C, unsigned long long *out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) out[i] = V.cand_key[i];
}
__global__ void unpack_level_order_kernel(SortedView V, const unsigned long long *sorted, unsigned C, unsigned off, unsigned long long *next)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) {
        const unsigned sp = V.inv[static_cast<unsigned>(sorted[i] & 0xffffffffu)];
        V.order[off + i] = sp;
        next[i] = (static_cast<unsigned long long>(sp) << 32) | i;
    }
}
__global__ void unpack_level_pos_kernel(SortedView V, const unsigned long long *sorted, unsigned C, unsigned off)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) {
        const unsigned pos = static_cast<unsigned>(sorted[i] >> 32);
        V.pm_pos[off + i] = pos;
        V.pm_fpos[off + i] = static_cast<unsigned>(sorted[i] & 0xffffffffu);
        for (int k = 0; k < V.m; ++k) V.pm_rows[static_cast<size_t>(off + i) * V.m + k] = V.rs[static_cast<size_t>(pos) * V.m + k];
    }
}

// back to original indices
__global__ void invert_perm_kernel(const unsigned *src, unsigned n, unsigned *inv)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[src[i]] = i;
}
__global__ void fnds_unsort_kernel(SortedView V, unsigned *rank_out, unsigned *key_out, unsigned *order_out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V.n) return;
    const unsigned o = V.src[i];
    rank_out[o] = V.rank[i];
    if (key_out) key_out[o] = V.key[i];
    if (i < V.meta->assigned) order_out[i] = V.src[V.order[i]];
}

// dense ranks of one objective: order-preserving keys (NaN last, -0 == +0), sorted, flag the value changes, scan, scatter
__global__ void objective_keys_kernel(const double *f, unsigned n, int m, int obj, unsigned long long *keys, unsigned *idx)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = f[static_cast<size_t>(i) * m + obj] + 0.0;
    const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
    keys[i] = (v != v) ? ~0ull : ((b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull));
    idx[i] = i;
}
__global__ void key_change_flags_kernel(const unsigned long long *sorted, unsigned n, unsigned *flags)
{
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) flags[j] = (j > 0 && sorted[j] != sorted[j - 1]) ? 1u : 0u;
}
__global__ void scatter_ranks_kernel(const unsigned *sorted_idx, const unsigned *dense, unsigned n, int m, int obj, unsigned *r)
{
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) r[static_cast<size_t>(sorted_idx[j]) * m + obj] = dense[j];
}
__global__ void gather_rows_u32m_kernel(const unsigned *r, const unsigned *src, unsigned n, int m, unsigned *out)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(n) * m) return;
    const unsigned row = static_cast<unsigned>(e / m), c = static_cast<unsigned>(e - static_cast<size_t>(row) * m);
    out[e] = r[static_cast<size_t>(src[row]) * m + c];
}

// first level: the candidates are the points with no dominator, key 0 (front 0 is in index order, :228-233)
__global__ void fnds_front0_kernel(unsigned n, const unsigned *count, unsigned *rank, unsigned *key, unsigned *cand, Meta *meta)
{
    const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    if (count[q] == 0) {
        rank[q] = 0;
        key[q] = 0;
        cand[atomicAdd(&meta->ncand, 1u)] = q;
    } else {
        rank[q] = kUnassigned;
    }
}

// close a level: order the candidates by (key, q) and append them to `order` as the next front
__global__ void __launch_bounds__(1024) fnds_order_kernel(const unsigned *cand, const unsigned *key, unsigned *order,
                                                          unsigned *front_off_out, Meta *meta, int first)
{
    __shared__ unsigned long long s[kOrderCap];
    if (meta->overflow || meta->done) return;
    const unsigned C = meta->ncand;
    if (!first && meta->front_size == 0) return; // finished earlier
    if (C > kOrderCap) {
        if (threadIdx.x == 0) meta->overflow = 1;
        return;
    }
    const unsigned off = first ? 0u : meta->front_off + meta->front_size;
    if (C <= 1024u) {
        // small levels (the common case): rank sort - every candidate counts the candidates that precede it; C*C/1024 compares
        // per thread and two barriers, instead of the ~50 barriers of a bitonic network
        for (unsigned i = threadIdx.x; i < C; i += blockDim.x) s[i] = (static_cast<unsigned long long>(key[cand[i]]) << 32) | cand[i];
        __syncthreads();
        for (unsigned i = threadIdx.x; i < C; i += blockDim.x) {
            const unsigned long long mine = s[i];
            unsigned before = 0;
            for (unsigned j = 0; j < C; ++j) before += s[j] < mine ? 1u : 0u; // (key, index) pairs are distinct
            order[off + before] = static_cast<unsigned>(mine & 0xffffffffu);
        }
    } else {
    unsigned P = 1;
    while (P < C) P <<= 1;
    for (unsigned i = threadIdx.x; i < P; i += blockDim.x)
        s[i] = (i < C) ? ((static_cast<unsigned long long>(key[cand[i]]) << 32) | cand[i]) : ~0ull;
    __syncthreads();
    for (unsigned k = 2; k <= P; k <<= 1)
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = threadIdx.x; i < P; i += blockDim.x) {
                const unsigned l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long a = s[i], b = s[l];
                    if ((a > b) == up) {
                        s[i] = b;
                        s[l] = a;
                    }
                }
            }
            __syncthreads();
        }
    for (unsigned i = threadIdx.x; i < C; i += blockDim.x) order[off + i] = static_cast<unsigned>(s[i] & 0xffffffffu);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned level = first ? 0u : meta->level + 1;
        if (C) {
            front_off_out[level] = off;
            front_off_out[level + 1] = off + C;
            meta->nfronts = level + 1;
        }
        meta->level = level;
        meta->front_off = off;
        meta->front_size = C;
        meta->assigned += C;
        meta->ncand = 0;
        if (meta->stop_after && meta->assigned >= meta->stop_after) meta->done = 1;
    }
}

// big path helpers (the host sorts packed (key, q) with CUB when a level has more than kOrderCap candidates)
__global__ void pack_keys_kernel(const unsigned *cand, const unsigned *key, unsigned C, unsigned long long *out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) out[i] = (static_cast<unsigned long long>(key[cand[i]]) << 32) | cand[i];
}

__global__ void unpack_front_kernel(const unsigned long long *sorted, unsigned C, unsigned off, unsigned *order)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) order[off + i] = static_cast<unsigned>(sorted[i] & 0xffffffffu);
}

__global__ void close_big_level_kernel(unsigned *front_off_out, Meta *meta, int first)
{
    const unsigned C = meta->ncand;
    const unsigned off = first ? 0u : meta->front_off + meta->front_size;
    const unsigned level = first ? 0u : meta->level + 1;
    front_off_out[level] = off;
    front_off_out[level + 1] = off + C;
    meta->nfronts = level + 1;
    meta->level = level;
    meta->level_pub = level;
    meta->front_off = off;
    meta->front_size = C;
    meta->assigned += C;
    meta->ncand = 0;
    meta->overflow = 0;
    if (meta->stop_after && meta->assigned >= meta->stop_after) meta->done = 1;
}

// ---- crowding distance ----------------------------------------------------------------------------------------
__global__ void gather_objective_kernel(const double *f, int m, int obj, const unsigned *order, unsigned n, double *keys,
                                        unsigned *vals)
{
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) {
        const unsigned idx = order ? order[j] : j;
        keys[j] = f[static_cast<size_t>(idx) * m + obj] + 0.0; // -0.0 -> +0.0 so that the radix order is the numeric order
        vals[j] = idx;
    }
}

// multi_objective.cpp:303-313 for every front (segment) at once.  small_rule: 0 = crowding_distance semantics only
// (segments of size < 2 untouched), 1 = nsga2.cpp:188-198 (size 1 or 2 -> inf), 2 = sort_population_mo :441-443 (size 1 -> 0)
__global__ void crowding_accumulate_kernel(const double *keys, const unsigned *vals, const unsigned *seg_of,
                                           const unsigned *front_off, unsigned n, double *cd)
{
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned s = seg_of[j], b = front_off[s], e = front_off[s + 1];
    if (e - b < 2) return;
    const unsigned idx = vals[j];
    if (j == b || j == e - 1) {
        cd[idx] = INFINITY;
    } else {
        const double df = keys[e - 1] - keys[b];
        cd[idx] += (keys[j + 1] - keys[j - 1]) / df;
    }
}

__global__ void segment_ids_kernel(const unsigned *front_off, unsigned nfronts, unsigned *seg_of)
{
    const unsigned s = blockIdx.x;
    if (s >= nfronts) return;
    for (unsigned j = front_off[s] + threadIdx.x; j < front_off[s + 1]; j += blockDim.x) seg_of[j] = s;
}

__global__ void small_front_rule_kernel(const unsigned *order, const unsigned *front_off, unsigned nfronts, int rule, double *cd)
{
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nfronts) return;
    const unsigned b = front_off[s], sz = front_off[s + 1] - b;
    if (rule == 1 && sz <= 2) {
        for (unsigned i = 0; i < sz; ++i) cd[order[b + i]] = INFINITY;
    } else if (rule == 2 && sz == 1) {
        cd[order[b]] = 0.0;
    }
}

__global__ void fill_double_kernel(double *p, size_t n, double v)
{
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void iota_kernel(unsigned *p, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

__global__ void gather_double_kernel(const double *src, const unsigned *idx, unsigned n, double *dst)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

__global__ void gather_u32_kernel(const unsigned *src, const unsigned *idx, unsigned n, unsigned *dst)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// descending sort keys with the reference's NaN placement (greater_than_f: NaN is "greater" than everything, so NaNs come
// FIRST in a descending sort): map to an ascending unsigned key
__global__ void cd_desc_key_kernel(const double *cd, const unsigned *idx, unsigned n, unsigned long long *key)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = cd[idx ? idx[i] : i] + 0.0;
    unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(v));
    unsigned long long asc = (bits & 0x8000000000000000ull) ? ~bits : (bits | 0x8000000000000000ull); // ascending numeric
    if (isnan(v)) asc = ~0ull;                                                                          // NaN largest
    key[i] = ~asc; // descending
}

__global__ void gather_rows_kernel(const double *f, const unsigned *ord, unsigned sz, int m, double *dst)
{
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < sz * m) dst[e] = f[static_cast<size_t>(ord[e / m]) * m + e % m];
}

// Stream-ordered scratch from the device's memory pool (kept warm: pgc_ctx_create sets the release threshold to
// "never"), so the level loop does not pay cudaMalloc / cudaFree (tens of ms each on a busy box) per call.
struct Workspace {
    pgc_ctx *ctx;
    cudaStream_t st;
    std::vector<void *> owned;
    Workspace(pgc_ctx *c, cudaStream_t s) : ctx(c), st(s) {}
    ~Workspace()
    {
        for (void *p : owned) cudaFreeAsync(p, st);
    }
    int alloc_bytes(void **out, size_t bytes)
    {
        void *p = nullptr;
        PGC_CUDA(cudaMallocAsync(&p, bytes ? bytes : 1, st));
        owned.push_back(p);
        *out = p;
        return PGC_OK;
    }
    template <class T> int alloc(T **out, size_t count)
    {
        void *p = nullptr;
        int rc = alloc_bytes(&p, sizeof(T) * std::max<size_t>(count, 1));
        *out = static_cast<T *>(p);
        return rc;
    }
};

inline unsigned blocks_for(size_t n, unsigned t) { return static_cast<unsigned>((n + t - 1) / t); }

} // namespace

// fast_non_dominated_sorting on device-resident f [n x m]; d_rank / d_order [n], d_front_off [n+1] are device outputs.
int fnds_device(pgc_ctx *ctx, const double *d_f, size_t n_, size_t m_, unsigned *d_rank, unsigned *d_dom_count,
                unsigned *d_order, unsigned *d_front_off, unsigned *nfronts_out, cudaStream_t st, unsigned stop_after,
                unsigned *d_key_out)
{
    PGC_REQUIRE(n_ >= 2, "At least two points are needed for fast_non_dominated_sorting: %zu detected.", n_); // :204-207
    PGC_REQUIRE(n_ < 0x7fffffffu, "fast_non_dominated_sorting: too many points (%zu)", n_);
    PGC_REQUIRE(m_ <= static_cast<size_t>(kMaxM), "fast_non_dominated_sorting: at most %d objectives are supported on the device, got %zu", kMaxM, m_);
    const unsigned n = static_cast<unsigned>(n_);
    const int m = static_cast<int>(m_);
    Workspace ws(ctx, st);
    unsigned *count, *key, *cand, *flag;
    Meta *meta;
    unsigned long long *pk0 = nullptr, *pk1 = nullptr;
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0;
    int rc;
    if ((rc = ws.alloc(&count, n)) || (rc = ws.alloc(&cand, n)) || (rc = ws.alloc(&flag, 1)) || (rc = ws.alloc(&meta, 1))) return rc;
    key = d_key_out; // the ordering keys (position of the last dominator in the previous front) stay with the caller if it asks
    if (!key && (rc = ws.alloc(&key, n))) return rc;
    {
        Meta m0{};
        m0.stop_after = stop_after;
        m0.big_pub = 0xffffffffu;
        PGC_CUDA(cudaMemcpyAsync(meta, &m0, sizeof(Meta), cudaMemcpyHostToDevice, st));
        PGC_CUDA(cudaStreamSynchronize(st)); // m0 is a stack object
    }
    PGC_CUDA(cudaMemsetAsync(flag, 0, sizeof(unsigned), st));
    bool nanaware = false;
    if (m > 0) {
        has_nan_kernel<<<std::min(1024u, blocks_for(n_ * m_, 256)), 256, 0, st>>>(d_f, n_ * m_, flag);
        unsigned h = 0;
        PGC_CUDA(cudaMemcpyAsync(&h, flag, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        nanaware = h != 0;
    }
    const unsigned gb = blocks_for(n, kTP);
    using count_fn = void (*)(const double *, unsigned, unsigned *, unsigned *);
    using peel_fn = void (*)(const double *, unsigned, const unsigned *, unsigned *, unsigned *, unsigned *, unsigned *, Meta *);
    count_fn count_k = nullptr;
    peel_fn peel_k = nullptr;
#define PGC_MO_CASE(MM)                                                                                                \
    case MM:                                                                                                           \
        count_k = nanaware ? fnds_count_kernel<MM, true> : fnds_count_kernel<MM, false>;                               \
        peel_k = nanaware ? fnds_peel_kernel<MM, true> : fnds_peel_kernel<MM, false>;                                  \
        break;
    switch (m) {
        PGC_MO_CASE(0) PGC_MO_CASE(1) PGC_MO_CASE(2) PGC_MO_CASE(3) PGC_MO_CASE(4) PGC_MO_CASE(5) PGC_MO_CASE(6) PGC_MO_CASE(7)
        PGC_MO_CASE(8)
    }
#undef PGC_MO_CASE
    if (m >= 1 && n >= 4096) {
        // ---- large inputs: dense integer ranks per objective, then the level loop in sorted space (see SortedView) ----
        unsigned long long *k0, *k1;
        unsigned *i0, *i1, *src, *flags, *dense, *ranks, *rs, *inv, *run_lo, *run_end, *act_a, *act_b, *rank_s, *key_s, *order_s, *pm_pos, *pm_fpos, *cand_rows, *pm_rows;
        unsigned long long *cand_key;
        if ((rc = ws.alloc(&k0, n)) || (rc = ws.alloc(&k1, n)) || (rc = ws.alloc(&i0, n)) || (rc = ws.alloc(&i1, n)) || (rc = ws.alloc(&src, n))
            || (rc = ws.alloc(&flags, n)) || (rc = ws.alloc(&dense, n)) || (rc = ws.alloc(&ranks, static_cast<size_t>(n) * m))
            || (rc = ws.alloc(&rs, static_cast<size_t>(n) * m)) || (rc = ws.alloc(&inv, n)) || (rc = ws.alloc(&run_lo, n))
            || (rc = ws.alloc(&run_end, n)) || (rc = ws.alloc(&act_a, n)) || (rc = ws.alloc(&act_b, n)) || (rc = ws.alloc(&rank_s, n)) || (rc = ws.alloc(&key_s, n)) || (rc = ws.alloc(&order_s, n))
            || (rc = ws.alloc(&pm_pos, n)) || (rc = ws.alloc(&pm_fpos, n)) || (rc = ws.alloc(&cand_rows, static_cast<size_t>(n) * m))
            || (rc = ws.alloc(&pm_rows, static_cast<size_t>(n) * m)) || (rc = ws.alloc(&cand_key, n)))
            return rc;
        size_t b1 = 0, b2 = 0, b3 = 0, b4 = 0, b5 = 0;
        void *tmp = nullptr;
        const StillActive still_active{count};
        thrust::counting_iterator<unsigned> all_positions(0u);
        PGC_CUDA(cub::DeviceSelect::If(nullptr, b4, all_positions, act_a, &meta->nact, static_cast<int>(n), still_active, st));
        PGC_CUDA(cub::DeviceSelect::If(nullptr, b5, act_a, act_b, &meta->nact, static_cast<int>(n), still_active, st));
        PGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b1, k0, k1, i0, i1, static_cast<int>(n), 0, 64, st));
        PGC_CUDA(cub::DeviceScan::InclusiveSum(nullptr, b2, flags, dense, static_cast<int>(n), st));
        PGC_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, b3, k0, k1, static_cast<int>(n), 0, 64, st));
        const size_t tmp_bytes = std::max(std::max(b1, b2), std::max(b3, std::max(b4, b5)));
        if ((rc = ws.alloc_bytes(&tmp, tmp_bytes))) return rc;
        for (int obj = 0; obj < m; ++obj) {
            unsigned *sorted_idx = obj == 0 ? src : i1;
            objective_keys_kernel<<<blocks_for(n, 256), 256, 0, st>>>(d_f, n, m, obj, k0, i0);
            PGC_CUDA(cub::DeviceRadixSort::SortPairs(tmp, b1, k0, k1, i0, sorted_idx, static_cast<int>(n), 0, 64, st));
            key_change_flags_kernel<<<blocks_for(n, 256), 256, 0, st>>>(k1, n, flags);
            PGC_CUDA(cub::DeviceScan::InclusiveSum(tmp, b2, flags, dense, static_cast<int>(n), st));
            scatter_ranks_kernel<<<blocks_for(n, 256), 256, 0, st>>>(sorted_idx, dense, n, m, obj, ranks);
        }
        gather_rows_u32m_kernel<<<blocks_for(static_cast<size_t>(n) * m, 256), 256, 0, st>>>(ranks, src, n, m, rs);
        invert_perm_kernel<<<blocks_for(n, 256), 256, 0, st>>>(src, n, inv);
        // two objectives: sort-based dominance count (PGC_FNDS_COUNT2=0: the all-pairs pass)
        const char *count2_env = std::getenv("PGC_FNDS_COUNT2");
        const bool count2_ok = m == 2 && !(count2_env && count2_env[0] == '0');
        unsigned *c2sorted = nullptr; // per-block sorted second ranks, padded to whole blocks
        if (count2_ok && (rc = ws.alloc(&c2sorted, static_cast<size_t>(blocks_for(n, kCB)) * kCB))) return rc;
        SortedView V{rs,    src,      inv,       run_lo,  run_end, act_a,   0u,      count,       rank_s, key_s, cand,
                     cand_key, cand_rows, order_s, pm_pos,  pm_fpos, pm_rows, d_front_off, meta,   n,     m};
        switch (m) {
#define PGC_MO_SORTED(MM)                                                                                              \
    case MM:                                                                                                           \
        run_bounds_kernel<MM><<<blocks_for(n, 256), 256, 0, st>>>(rs, n, run_lo, run_end);                             \
        if (MM == 2 && count2_ok) {                                                                                    \
            count2_sort_blocks_kernel<<<blocks_for(n, kCB), kCB / 2, 0, st>>>(rs, n, c2sorted);                        \
            fnds_count2_kernel<<<blocks_for(n, 256), 256, 0, st>>>(V, c2sorted, d_dom_count);                          \
        } else {                                                                                                       \
            fnds_count_sorted_kernel<MM><<<gb, kTP, 0, st>>>(V, d_dom_count);                                          \
        }                                                                                                              \
        break;
            PGC_MO_SORTED(1) PGC_MO_SORTED(2) PGC_MO_SORTED(3) PGC_MO_SORTED(4) PGC_MO_SORTED(5) PGC_MO_SORTED(6) PGC_MO_SORTED(7)
            PGC_MO_SORTED(8)
#undef PGC_MO_SORTED
        }
        fnds_front0_sorted_kernel<<<blocks_for(n, 256), 256, 0, st>>>(V);
        fnds_order_sorted_kernel<<<1, 1024, 0, st>>>(V, 1);
        {   // the active list: every position that still has dominators outside the closed fronts
            size_t bytes = tmp_bytes;
            PGC_CUDA(cub::DeviceSelect::If(tmp, bytes, all_positions, act_a, &meta->nact, static_cast<int>(n), still_active, st));
        }
        ctx->launches.fetch_add(5 * m + 8, std::memory_order_relaxed);

        Meta h;
        auto poll = [&]() -> int {
            PGC_CUDA(cudaMemcpyAsync(&h, meta, sizeof(Meta), cudaMemcpyDeviceToHost, st));
            PGC_CUDA(cudaStreamSynchronize(st));
            return PGC_OK;
        };
        auto big_level = [&](int first) -> int { // a level with more than kOrderCap candidates: both orders through CUB
            const unsigned C = h.ncand;
            const unsigned off = first ? 0u : h.front_off + h.front_size;
            size_t bytes = tmp_bytes;
            pack_level_keys_kernel<<<blocks_for(C, 256), 256, 0, st>>>(V, C, k0);
            PGC_CUDA(cub::DeviceRadixSort::SortKeys(tmp, bytes, k0, k1, static_cast<int>(C), 0, 64, st));
            unpack_level_order_kernel<<<blocks_for(C, 256), 256, 0, st>>>(V, k1, C, off, k0);
            bytes = tmp_bytes;
            PGC_CUDA(cub::DeviceRadixSort::SortKeys(tmp, bytes, k0, k1, static_cast<int>(C), 0, 64, st));
            unpack_level_pos_kernel<<<blocks_for(C, 256), 256, 0, st>>>(V, k1, C, off);
            close_big_level_kernel<<<1, 1, 0, st>>>(d_front_off, meta, first);
            ctx->launches.fetch_add(8, std::memory_order_relaxed);
            return poll();
        };
        if ((rc = poll())) return rc;
        if (h.overflow && (rc = big_level(1))) return rc;
        const char *fuse_env = std::getenv("PGC_FNDS_FUSE"); // PGC_FNDS_FUSE=0: always two launches per level (A/B switch)
        const bool fuse_ok = !(fuse_env && fuse_env[0] == '0');
        const char *batch_env = std::getenv("PGC_FNDS_BATCH"); // levels launched between two host polls (experiments)
        const int batch = batch_env ? std::max(1, std::atoi(batch_env)) : kBatch;
        // resident level loop (fnds_persistent_kernel) when every position gets its own thread of a one-block-per-SM grid;
        // PGC_FNDS_PERSIST=0 selects the launch-per-level loop below (A/B switch; results are identical)
        int coop_attr = 0;
        PGC_CUDA(cudaDeviceGetAttribute(&coop_attr, cudaDevAttrCooperativeLaunch, ctx->device));
        const char *persist_env = std::getenv("PGC_FNDS_PERSIST");
        const unsigned pgrid = static_cast<unsigned>(ctx->sm_count);
        const unsigned slice = (n + pgrid - 1) / pgrid;
        bool persist = coop_attr != 0 && slice <= static_cast<unsigned>(kPersistThreads) && !(persist_env && persist_env[0] == '0');
        while (h.assigned < n && !h.done) {
            if (h.front_size == 0) {
                set_error("fast_non_dominated_sorting: internal error, empty front with %u of %u points assigned", h.assigned, n);
                return PGC_ERR_CUDA;
            }
            if (persist) {
                void *fn = nullptr;
                switch (m) {
#define PGC_MO_PERSIST(MM) case MM: fn = reinterpret_cast<void *>(fnds_persistent_kernel<MM>); break;
                    PGC_MO_PERSIST(1) PGC_MO_PERSIST(2) PGC_MO_PERSIST(3) PGC_MO_PERSIST(4) PGC_MO_PERSIST(5) PGC_MO_PERSIST(6)
                    PGC_MO_PERSIST(7) PGC_MO_PERSIST(8)
#undef PGC_MO_PERSIST
                }
                const char *big_env = std::getenv("PGC_FNDS_BIG_INKERNEL"); // 0: big levels always through the host / CUB path
                unsigned big_arg = (big_env && big_env[0] == '0') ? 0u : 1u;
                // test hooks: PGC_FNDS_INJECT_STUCK=1 makes block 0 report a lost wake-up (the watchdog path must surface as an
                // error, not a hang); PGC_FNDS_FORCE_NOCOOP=1 behaves as if the cooperative launch had been refused
                const char *stuck_env = std::getenv("PGC_FNDS_INJECT_STUCK"), *nocoop_env = std::getenv("PGC_FNDS_FORCE_NOCOOP");
                if (stuck_env && stuck_env[0] == '1') big_arg |= 2u;
                const bool refuse = nocoop_env && nocoop_env[0] == '1';
                void *args[] = {&V, &big_arg};
                if (refuse || cudaLaunchCooperativeKernel(fn, dim3(pgrid), dim3(kPersistThreads), args, 0, st) != cudaSuccess) {
                    cudaGetLastError(); // the grid cannot be co-resident here (e.g. a partitioned device): launch-per-level loop instead
                    persist = false;
                    continue;
                }
                ctx->launches.fetch_add(1, std::memory_order_relaxed);
                if ((rc = poll())) return rc;
                if (h.stuck) {
                    set_error("fast_non_dominated_sorting: internal error, the resident level loop lost a wake-up at level %u", h.level);
                    return PGC_ERR_CUDA;
                }
                if (h.overflow && (rc = big_level(0))) return rc;
                continue;
            }
            const unsigned grid = std::max(1u, blocks_for(h.nact, kPT)); // h.nact: length of the list at the last compaction
            V.nact_cap = h.nact;
            // small fronts: one fused launch per level; a front above kFusedCap switches the next batches to two launches
            const bool fused = fuse_ok && h.front_size <= kFusedCap;
            for (int b = 0; b < batch; ++b) {
                switch (m) {
#define PGC_MO_PEEL(MM)                                                                                                \
    case MM:                                                                                                           \
        if (fused) fnds_peel_sorted_kernel<MM, true><<<grid, kPT, 0, st>>>(V);                                         \
        else fnds_peel_sorted_kernel<MM, false><<<grid, kPT, 0, st>>>(V);                                              \
        break;
                    PGC_MO_PEEL(1) PGC_MO_PEEL(2) PGC_MO_PEEL(3) PGC_MO_PEEL(4) PGC_MO_PEEL(5) PGC_MO_PEEL(6) PGC_MO_PEEL(7) PGC_MO_PEEL(8)
#undef PGC_MO_PEEL
                }
                if (!fused) fnds_order_sorted_kernel<<<1, 1024, 0, st>>>(V, 0);
            }
            {   // drop the points that joined a front during this batch (order is kept: the list stays ascending)
                unsigned *next = V.act == act_a ? act_b : act_a;
                size_t bytes = tmp_bytes;
                PGC_CUDA(cub::DeviceSelect::If(tmp, bytes, V.act, next, &meta->nact, static_cast<int>(h.nact), still_active, st));
                V.act = next;
            }
            ctx->launches.fetch_add((fused ? 1 : 2) * batch + 2, std::memory_order_relaxed);
            if ((rc = poll())) return rc;
            if (h.overflow == 2) { // a fused pass left a level of more than kFusedCap candidates open
                PGC_CUDA(cudaMemsetAsync(&meta->overflow, 0, sizeof(unsigned), st));
                fnds_order_sorted_kernel<<<1, 1024, 0, st>>>(V, 0);
                ctx->launches.fetch_add(1, std::memory_order_relaxed);
                if ((rc = poll())) return rc;
            }
            if (h.overflow && (rc = big_level(0))) return rc;
        }
        fnds_unsort_kernel<<<blocks_for(n, 256), 256, 0, st>>>(V, d_rank, d_key_out, d_order);
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        PGC_CUDA(cudaGetLastError());
        PGC_CUDA(cudaStreamSynchronize(st)); // the workspace is released on return
        if (nfronts_out) *nfronts_out = h.nfronts;
        return PGC_OK;
    }
    count_k<<<gb, kTP, 0, st>>>(d_f, n, count, d_dom_count);
    fnds_front0_kernel<<<blocks_for(n, 256), 256, 0, st>>>(n, count, d_rank, key, cand, meta);
    ctx->launches.fetch_add(3, std::memory_order_relaxed);

    auto big_order = [&](int first) -> int { // order a level with more than kOrderCap candidates through CUB
        Meta h;
        PGC_CUDA(cudaMemcpyAsync(&h, meta, sizeof(Meta), cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        const unsigned C = h.ncand;
        if (!pk0) {
            int r2;
            if ((r2 = ws.alloc(&pk0, n)) || (r2 = ws.alloc(&pk1, n))) return r2;
            PGC_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, pk0, pk1, static_cast<int>(n), 0, 64, st));
            if ((r2 = ws.alloc_bytes(&cub_tmp, cub_bytes))) return r2;
        }
        pack_keys_kernel<<<blocks_for(C, 256), 256, 0, st>>>(cand, key, C, pk0);
        size_t bytes = cub_bytes;
        PGC_CUDA(cub::DeviceRadixSort::SortKeys(cub_tmp, bytes, pk0, pk1, static_cast<int>(C), 0, 64, st));
        const unsigned off = first ? 0u : h.front_off + h.front_size;
        unpack_front_kernel<<<blocks_for(C, 256), 256, 0, st>>>(pk1, C, off, d_order);
        close_big_level_kernel<<<1, 1, 0, st>>>(d_front_off, meta, first);
        ctx->launches.fetch_add(5, std::memory_order_relaxed);
        return PGC_OK;
    };

    fnds_order_kernel<<<1, 1024, 0, st>>>(cand, key, d_order, d_front_off, meta, 1);
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    Meta h;
    PGC_CUDA(cudaMemcpyAsync(&h, meta, sizeof(Meta), cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    if (h.overflow) {
        if ((rc = big_order(1))) return rc;
        PGC_CUDA(cudaMemcpyAsync(&h, meta, sizeof(Meta), cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
    }
    while (h.assigned < n && !h.done) {
        if (h.front_size == 0) {
            set_error("fast_non_dominated_sorting: internal error, empty front with %u of %u points assigned", h.assigned, n);
            return PGC_ERR_CUDA;
        }
        for (int b = 0; b < kBatch; ++b) {
            peel_k<<<gb, kTP, 0, st>>>(d_f, n, d_order, count, d_rank, key, cand, meta);
            fnds_order_kernel<<<1, 1024, 0, st>>>(cand, key, d_order, d_front_off, meta, 0);
        }
        ctx->launches.fetch_add(2 * kBatch, std::memory_order_relaxed);
        PGC_CUDA(cudaMemcpyAsync(&h, meta, sizeof(Meta), cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        if (h.overflow) {
            if ((rc = big_order(0))) return rc;
            PGC_CUDA(cudaMemcpyAsync(&h, meta, sizeof(Meta), cudaMemcpyDeviceToHost, st));
            PGC_CUDA(cudaStreamSynchronize(st));
        }
    }
    PGC_CUDA(cudaGetLastError());
    if (nfronts_out) *nfronts_out = h.nfronts;
    return PGC_OK;
}

// Crowding distance of every front at once (segments of `d_order` delimited by `d_front_off`), written at the points'
// own indices in d_cd.  d_order == nullptr: one segment = all n points in index order (plain crowding_distance()).
int crowding_device(pgc_ctx *ctx, const double *d_f, size_t n_, size_t m_, const unsigned *d_order, const unsigned *d_front_off,
                    unsigned nfronts, int small_rule, double *d_cd, cudaStream_t st)
{
    const unsigned n = static_cast<unsigned>(n_);
    const int m = static_cast<int>(m_);
    Workspace ws(ctx, st);
    double *keys_in, *keys_out;
    unsigned *vals_in, *vals_out, *seg_of, *one_off = nullptr;
    int rc;
    if ((rc = ws.alloc(&keys_in, n)) || (rc = ws.alloc(&keys_out, n)) || (rc = ws.alloc(&vals_in, n)) || (rc = ws.alloc(&vals_out, n))
        || (rc = ws.alloc(&seg_of, n)))
        return rc;
    if (!d_order) {
        if ((rc = ws.alloc(&one_off, 2))) return rc;
        const unsigned h[2] = {0u, n};
        PGC_CUDA(cudaMemcpyAsync(one_off, h, sizeof(h), cudaMemcpyHostToDevice, st));
        d_front_off = one_off;
        nfronts = 1;
    }
    fill_double_kernel<<<blocks_for(n, 256), 256, 0, st>>>(d_cd, n, 0.0);
    segment_ids_kernel<<<nfronts, 128, 0, st>>>(d_front_off, nfronts, seg_of);
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    PGC_CUDA(cub::DeviceSegmentedSort::StableSortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, vals_out, static_cast<int>(n),
                                                       static_cast<int>(nfronts), d_front_off, d_front_off + 1, st));
    if ((rc = ws.alloc_bytes(&tmp, tmp_bytes))) return rc;
    unsigned *seq = nullptr; // the sequence objective `obj` is sorted from
    if ((rc = ws.alloc(&seq, n))) return rc;
    for (int obj = 0; obj < m; ++obj) {
        // The reference sorts ONE index vector objective after objective without resetting it (multi_objective.cpp:296-313), so among
        // equal values of objective k the boundary point (and every neighbour) is decided by the order objective k - 1 left.  The
        // sorts are stable here, so starting each one from the previous result reproduces that; objective 0 starts from the
        // front's own order.
        gather_objective_kernel<<<blocks_for(n, 256), 256, 0, st>>>(d_f, m, obj, obj == 0 ? d_order : seq, n, keys_in, vals_in);
        size_t bytes = tmp_bytes;
        PGC_CUDA(cub::DeviceSegmentedSort::StableSortPairs(tmp, bytes, keys_in, keys_out, vals_in, vals_out, static_cast<int>(n),
                                                           static_cast<int>(nfronts), d_front_off, d_front_off + 1, st));
        crowding_accumulate_kernel<<<blocks_for(n, 256), 256, 0, st>>>(keys_out, vals_out, seg_of, d_front_off, n, d_cd);
        if (obj + 1 < m) PGC_CUDA(cudaMemcpyAsync(seq, vals_out, sizeof(unsigned) * n, cudaMemcpyDeviceToDevice, st));
        ctx->launches.fetch_add(4, std::memory_order_relaxed);
    }
    if (small_rule && d_order) {
        small_front_rule_kernel<<<blocks_for(nfronts, 128), 128, 0, st>>>(d_order, d_front_off, nfronts, small_rule, d_cd);
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
    }
    PGC_CUDA(cudaGetLastError());
    PGC_CUDA(cudaStreamSynchronize(st)); // workspace is freed on return
    return PGC_OK;
}

// stable sort of `count` indices (d_idx, in place) by crowding distance descending (greater_than_f order)
static int sort_by_cd_desc(pgc_ctx *ctx, const double *d_cd, unsigned *d_idx, unsigned count, cudaStream_t st)
{
    if (count < 2) return PGC_OK;
    Workspace ws(ctx, st);
    unsigned long long *k0, *k1;
    unsigned *v1;
    int rc;
    if ((rc = ws.alloc(&k0, count)) || (rc = ws.alloc(&k1, count)) || (rc = ws.alloc(&v1, count))) return rc;
    cd_desc_key_kernel<<<blocks_for(count, 256), 256, 0, st>>>(d_cd, d_idx, count, k0);
    void *tmp = nullptr;
    size_t bytes = 0;
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0, k1, d_idx, v1, static_cast<int>(count), 0, 64, st));
    if ((rc = ws.alloc_bytes(&tmp, bytes))) return rc;
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, k0, k1, d_idx, v1, static_cast<int>(count), 0, 64, st));
    PGC_CUDA(cudaMemcpyAsync(d_idx, v1, sizeof(unsigned) * count, cudaMemcpyDeviceToDevice, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    ctx->launches.fetch_add(3, std::memory_order_relaxed);
    return PGC_OK;
}

// select_best_N_mo, multi_objective.cpp:344-396, on device-resident f; d_out receives min(N, n) indices
__global__ void add_offset_kernel(const unsigned *in, unsigned n, unsigned off, unsigned *out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + off;
}

// `ranking` (optional): fast_non_dominated_sorting of the SELECTED individuals, indexed by their position in d_out, derived
// from this run instead of a second sort.  Why it is exact: a survivor's dominators all sit in lower fronts, which survive
// whole, so ranks carry over; whole fronts keep their member sequence (d_out lists them in front order, and the order inside
// a front is (position of the last dominator in the previous front, index) - positions are unchanged and the new indices
// ascend in the old order); only the cut front must be re-ordered, by (its old keys, new index).
int select_best_device(pgc_ctx *ctx, const double *d_f, size_t n_, size_t m_, size_t N_, unsigned *d_out, unsigned *nout,
                       cudaStream_t st, SelectedRanking *ranking)
{
    if (ranking) ranking->valid = false;
    const unsigned n = static_cast<unsigned>(n_);
    if (N_ == 0 || n == 0) { // :346-351
        *nout = 0;
        return PGC_OK;
    }
    if (n == 1) { // :352-354
        const unsigned z = 0;
        PGC_CUDA(cudaMemcpyAsync(d_out, &z, sizeof(z), cudaMemcpyHostToDevice, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        *nout = 1;
        return PGC_OK;
    }
    if (N_ >= n) { // :355-359
        iota_kernel<<<blocks_for(n, 256), 256, 0, st>>>(d_out, n);
        *nout = n;
        return PGC_OK;
    }
    const unsigned N = static_cast<unsigned>(N_);
    Workspace ws(ctx, st);
    unsigned *rank, *order, *foff, *key;
    int rc;
    if ((rc = ws.alloc(&rank, n)) || (rc = ws.alloc(&order, n)) || (rc = ws.alloc(&foff, n + 1)) || (rc = ws.alloc(&key, n))) return rc;
    unsigned nfronts = 0;
    // only the fronts that hold the best N are needed: the level loop stops once N points sit in closed fronts
    if ((rc = fnds_device(ctx, d_f, n, m_, rank, nullptr, order, foff, &nfronts, st, N, key))) return rc;
    std::vector<unsigned> hoff(nfronts + 1);
    PGC_CUDA(cudaMemcpyAsync(hoff.data(), foff, sizeof(unsigned) * (nfronts + 1), cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    unsigned front_id = 0, taken = 0; // whole fronts while they fit, :365-377
    while (front_id < nfronts && taken + (hoff[front_id + 1] - hoff[front_id]) <= N) {
        taken += hoff[front_id + 1] - hoff[front_id];
        ++front_id;
    }
    PGC_CUDA(cudaMemcpyAsync(d_out, order, sizeof(unsigned) * taken, cudaMemcpyDeviceToDevice, st));
    if (taken < N) { // the cut front, by crowding distance descending, :378-394
        const unsigned b = hoff[front_id], sz = hoff[front_id + 1] - b;
        double *cd, *fsub;
        unsigned *idx;
        if ((rc = ws.alloc(&cd, sz)) || (rc = ws.alloc(&fsub, static_cast<size_t>(sz) * m_)) || (rc = ws.alloc(&idx, sz))) return rc;
        // crowding_distance() of the front's own fitness list (local indices 0..sz-1 in front order)
        gather_rows_kernel<<<blocks_for(static_cast<size_t>(sz) * m_, 256), 256, 0, st>>>(d_f, order + b, sz, static_cast<int>(m_), fsub);
        if ((rc = crowding_device(ctx, fsub, sz, m_, nullptr, nullptr, 1, 0, cd, st))) return rc;
        iota_kernel<<<blocks_for(sz, 256), 256, 0, st>>>(idx, sz);
        if ((rc = sort_by_cd_desc(ctx, cd, idx, sz, st))) return rc;
        gather_u32_kernel<<<blocks_for(N - taken, 256), 256, 0, st>>>(order + b, idx, N - taken, d_out + taken);
    }
    if (ranking) {
        // whole fronts: position i of d_out is new individual i, fronts keep their offsets
        const unsigned nf_new = front_id + (taken < N ? 1u : 0u);
        std::vector<unsigned> noff(hoff.begin(), hoff.begin() + front_id + 1);
        if (taken < N) noff.push_back(N);
        PGC_CUDA(cudaMemcpyAsync(ranking->front_off, noff.data(), sizeof(unsigned) * noff.size(), cudaMemcpyHostToDevice, st));
        iota_kernel<<<blocks_for(N, 256), 256, 0, st>>>(ranking->order, N);
        segment_ids_kernel<<<nf_new, 256, 0, st>>>(ranking->front_off, nf_new, ranking->rank);
        if (taken < N && front_id > 0) { // the cut front: stable sort of its survivors by their old keys
            const unsigned c = N - taken;
            unsigned *k0, *k1, *j0, *j1;
            if ((rc = ws.alloc(&k0, c)) || (rc = ws.alloc(&k1, c)) || (rc = ws.alloc(&j0, c)) || (rc = ws.alloc(&j1, c))) return rc;
            gather_u32_kernel<<<blocks_for(c, 256), 256, 0, st>>>(key, d_out + taken, c, k0);
            iota_kernel<<<blocks_for(c, 256), 256, 0, st>>>(j0, c);
            size_t bytes = 0;
            void *tmp = nullptr;
            PGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0, k1, j0, j1, static_cast<int>(c), 0, 32, st));
            if ((rc = ws.alloc_bytes(&tmp, bytes))) return rc;
            PGC_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, k0, k1, j0, j1, static_cast<int>(c), 0, 32, st));
            add_offset_kernel<<<blocks_for(c, 256), 256, 0, st>>>(j1, c, taken, ranking->order + taken);
        }
        PGC_CUDA(cudaStreamSynchronize(st)); // noff is a host vector
        ranking->nfronts = nf_new;
        ranking->valid = true;
    }
    PGC_CUDA(cudaGetLastError());
    PGC_CUDA(cudaStreamSynchronize(st));
    *nout = N;
    return PGC_OK;
}

// sort_population_mo, multi_objective.cpp:425-465: indices by (rank ascending, crowding distance descending)
// `limit` (0 = everything): only d_out[0 .. limit) is needed.  The level loop then stops once `limit` points sit in closed fronts,
// and the points of those fronts - every other point ranks behind them - are ordered exactly as the full sort would order them
// (rank, then crowding distance descending, ties in index order); d_out receives all of them (at least `limit`, at most n).
int sort_population_device(pgc_ctx *ctx, const double *d_f, size_t n_, size_t m_, unsigned *d_out, cudaStream_t st, size_t limit)
{
    const unsigned n = static_cast<unsigned>(n_);
    if (n == 0) return PGC_OK;
    if (n == 1) {
        const unsigned z = 0;
        PGC_CUDA(cudaMemcpyAsync(d_out, &z, sizeof(z), cudaMemcpyHostToDevice, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        return PGC_OK;
    }
    const bool partial = limit > 0 && limit < n;
    Workspace ws(ctx, st);
    unsigned *rank, *order, *foff, *rk_in, *rk_out, *v_out, *key = nullptr;
    double *cd;
    int rc;
    if ((rc = ws.alloc(&rank, n)) || (rc = ws.alloc(&order, n)) || (rc = ws.alloc(&foff, n + 1)) || (rc = ws.alloc(&cd, n))
        || (rc = ws.alloc(&rk_in, n)) || (rc = ws.alloc(&rk_out, n)) || (rc = ws.alloc(&v_out, n)))
        return rc;
    if (partial && (rc = ws.alloc(&key, n))) return rc;
    unsigned nfronts = 0, count = n;
    if ((rc = fnds_device(ctx, d_f, n, m_, rank, nullptr, order, foff, &nfronts, st, partial ? static_cast<unsigned>(limit) : 0u, key))) return rc;
    if (partial) {
        PGC_CUDA(cudaMemcpyAsync(&count, foff + nfronts, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        fill_double_kernel<<<blocks_for(n, 256), 256, 0, st>>>(cd, n, 0.0); // crowding_device clears only the entries it is given
    }
    if ((rc = crowding_device(ctx, d_f, count, m_, order, foff, nfronts, 2, cd, st))) return rc;
    if (partial) { // the points of the closed fronts, in index order
        void *tmp = nullptr;
        size_t bytes = 0;
        PGC_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, order, d_out, static_cast<int>(count), 0, 32, st));
        if ((rc = ws.alloc_bytes(&tmp, bytes))) return rc;
        PGC_CUDA(cub::DeviceRadixSort::SortKeys(tmp, bytes, order, d_out, static_cast<int>(count), 0, 32, st));
    } else {
        iota_kernel<<<blocks_for(n, 256), 256, 0, st>>>(d_out, n);
    }
    if ((rc = sort_by_cd_desc(ctx, cd, d_out, count, st))) return rc;
    gather_u32_kernel<<<blocks_for(count, 256), 256, 0, st>>>(rank, d_out, count, rk_in);
    void *tmp = nullptr;
    size_t bytes = 0;
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, rk_in, rk_out, d_out, v_out, static_cast<int>(count), 0, 32, st));
    if ((rc = ws.alloc_bytes(&tmp, bytes))) return rc;
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, rk_in, rk_out, d_out, v_out, static_cast<int>(count), 0, 32, st));
    PGC_CUDA(cudaMemcpyAsync(d_out, v_out, sizeof(unsigned) * count, cudaMemcpyDeviceToDevice, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    return PGC_OK;
}

} // namespace pgc
