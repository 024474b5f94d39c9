// de.cu - differential evolution family (de, sade, de1220) as a GENERATIONAL device loop.
//
// Reference: src/algorithms/de.cpp:76-345 (10 variants :154-275, selection :277-299, exit :302-321),
// src/algorithms/sade.cpp:78-560 and src/algorithms/de1220.cpp:80-600 (18 variants :199-505, F/CR initialisation :147-165,
// jDE adaptation :193-197, iDE adaptation inside every variant, feasibility :507-513, selection :515-536).
//
// The reference evaluates `prob.fitness(tmp)` one individual at a time inside the population loop (no bfe hook).
// Its trial vectors only read `popold` and `gbIter` (the previous generation), so building all NP trials first, evaluating
// them as one batch and then applying the selection is the same algorithm, with two documented differences: the self-adapted
// F/CR/variant of individuals accepted earlier IN THE SAME generation are not yet visible to later individuals (iDE reads
// m_F[r[k]]), and the global best (gbX, gbF, gbCR) is updated once per generation with the reference's tie rule (`<=`, so the
// last of equal minima wins).
// Draws: individual i owns the Philox substream (seed, kTagDe, generation, i) and consumes it in the reference's order:
// 7 (de: 5) Durstenfeld index picks, [de1220: variant gate], [jDE: F gate (+1), CR gate (+1)] | [iDE: normals, Box-Muller],
// start gene, crossover draws, one draw per out-of-bounds gene.  uniform_int(a, b) = a + floor(u * (b - a + 1)).
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <algorithm>
#include <vector>

#include <cooperative_groups.h>

#include "pgc_internal.cuh"
#include "philox.cuh"
#include "simple_device.cuh"

namespace cg = cooperative_groups;

namespace pgc
{

namespace
{

struct DeConfig {
    unsigned algo;          // 0 = de, 1 = sade, 2 = de1220
    unsigned variant;       // de / sade: mutation variant
    unsigned variant_adptv; // sade / de1220: 1 = jDE, 2 = iDE
    double F, CR;           // de
    unsigned n_allowed;
    unsigned allowed[18];
};

struct TrialParams {
    const double *popold; // [NP x dim]
    const double *gbIter; // [dim]
    const double *lb, *ub;
    const double *F_in, *CR_in;     // per individual (sade / de1220)
    const unsigned *variant_in;     // per individual (de1220)
    const double *gbIterF, *gbIterCR; // device scalars
    const unsigned *stopped;          // device flag: an exit condition fired in an earlier generation of this batch of launches
    double *trial;                  // [NP x dim]
    double *F_out, *CR_out;
    unsigned *variant_out;
    unsigned NP, dim;
    unsigned long long seed;
    const unsigned *gen_base, *gens_done; // device: generation index = *gen_base + *gens_done (no per-launch parameter: graph-replayable)
    DeConfig cfg;
};

__device__ __forceinline__ double normal01(PhiloxStream &rs) // Box-Muller on two uniforms
{
    const double u1 = 1.0 - rs.next();
    const double u2 = rs.next();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

__device__ __forceinline__ unsigned uint_below(PhiloxStream &rs, unsigned n) // uniform in [0, n)
{
    const unsigned v = static_cast<unsigned>(rs.next() * static_cast<double>(n));
    return v < n ? v : n - 1;
}

// mutation formulas.  de.cpp and de1220/sade.cpp share variants 1-3 / 6-8; 4,5,9,10 differ in form (de: one product).
__device__ __forceinline__ double mutate(unsigned algo, unsigned base, double t, double gb, const double *p, double pi, double F)
{
    // p[k] = popold[r[k]][n], pi = popold[i][n], t = tmp[n] (== pi)
    switch (base) {
        case 1: return gb + F * (p[1] - p[2]);
        case 2: return p[0] + F * (p[1] - p[2]);
        case 3: return t + F * (gb - t) + F * (p[0] - p[1]);
        case 4: return algo == 0 ? gb + (p[0] + p[1] - p[2] - p[3]) * F : gb + (p[0] - p[1]) * F + (p[2] - p[3]) * F;
        case 5: return algo == 0 ? p[4] + (p[0] + p[1] - p[2] - p[3]) * F : p[4] + (p[0] - p[1]) * F + (p[2] - p[3]) * F;
        case 6: return p[0] + (p[1] - p[2]) * F + (p[3] - p[4]) * F + (p[5] - p[6]) * F; // variants 11/12
        case 7: return gb + (p[1] - p[2]) * F + (p[3] - p[4]) * F + (p[5] - p[6]) * F;   // 13/14
        case 8: return p[0] + (p[1] - pi) * F + (p[2] - p[3]) * F;                       // 15/16
        default: return p[0] + (p[1] - pi) * F - (p[2] - gb) * F;                        // 17/18
    }
}

// variant -> (formula id, exponential crossover?)
__device__ __forceinline__ void decode_variant(unsigned v, unsigned &base, bool &expo)
{
    if (v <= 10u) {
        expo = v <= 5u;
        base = expo ? v : v - 5u;
    } else {
        expo = (v & 1u) != 0u; // 11, 13, 15, 17 exponential; 12, 14, 16, 18 binomial
        base = 6u + (v - 11u) / 2u;
    }
}

// ---- trial vector of individual i: the scalar prelude (index picks, variant gate, F / CR) -------------------------------------
// DRAW is the individual's Philox substream read in order (next()): the thread-per-individual kernel passes the stream itself,
// the warp-per-individual kernel a view of draws its lanes computed in parallel.
struct TrialHead {
    unsigned r[7];
    double F, CR;
    unsigned variant, base;
    bool expo;
};

template <class DRAW>
__device__ __forceinline__ unsigned uint_below_from(DRAW &rs, unsigned n) // uniform in [0, n)
{
    const unsigned v = static_cast<unsigned>(rs.next() * static_cast<double>(n));
    return v < n ? v : n - 1;
}

template <class DRAW>
__device__ __forceinline__ double normal01_from(DRAW &rs) // Box-Muller on two uniforms
{
    const double u1 = 1.0 - rs.next();
    const double u2 = rs.next();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

template <class DRAW>
__device__ __forceinline__ TrialHead trial_head(const TrialParams &P, unsigned i, DRAW &rs)
{
    TrialHead H;
    const unsigned NP = P.NP;
    // the individual's own self-adapted state: loaded before the draws decide whether it is used (latency off the chain)
    const double own_F = (P.cfg.algo != 0u && P.cfg.variant_adptv == 1u) ? P.F_in[i] : 0.0;
    const double own_CR = (P.cfg.algo != 0u && P.cfg.variant_adptv == 1u) ? P.CR_in[i] : 0.0;
    const unsigned own_variant = P.cfg.algo == 2u ? P.variant_in[i] : 0u;
    // Durstenfeld partial shuffle of 0..NP-1 (de.cpp:143-149, de1220.cpp:181-187) on a virtual array: only the picked
    // positions are ever overwritten (with the value of the current last position)
    const unsigned npick = P.cfg.algo == 0 ? 5u : 7u;
    unsigned pos[7], val[7];
    for (unsigned j = 0; j < npick; ++j) {
        const unsigned last = NP - 1u - j;
        const unsigned idx = uint_below_from(rs, NP - j);
        unsigned at_idx = idx, at_last = last;
        for (unsigned k = 0; k < j; ++k) {
            if (pos[k] == idx) at_idx = val[k];
            if (pos[k] == last) at_last = val[k];
        }
        H.r[j] = at_idx;
        pos[j] = idx;
        val[j] = at_last;
    }
    for (unsigned j = npick; j < 7u; ++j) H.r[j] = 0;
    const unsigned *r = H.r;

    double F = P.cfg.F, CR = P.cfg.CR;
    unsigned variant = P.cfg.variant;
    if (P.cfg.algo == 2u) { // de1220.cpp:192
        variant = (rs.next() < 0.9) ? own_variant : P.cfg.allowed[uint_below_from(rs, P.cfg.n_allowed)];
    }
    if (P.cfg.algo != 0u && P.cfg.variant_adptv == 1u) { // jDE, de1220.cpp:193-196 / sade.cpp:178-182
        F = (rs.next() < 0.9) ? own_F : rs.next() * 0.9 + 0.1;
        CR = (rs.next() < 0.9) ? own_CR : rs.next();
    }
    unsigned base;
    bool expo;
    decode_variant(variant, base, expo);
    if (P.cfg.algo != 0u && P.cfg.variant_adptv == 2u) { // iDE: the per-variant F/CR formulas (e.g. de1220.cpp:201-202)
        // the normal draws are taken in the order they appear in the reference's expressions, F first then CR (the C++
        // evaluation order inside one expression is unspecified, so it is fixed here explicitly)
        const double *mF = P.F_in, *mC = P.CR_in;
        const double gF = *P.gbIterF, gC = *P.gbIterCR;
        double a1, a2, a3, c1, c2;
        switch (base) {
            case 1:
                a1 = normal01_from(rs); c1 = normal01_from(rs);
                F = gF + a1 * 0.5 * (mF[r[1]] - mF[r[2]]);
                CR = gC + c1 * 0.5 * (mC[r[1]] - mC[r[2]]);
                break;
            case 2:
                a1 = normal01_from(rs); c1 = normal01_from(rs);
                F = mF[r[0]] + a1 * 0.5 * (mF[r[1]] - mF[r[2]]);
                CR = mC[r[0]] + c1 * 0.5 * (mC[r[1]] - mC[r[2]]);
                break;
            case 3:
                a1 = normal01_from(rs); a2 = normal01_from(rs); c1 = normal01_from(rs); c2 = normal01_from(rs);
                F = mF[i] + a1 * 0.5 * (gF - mF[i]) + a2 * 0.5 * (mF[r[0]] - mF[r[1]]);
                CR = mC[i] + c1 * 0.5 * (gC - mC[i]) + c2 * 0.5 * (mC[r[0]] - mC[r[1]]);
                break;
            case 4:
                a1 = normal01_from(rs); a2 = normal01_from(rs); c1 = normal01_from(rs); c2 = normal01_from(rs);
                F = gF + a1 * 0.5 * (mF[r[0]] - mF[r[1]]) + a2 * 0.5 * (mF[r[2]] - mF[r[3]]);
                CR = gC + c1 * 0.5 * (mC[r[0]] - mC[r[1]]) + c2 * 0.5 * (mC[r[2]] - mC[r[3]]);
                break;
            case 5:
                a1 = normal01_from(rs); a2 = normal01_from(rs); c1 = normal01_from(rs); c2 = normal01_from(rs);
                F = mF[r[4]] + a1 * 0.5 * (mF[r[0]] - mF[r[1]]) + a2 * 0.5 * (mF[r[2]] - mF[r[3]]);
                CR = mC[r[4]] + c1 * 0.5 * (mC[r[0]] - mC[r[1]]) + c2 * 0.5 * (mC[r[2]] - mC[r[3]]);
                break;
            case 6:
                a1 = normal01_from(rs); a2 = normal01_from(rs); a3 = normal01_from(rs); c1 = normal01_from(rs);
                F = mF[r[0]] + a1 * 0.5 * (mF[r[1]] - mF[r[2]]) + a2 * 0.5 * (mF[r[3]] - mF[r[4]]) + a3 * 0.5 * (mF[r[5]] - mF[r[6]]);
                CR = mC[r[4]] + c1 * 0.5 * (mC[r[0]] + mC[r[1]] - mC[r[2]] - mC[r[3]]);
                break;
            case 7:
                a1 = normal01_from(rs); a2 = normal01_from(rs); a3 = normal01_from(rs); c1 = normal01_from(rs);
                F = gF + a1 * 0.5 * (mF[r[1]] - mF[r[2]]) + a2 * 0.5 * (mF[r[3]] - mF[r[4]]) + a3 * 0.5 * (mF[r[5]] - mF[r[6]]);
                CR = gC + c1 * 0.5 * (mC[r[0]] + mC[r[1]] - mC[r[2]] - mC[r[3]]);
                break;
            case 8:
                a1 = normal01_from(rs); a2 = normal01_from(rs); c1 = normal01_from(rs); c2 = normal01_from(rs);
                F = mF[r[0]] + a1 * 0.5 * (mF[r[1]] - mF[i]) + a2 * 0.5 * (mF[r[3]] - mF[r[4]]);
                CR = mC[r[0]] + c1 * 0.5 * (mC[r[1]] - mC[i]) + c2 * 0.5 * (mC[r[3]] - mC[r[4]]);
                break;
            default:
                a1 = normal01_from(rs); a2 = normal01_from(rs); c1 = normal01_from(rs); c2 = normal01_from(rs);
                F = mF[r[0]] + a1 * 0.5 * (mF[r[1]] - mF[i]) - a2 * 0.5 * (mF[r[2]] - gF);
                CR = mC[r[0]] + c1 * 0.5 * (mC[r[1]] - mC[i]) - c2 * 0.5 * (mC[r[3]] - gC);
        }
    }
    H.F = F;
    H.CR = CR;
    H.variant = variant;
    H.base = base;
    H.expo = expo;
    return H;
}

// one thread per individual (large populations: every thread has its own long dependent chain, the SMs are full anyway)
__global__ void de_trial_kernel(const TrialParams P)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.NP || *P.stopped) return;
    const unsigned dim = P.dim;
    PhiloxStream rs(P.seed, kTagDe, *P.gen_base + *P.gens_done, i);
    const TrialHead H = trial_head(P, i, rs);
    const unsigned *r = H.r;
    const double F = H.F, CR = H.CR;
    const unsigned base = H.base;
    const bool expo = H.expo;

    const double *xi = P.popold + static_cast<size_t>(i) * dim;
    double *tmp = P.trial + static_cast<size_t>(i) * dim;
    for (unsigned d = 0; d < dim; ++d) tmp[d] = xi[d];
    unsigned n = uint_below(rs, dim); // c_idx(m_e)
    auto gene = [&](unsigned nn) {
        double p[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) p[k] = P.popold[static_cast<size_t>(r[k]) * dim + nn];
        return mutate(P.cfg.algo, base, tmp[nn], P.gbIter[nn], p, xi[nn], F);
    };
    if (expo) {
        unsigned L = 0u;
        do {
            tmp[n] = gene(n);
            n = (n + 1u) % dim;
            ++L;
        } while ((rs.next() < CR) && (L < dim));
    } else {
        for (unsigned L = 0u; L < dim; ++L) {
            if ((rs.next() < CR) || L + 1u == dim) tmp[n] = gene(n);
            n = (n + 1u) % dim;
        }
    }
    // feasibility: out-of-bounds genes are resampled uniformly, de1220.cpp:507-513 / force_bounds_random generic.hpp:403-412
    for (unsigned j = 0; j < dim; ++j) {
        if ((tmp[j] < P.lb[j]) || (tmp[j] > P.ub[j])) {
            const double lo = P.lb[j], hi = P.ub[j];
            tmp[j] = (lo == hi) ? lo : (hi - lo) * rs.next() + lo;
        }
    }
    if (P.F_out) {
        P.F_out[i] = H.F;
        P.CR_out[i] = H.CR;
    }
    if (P.variant_out) P.variant_out[i] = H.variant;
}

// ---- one WARP per individual (launch-bound populations): the same trial vector, bit for bit, with the substream read by
// position instead of in sequence.  Draw k of a Philox substream is addressable (block k / 2, half k % 2), so
//   - the lanes compute draws 0..31 in one go and the scalar prelude reads them by shuffle (it needs at most 21);
//   - binomial crossover: the draw of the L-th visited gene is draw c0 + L, one lane per gene;
//   - exponential crossover: the run length is the first L >= 1 with draw c0 + L - 1 >= CR (a ballot), capped at dim;
//   - feasibility: the out-of-bounds genes take the draws after the crossover in gene order (ballot + popc prefix).
__device__ __forceinline__ double philox_draw_at(unsigned long long seed, unsigned generation, unsigned index, unsigned k)
{
    const Philox4 r = philox4x32_10(k >> 1, index, generation, kTagDe, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    const uint64_t bits = (k & 1u) ? ((static_cast<uint64_t>(r.v[3]) << 32) | r.v[2]) : ((static_cast<uint64_t>(r.v[1]) << 32) | r.v[0]);
    return static_cast<double>(bits >> 11) * (1.0 / 9007199254740992.0);
}

struct WarpDraws { // the substream as the prelude sees it: draws 0..31 live one per lane
    double mine;
    unsigned long long seed;
    unsigned generation, index, slot;
    __device__ __forceinline__ double next()
    {
        const unsigned k = slot++;
        if (k < 32u) return __shfl_sync(0xffffffffu, mine, static_cast<int>(k)); // k is warp-uniform
        return philox_draw_at(seed, generation, index, k);
    }
};

constexpr unsigned kTrialWarps = 8;

// the trial vector of individual i, gene j = lane + 32 c written to out[j]; every lane returns the head (F, CR, variant)
__device__ __forceinline__ TrialHead trial_warp_body(const TrialParams &P, unsigned i, unsigned generation, unsigned lane, double *out)
{
    const unsigned dim = P.dim;
    WarpDraws rs{philox_draw_at(P.seed, generation, i, lane), P.seed, generation, i, 0u};
    const TrialHead H = trial_head(P, i, rs);
    const unsigned n0 = uint_below_from(rs, dim); // c_idx(m_e)
    const unsigned c0 = rs.slot;                  // first crossover draw
    unsigned run = 0;                             // exponential: genes n0 .. n0 + run - 1 (mod dim) are mutated
    unsigned used;                                // crossover draws consumed
    if (H.expo) {
        // do { mutate; ++L; } while (draw < CR && L < dim): iteration L = 1, 2, ... consumes draw c0 + L - 1 and is the last one
        // when that draw is >= CR or L == dim
        run = dim;
        for (unsigned b = 0; b < dim; b += 32u) {
            const unsigned L = b + lane + 1u;
            const bool stop = L <= dim && (L == dim || !(philox_draw_at(P.seed, generation, i, c0 + L - 1u) < H.CR));
            const unsigned m = __ballot_sync(0xffffffffu, stop);
            if (m) {
                run = b + static_cast<unsigned>(__ffs(static_cast<int>(m)));
                break;
            }
        }
        used = run;
    } else {
        used = dim;
    }
    const double *xi = P.popold + static_cast<size_t>(i) * dim;
    unsigned resampled = 0; // out-of-bounds genes before this chunk
    for (unsigned b = 0; b < dim; b += 32u) {
        const unsigned j = b + lane;
        bool oob = false;
        double v = 0.0, lo = 0.0, hi = 0.0;
        if (j < dim) {
            const unsigned L = j >= n0 ? j - n0 : j + dim - n0; // position of gene j in the visiting order n0, n0 + 1, ...
            const bool take = H.expo ? (L < run) : (L + 1u == dim || philox_draw_at(P.seed, generation, i, c0 + L) < H.CR);
            v = xi[j];
            if (take) {
                double p[7];
#pragma unroll
                for (int k = 0; k < 7; ++k) p[k] = P.popold[static_cast<size_t>(H.r[k]) * dim + j];
                v = mutate(P.cfg.algo, H.base, v, P.gbIter[j], p, v, H.F);
            }
            lo = P.lb[j];
            hi = P.ub[j];
            oob = (v < lo) || (v > hi);
        }
        // feasibility: out-of-bounds genes are resampled uniformly, de1220.cpp:507-513 / force_bounds_random generic.hpp:403-412
        const unsigned m = __ballot_sync(0xffffffffu, oob && lo != hi);
        if (oob) { // (a gene with lb == ub is reset without a draw)
            const unsigned k = c0 + used + resampled + static_cast<unsigned>(__popc(m & ((1u << lane) - 1u)));
            v = (lo == hi) ? lo : (hi - lo) * philox_draw_at(P.seed, generation, i, k) + lo;
        }
        resampled += static_cast<unsigned>(__popc(m));
        if (j < dim) out[j] = v;
    }
    return H;
}

__global__ void __launch_bounds__(kTrialWarps * 32) de_trial_warp_kernel(const TrialParams P)
{
    const unsigned i = blockIdx.x * kTrialWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (i >= P.NP || *P.stopped) return; // warp-uniform
    const TrialHead H = trial_warp_body(P, i, *P.gen_base + *P.gens_done, lane, P.trial + static_cast<size_t>(i) * P.dim);
    if (lane == 0) {
        if (P.F_out) {
            P.F_out[i] = H.F;
            P.CR_out[i] = H.CR;
        }
        if (P.variant_out) P.variant_out[i] = H.variant;
    }
}

// selection, de.cpp:281-299 / de1220.cpp:515-536: one warp per individual - lane 0 decides, the lanes copy the accepted trial
__global__ void de_select_kernel(const double *trial, const double *ftrial, double *x, double *f, unsigned NP, unsigned dim,
                                 unsigned char *accepted, const double *F_try, const double *CR_try, const unsigned *var_try,
                                 double *F, double *CR, unsigned *variant, const unsigned *stopped)
{
    const unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= NP || *stopped) return;
    const bool ok = ftrial[i] <= f[i];
    __syncwarp();
    if (ok) {
        for (unsigned d = lane; d < dim; d += 32) x[static_cast<size_t>(i) * dim + d] = trial[static_cast<size_t>(i) * dim + d];
    }
    if (lane == 0) {
        accepted[i] = ok;
        if (ok) {
            f[i] = ftrial[i];
            if (F) {
                F[i] = F_try[i];
                CR[i] = CR_try[i];
            }
            if (variant) variant[i] = var_try[i];
        }
    }
}

// Large populations: the same selection as two element-parallel kernels (a warp per individual wastes most lanes on short rows)
__global__ void de_accept_kernel(const double *ftrial, const double *f, unsigned NP, unsigned char *accepted, const double *F_try,
                                 const double *CR_try, const unsigned *var_try, double *F, double *CR, unsigned *variant, const unsigned *stopped)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NP || *stopped) return;
    const bool ok = ftrial[i] <= f[i];
    accepted[i] = ok;
    if (ok) {
        if (F) {
            F[i] = F_try[i];
            CR[i] = CR_try[i];
        }
        if (variant) variant[i] = var_try[i];
    }
}
__global__ void de_copy_accepted_kernel(const double *trial, const double *ftrial, double *x, double *f, const unsigned char *accepted, unsigned NP,
                                        unsigned dim, const unsigned *stopped)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (*stopped || e >= static_cast<size_t>(NP) * dim) return;
    const unsigned i = static_cast<unsigned>(e / dim);
    if (accepted[i]) {
        x[e] = trial[e];
        if (e == static_cast<size_t>(i) * dim) f[i] = ftrial[i];
    }
}

struct DeGlobal { // device-side global best + exit-condition data
    double gbfit, gbF, gbCR;
    unsigned gbidx, gbvariant;
    unsigned best_idx, worst_idx;
    double dx, df;
    unsigned stopped;   // set when dx < xtol or df < ftol (de.cpp:302-321): later generations already queued become no-ops
    unsigned gens_done;
    unsigned gen_base;  // generation counter of the first generation of this evolve() call (the Philox substream index)
};

// the reference's log line of a generation (de: gen, fevals, best, dx, df; sade adds F, CR of the best; de1220 also its variant),
// read from the quantities the exit test has just computed.  Nothing is logged for the generation whose exit test fired: the
// reference returns before it gets there (de.cpp:308-321).
__global__ void de_log_kernel(const DeGlobal *G, const double *f, unsigned algo, double gen, double fevals, double *rows, unsigned *count,
                              unsigned max_rows, unsigned row_len)
{
    if (G->stopped) return;
    const unsigned row = *count;
    if (row >= max_rows) return;
    double *o = rows + static_cast<size_t>(row) * row_len;
    unsigned c = 0;
    o[c++] = gen;
    o[c++] = fevals;
    o[c++] = f[G->best_idx];
    if (algo >= 1u) {
        o[c++] = G->gbF;
        o[c++] = G->gbCR;
    }
    if (algo == 2u) o[c++] = static_cast<double>(G->gbvariant);
    o[c++] = G->dx;
    o[c++] = G->df;
    *count = row + 1u;
}

struct DePartial { // per-CTA result of the scan below (large populations: de_global_partial_kernel)
    double fa, fb, fw;
    unsigned ia, ib, iw;
};

// Large populations: the scan over f is spread over many CTAs; de_global_kernel then only combines their partial results.
__global__ void de_global_partial_kernel(const double *f, const unsigned char *accepted, unsigned NP, int init, const unsigned *stopped,
                                         DePartial *out)
{
    __shared__ double sfa[256], sfb[256], sfw[256];
    __shared__ unsigned sia[256], sib[256], siw[256];
    if (!init && *stopped) return;
    const unsigned t = threadIdx.x, kNone = 0xffffffffu;
    double fa = 0, fb = 0, fw = 0;
    unsigned ia = kNone, ib = kNone, iw = kNone;
    const unsigned per = (NP + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * per, hi = min(NP, lo + per);
    for (unsigned i = lo + t; i < hi; i += blockDim.x) {
        const double v = f[i];
        if (ib == kNone || v < fb) { fb = v; ib = i; }
        if (iw == kNone || v > fw) { fw = v; iw = i; }
        if (!init && accepted[i] && (ia == kNone || v <= fa)) { fa = v; ia = i; }
    }
    sfa[t] = fa; sia[t] = ia; sfb[t] = fb; sib[t] = ib; sfw[t] = fw; siw[t] = iw;
    __syncthreads();
    for (unsigned h = blockDim.x >> 1; h > 0; h >>= 1) {
        if (t < h) {
            const unsigned o = t + h;
            if (sib[o] != kNone && (sib[t] == kNone || sfb[o] < sfb[t] || (sfb[o] == sfb[t] && sib[o] < sib[t]))) { sfb[t] = sfb[o]; sib[t] = sib[o]; }
            if (siw[o] != kNone && (siw[t] == kNone || sfw[o] > sfw[t] || (sfw[o] == sfw[t] && siw[o] < siw[t]))) { sfw[t] = sfw[o]; siw[t] = siw[o]; }
            if (sia[o] != kNone && (sia[t] == kNone || sfa[o] < sfa[t] || (sfa[o] == sfa[t] && sia[o] > sia[t]))) { sfa[t] = sfa[o]; sia[t] = sia[o]; }
        }
        __syncthreads();
    }
    if (t == 0) out[blockIdx.x] = DePartial{sfa[0], sfb[0], sfw[0], sia[0], sib[0], siw[0]};
}

// sequential "if accepted and f <= gbfit: gb = i" over ascending i == smallest accepted fitness, last index on ties, if <= gbfit;
// plus pop.best_idx() / worst_idx() (first min / first max) and the exit quantities dx, df (de.cpp:302-316).  Single CTA.
__global__ void de_global_kernel(const double *x, const double *f, const unsigned char *accepted, unsigned NP, unsigned dim,
                                 const double *F, const double *CR, const unsigned *variant, double *gbX, DeGlobal *G, int init, double xtol,
                                 double ftol, const DePartial *partials, unsigned nparts, unsigned first_generation)
{
    __shared__ double sfa[256], sfb[256], sfw[256];
    __shared__ unsigned sia[256], sib[256], siw[256];
    const unsigned t = threadIdx.x;
    if (!init && G->stopped) return; // uniform: written only by this kernel, at the end of an earlier launch
    if (init && t == 0) {
        G->stopped = 0;
        G->gens_done = 0;
        G->gen_base = first_generation;
    }
    const unsigned kNone = 0xffffffffu;
    double fa = 0, fb = 0, fw = 0;
    unsigned ia = kNone, ib = kNone, iw = kNone;
    if (partials) { // combine the per-CTA scans (slices ascend with the CTA index, so the same tie rules apply)
        for (unsigned k = t; k < nparts; k += blockDim.x) {
            const DePartial q = partials[k];
            if (q.ib != kNone && (ib == kNone || q.fb < fb)) { fb = q.fb; ib = q.ib; }
            if (q.iw != kNone && (iw == kNone || q.fw > fw)) { fw = q.fw; iw = q.iw; }
            if (q.ia != kNone && (ia == kNone || q.fa <= fa)) { fa = q.fa; ia = q.ia; }
        }
    } else {
        for (unsigned i = t; i < NP; i += blockDim.x) {
            const double v = f[i];
            if (ib == kNone || v < fb) { fb = v; ib = i; }
            if (iw == kNone || v > fw) { fw = v; iw = i; }
            if (!init && accepted[i] && (ia == kNone || v <= fa)) { fa = v; ia = i; }
        }
    }
    sfa[t] = fa; sia[t] = ia; sfb[t] = fb; sib[t] = ib; sfw[t] = fw; siw[t] = iw;
    __syncthreads();
    // tree reduction with the reference's tie rules: best = first minimum, worst = first maximum, accepted = LAST minimum
    for (unsigned h = blockDim.x >> 1; h > 0; h >>= 1) {
        if (t < h) {
            const unsigned o = t + h;
            if (sib[o] != kNone && (sib[t] == kNone || sfb[o] < sfb[t] || (sfb[o] == sfb[t] && sib[o] < sib[t]))) { sfb[t] = sfb[o]; sib[t] = sib[o]; }
            if (siw[o] != kNone && (siw[t] == kNone || sfw[o] > sfw[t] || (sfw[o] == sfw[t] && siw[o] < siw[t]))) { sfw[t] = sfw[o]; siw[t] = siw[o]; }
            if (sia[o] != kNone && (sia[t] == kNone || sfa[o] < sfa[t] || (sfa[o] == sfa[t] && sia[o] > sia[t]))) { sfa[t] = sfa[o]; sia[t] = sia[o]; }
        }
        __syncthreads();
    }
    if (t == 0) {
        fa = sfa[0]; ia = sia[0]; fb = sfb[0]; ib = sib[0]; fw = sfw[0]; iw = siw[0];
        G->best_idx = ib;
        G->worst_idx = iw;
        if (init) {
            G->gbidx = ib;
            G->gbfit = fb;
            G->gbF = F ? F[0] : 0.0;   // "initialization to the 0 ind, will soon be forgotten", de1220.cpp:168-170
            G->gbCR = CR ? CR[0] : 0.0;
            G->gbvariant = variant ? variant[0] : 0u;
        } else if (ia != 0xffffffffu && fa <= G->gbfit) {
            G->gbidx = ia;
            G->gbfit = fa;
            if (F) { G->gbF = F[ia]; G->gbCR = CR[ia]; }
            if (variant) G->gbvariant = variant[ia];
        }
        G->df = fabs(f[iw] - f[ib]);
    }
    __syncthreads();
    // gbX <- x[gbidx]: the individual holding the global best can only be replaced by a trial that is itself <= gbfit, in
    // which case gbidx moved with it, so x[gbidx] always equals the reference's gbX.  dx = sum |x_worst - x_best|.
    const unsigned gi = G->gbidx;
    for (unsigned d = t; d < dim; d += blockDim.x) gbX[d] = x[static_cast<size_t>(gi) * dim + d];
    double part = 0;
    for (unsigned d = t; d < dim; d += blockDim.x) part += fabs(x[static_cast<size_t>(G->worst_idx) * dim + d] - x[static_cast<size_t>(G->best_idx) * dim + d]);
    sfa[t] = part;
    __syncthreads();
    if (t == 0) {
        double s = 0;
        for (unsigned k = 0; k < blockDim.x; ++k) s += sfa[k];
        G->dx = s;
        if (!init) {
            G->gens_done += 1;
            if (s < xtol || G->df < ftol) G->stopped = 1; // de.cpp:308,316
        }
    }
}

// ---- launch-bound populations (NP < 16384): selection + global best + exit quantities in ONE launch ------------------------
// Same rules as de_select_kernel followed by de_global_kernel: the trial replaces its parent when ftrial <= f; best = first
// minimum, worst = first maximum, accepted-best = LAST minimum among the accepted trials; gb moves when that one is <= gbfit.
// A warp per individual does the selection (8 per CTA); every CTA leaves its candidates in `parts`, and the CTA that finishes
// last (ticket counter) combines them and updates the global state - no second launch, no single-CTA copy of the population.
constexpr unsigned kFinishWarps = 8;
constexpr unsigned kFinishMaxNP = 16384;

struct ArgVal {
    double v;
    unsigned i;
};

// MODE 0: smaller value wins, ties -> smaller index; 1: larger value wins, ties -> smaller index; 2: smaller value wins, ties -> larger index
template <int MODE> __device__ __forceinline__ ArgVal arg_combine(ArgVal a, ArgVal b)
{
    const unsigned kNone = 0xffffffffu;
    if (b.i == kNone) return a;
    if (a.i == kNone) return b;
    const bool better = MODE == 1 ? b.v > a.v : b.v < a.v;
    const bool tie = b.v == a.v && (MODE == 2 ? b.i > a.i : b.i < a.i);
    return (better || tie) ? b : a;
}

template <int MODE> __device__ __forceinline__ ArgVal arg_reduce_warp(ArgVal a)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        ArgVal o;
        o.v = __shfl_xor_sync(0xffffffffu, a.v, m);
        o.i = __shfl_xor_sync(0xffffffffu, a.i, m);
        a = arg_combine<MODE>(a, o);
    }
    return a;
}

__global__ void __launch_bounds__(kFinishWarps * 32) de_finish_kernel(const double *__restrict__ trial, const double *__restrict__ ftrial,
                                                                      double *x, double *f, unsigned NP, unsigned dim, const double *F_try,
                                                                      const double *CR_try, const unsigned *var_try, double *F, double *CR,
                                                                      unsigned *variant, double *gbX, DeGlobal *G, double xtol, double ftol,
                                                                      DePartial *parts, unsigned *ticket)
{
    __shared__ ArgVal s_best[kFinishWarps], s_worst[kFinishWarps], s_acc[kFinishWarps];
    __shared__ double s_diff[kFinishWarps * 32];
    __shared__ bool s_last;
    const unsigned t = threadIdx.x, warp = t >> 5, lane = t & 31u, kNone = 0xffffffffu;
    if (G->stopped) return; // uniform over the grid: written only by the last CTA of an earlier launch
    const unsigned i = blockIdx.x * kFinishWarps + warp;
    ArgVal best{0.0, kNone}, worst{0.0, kNone}, acc{0.0, kNone};
    if (i < NP) { // selection, de.cpp:281-299 / de1220.cpp:515-536
        const double ft = ftrial[i];
        double v = f[i];
        const bool ok = ft <= v;
        if (ok) {
            const double *src = trial + static_cast<size_t>(i) * dim;
            double *dst = x + static_cast<size_t>(i) * dim;
            for (unsigned d = lane; d < dim; d += 32u) dst[d] = src[d];
            v = ft;
            if (lane == 0) {
                f[i] = ft;
                if (F) {
                    F[i] = F_try[i];
                    CR[i] = CR_try[i];
                }
                if (variant) variant[i] = var_try[i];
            }
            acc = ArgVal{v, i};
        }
        best = worst = ArgVal{v, i};
    }
    if (lane == 0) {
        s_best[warp] = best;
        s_worst[warp] = worst;
        s_acc[warp] = acc;
    }
    __syncthreads();
    if (t == 0) {
        for (unsigned w = 1; w < kFinishWarps; ++w) {
            best = arg_combine<0>(best, s_best[w]);
            worst = arg_combine<1>(worst, s_worst[w]);
            acc = arg_combine<2>(acc, s_acc[w]);
        }
        parts[blockIdx.x] = DePartial{acc.v, best.v, worst.v, acc.i, best.i, worst.i};
        __threadfence(); // this CTA's rows, fitness and candidates are visible before its ticket is
        s_last = atomicAdd(ticket, 1u) + 1u == gridDim.x;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- the last CTA: every other CTA's writes are visible (read through L2: this SM's L1 never held those lines)
    best = worst = acc = ArgVal{0.0, kNone};
    for (unsigned k = t; k < gridDim.x; k += blockDim.x) {
        const DePartial *q = parts + k;
        best = arg_combine<0>(best, ArgVal{__ldcg(&q->fb), __ldcg(&q->ib)});
        worst = arg_combine<1>(worst, ArgVal{__ldcg(&q->fw), __ldcg(&q->iw)});
        acc = arg_combine<2>(acc, ArgVal{__ldcg(&q->fa), __ldcg(&q->ia)});
    }
    best = arg_reduce_warp<0>(best);
    worst = arg_reduce_warp<1>(worst);
    acc = arg_reduce_warp<2>(acc);
    if (lane == 0) {
        s_best[warp] = best;
        s_worst[warp] = worst;
        s_acc[warp] = acc;
    }
    __syncthreads();
    best = s_best[0];
    worst = s_worst[0];
    acc = s_acc[0];
    for (unsigned w = 1; w < kFinishWarps; ++w) {
        best = arg_combine<0>(best, s_best[w]);
        worst = arg_combine<1>(worst, s_worst[w]);
        acc = arg_combine<2>(acc, s_acc[w]);
    }
    const bool moved = acc.i != kNone && acc.v <= G->gbfit; // every thread reads the old G before thread 0 writes it
    const unsigned gi = moved ? acc.i : G->gbidx;
    __syncthreads();
    if (t == 0) {
        G->best_idx = best.i;
        G->worst_idx = worst.i;
        if (moved) {
            G->gbidx = gi;
            G->gbfit = acc.v;
            if (F) {
                G->gbF = __ldcg(F + gi);
                G->gbCR = __ldcg(CR + gi);
            }
            if (variant) G->gbvariant = __ldcg(variant + gi);
        }
        G->df = fabs(worst.v - best.v);
        *ticket = 0u;
    }
    // gbX <- x[gbidx] (see de_global_kernel); dx = sum_d |x_worst[d] - x_best[d]| in ascending d (de.cpp:302-306)
    for (unsigned d = t; d < dim; d += blockDim.x) gbX[d] = __ldcg(x + static_cast<size_t>(gi) * dim + d);
    double dx = 0.0;
    for (unsigned d0 = 0; d0 < dim; d0 += blockDim.x) {
        const unsigned d = d0 + t;
        if (d < dim) s_diff[t] = fabs(__ldcg(x + static_cast<size_t>(worst.i) * dim + d) - __ldcg(x + static_cast<size_t>(best.i) * dim + d));
        __syncthreads();
        if (t == 0)
            for (unsigned k = 0; k < min(blockDim.x, dim - d0); ++k) dx += s_diff[k];
        __syncthreads();
    }
    if (t == 0) {
        G->dx = dx;
        G->gens_done += 1;
        if (dx < xtol || G->df < ftol) G->stopped = 1; // de.cpp:308,316
    }
}

// ---- RESIDENT generation loop (launch-bound populations of the simple UDPs): every generation of an evolve() call inside ONE
// cooperative launch.  A warp builds the trial vector of its individual in shared memory, evaluates it there (the terms of
// simple_device.cuh in parallel, folded by one lane in the reference's order: the same bits as eval_simple.cu), and applies the
// selection into the OTHER copy of the population (x, f, F, CR, variant, accepted are double buffered, so a fast CTA can start
// generation g + 1 while a slow one still reduces generation g).  One grid-wide barrier per generation; after it every CTA
// reduces the whole fitness vector itself (<= 16384 values), so the global-best state needs no second barrier and no broadcast.
struct ResidentParams {
    double *x[2], *f[2], *F[2], *CR[2];
    unsigned *variant[2];
    unsigned char *accepted[2];
    const double *lb, *ub;
    double *gbX;
    DeGlobal *G;
    unsigned NP, dim, gens;
    unsigned long long seed;
    DeConfig cfg;
    double xtol, ftol;
    DePartial *parts[2];      // per-CTA (accepted-best, best, worst) candidates of a generation, double buffered like the population
    unsigned long long *prof; // PGC_DE_PROF=1: clock64 sums of CTA 0 (trial + evaluate + select | grid barrier | reduction), else null
};

#ifndef PGC_RES_WARPS
#define PGC_RES_WARPS 8
#endif
constexpr unsigned kResWarps = PGC_RES_WARPS; // warps (= individuals in flight) per CTA of the resident loop

template <int FAM> __global__ void __launch_bounds__(kResWarps * 32) de_resident_kernel(const ResidentParams R)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double res_smem[]; // per warp: trial row | a-terms | b-terms, dim doubles each
    __shared__ DeGlobal sG;
    __shared__ ArgVal s_best[kResWarps], s_worst[kResWarps], s_acc[kResWarps];
    __shared__ double s_diff[kResWarps * 32];
    const unsigned t = threadIdx.x, warp = t >> 5, lane = t & 31u, kNone = 0xffffffffu;
    const unsigned NP = R.NP, dim = R.dim;
    double *row = res_smem + static_cast<size_t>(warp) * 3 * dim, *ta = row + dim, *tb = ta + dim;
    if (t == 0) sG = *R.G;
    __syncthreads();
    unsigned cur = 0;
    long long c_work = 0, c_sync = 0, c_reduce = 0;
    for (unsigned g = 0; g < R.gens; ++g) {
        const long long c0 = clock64();
        const unsigned nxt = cur ^ 1u;
        TrialParams P{};
        P.popold = R.x[cur];
        P.gbIter = R.x[cur] + static_cast<size_t>(sG.gbidx) * dim; // == gbX: the row of the global best (see de_global_kernel)
        P.lb = R.lb;
        P.ub = R.ub;
        P.F_in = R.F[cur];
        P.CR_in = R.CR[cur];
        P.variant_in = R.variant[cur];
        P.gbIterF = &sG.gbF;
        P.gbIterCR = &sG.gbCR;
        P.NP = NP;
        P.dim = dim;
        P.seed = R.seed;
        P.cfg = R.cfg;
        ArgVal best{0.0, kNone}, worst{0.0, kNone}, acc{0.0, kNone}; // over this warp's individuals, after the selection
        for (unsigned i = blockIdx.x * kResWarps + warp; i < NP; i += gridDim.x * kResWarps) {
            const double fo = R.f[cur][i]; // issued early: its latency hides behind the trial construction
            const TrialHead H = trial_warp_body(P, i, sG.gen_base + g, lane, row);
            __syncwarp();
            for (unsigned j = lane; j < dim; j += 32u) {
                const bool has_next = j + 1u < dim;
                simple::acc_term<FAM>(row[j], static_cast<int>(j), has_next, has_next ? row[j + 1u] : 0.0, ta[j], tb[j]);
            }
            __syncwarp();
            double ft = 0.0;
            if (lane == 0) {
                simple::Acc acc;
                simple::acc_init<FAM>(acc);
                for (unsigned j = 0; j < dim; ++j) simple::acc_fold<FAM>(acc, ta[j], tb[j], j + 1u < dim);
                ft = simple::acc_final<FAM>(acc, static_cast<int>(dim));
            }
            ft = __shfl_sync(0xffffffffu, ft, 0);
            const bool ok = ft <= fo; // selection, de.cpp:281-299 / de1220.cpp:515-536
            const ArgVal mine{ok ? ft : fo, i};
            best = arg_combine<0>(best, mine);
            worst = arg_combine<1>(worst, mine);
            if (ok) acc = arg_combine<2>(acc, mine);
            const double *old_row = R.x[cur] + static_cast<size_t>(i) * dim;
            double *new_row = R.x[nxt] + static_cast<size_t>(i) * dim;
            for (unsigned j = lane; j < dim; j += 32u) new_row[j] = ok ? row[j] : old_row[j];
            if (lane == 0) {
                R.f[nxt][i] = ok ? ft : fo;
                if (R.F[0]) {
                    R.F[nxt][i] = ok ? H.F : R.F[cur][i];
                    R.CR[nxt][i] = ok ? H.CR : R.CR[cur][i];
                }
                if (R.variant[0]) R.variant[nxt][i] = ok ? H.variant : R.variant[cur][i];
            }
            __syncwarp();
        }
        if (lane == 0) {
            s_best[warp] = best;
            s_worst[warp] = worst;
            s_acc[warp] = acc;
        }
        __syncthreads();
        if (t == 0) {
            for (unsigned w = 1; w < kResWarps; ++w) {
                best = arg_combine<0>(best, s_best[w]);
                worst = arg_combine<1>(worst, s_worst[w]);
                acc = arg_combine<2>(acc, s_acc[w]);
            }
            R.parts[nxt][blockIdx.x] = DePartial{acc.v, best.v, worst.v, acc.i, best.i, worst.i};
        }
        const long long c1 = clock64();
        grid.sync();
        const long long c2 = clock64();
        // ---- every CTA combines the CTAs' candidates: best = first minimum, worst = first maximum, accepted-best = LAST minimum
        // among the accepted trials (the rules do not depend on the order of combination)
        best = worst = acc = ArgVal{0.0, kNone};
        for (unsigned k = t; k < gridDim.x; k += blockDim.x) {
            const DePartial q = R.parts[nxt][k];
            best = arg_combine<0>(best, ArgVal{q.fb, q.ib});
            worst = arg_combine<1>(worst, ArgVal{q.fw, q.iw});
            acc = arg_combine<2>(acc, ArgVal{q.fa, q.ia});
        }
        __syncthreads(); // s_best ... are about to be reused
        best = arg_reduce_warp<0>(best);
        worst = arg_reduce_warp<1>(worst);
        acc = arg_reduce_warp<2>(acc);
        if (lane == 0) {
            s_best[warp] = best;
            s_worst[warp] = worst;
            s_acc[warp] = acc;
        }
        __syncthreads();
        best = s_best[0];
        worst = s_worst[0];
        acc = s_acc[0];
        for (unsigned w = 1; w < kResWarps; ++w) {
            best = arg_combine<0>(best, s_best[w]);
            worst = arg_combine<1>(worst, s_worst[w]);
            acc = arg_combine<2>(acc, s_acc[w]);
        }
        // the new global best's self-adapted parameters: loaded now, next to the rows of the exit test
        const bool moved = acc.i != kNone && acc.v <= sG.gbfit; // sG is written again only after the next barrier
        double gF = 0.0, gCR = 0.0;
        unsigned gvar = 0u;
        if (t == 0 && moved) {
            if (R.F[0]) {
                gF = R.F[nxt][acc.i];
                gCR = R.CR[nxt][acc.i];
            }
            if (R.variant[0]) gvar = R.variant[nxt][acc.i];
        }
        double dx = 0.0; // sum_d |x_worst[d] - x_best[d]| in ascending d (de.cpp:302-306)
        for (unsigned d0 = 0; d0 < dim; d0 += blockDim.x) {
            const unsigned d = d0 + t;
            if (d < dim) s_diff[t] = fabs(R.x[nxt][static_cast<size_t>(worst.i) * dim + d] - R.x[nxt][static_cast<size_t>(best.i) * dim + d]);
            __syncthreads();
            if (t == 0)
                for (unsigned k = 0; k < min(blockDim.x, dim - d0); ++k) dx += s_diff[k];
            __syncthreads();
        }
        if (t == 0) {
            sG.best_idx = best.i;
            sG.worst_idx = worst.i;
            if (moved) {
                sG.gbidx = acc.i;
                sG.gbfit = acc.v;
                if (R.F[0]) {
                    sG.gbF = gF;
                    sG.gbCR = gCR;
                }
                if (R.variant[0]) sG.gbvariant = gvar;
            }
            sG.df = fabs(worst.v - best.v);
            sG.dx = dx;
            sG.gens_done += 1;
            if (dx < R.xtol || sG.df < R.ftol) sG.stopped = 1; // de.cpp:308,316
        }
        __syncthreads();
        c_work += c1 - c0;
        c_sync += c2 - c1;
        c_reduce += clock64() - c2;
        cur = nxt;
        if (sG.stopped) break; // the same decision in every CTA
    }
    // the caller's buffers are copy 0
    if (cur == 1u) {
        for (unsigned i = blockIdx.x * kResWarps + warp; i < NP; i += gridDim.x * kResWarps) {
            for (unsigned j = lane; j < dim; j += 32u) R.x[0][static_cast<size_t>(i) * dim + j] = R.x[1][static_cast<size_t>(i) * dim + j];
            if (lane == 0) {
                R.f[0][i] = R.f[1][i];
                if (R.F[0]) {
                    R.F[0][i] = R.F[1][i];
                    R.CR[0][i] = R.CR[1][i];
                }
                if (R.variant[0]) R.variant[0][i] = R.variant[1][i];
            }
        }
    }
    if (blockIdx.x == 0) {
        for (unsigned d = t; d < dim; d += blockDim.x) R.gbX[d] = R.x[cur][static_cast<size_t>(sG.gbidx) * dim + d];
        if (t == 0) *R.G = sG;
        if (t == 0 && R.prof) {
            R.prof[0] = static_cast<unsigned long long>(c_work);
            R.prof[1] = static_cast<unsigned long long>(c_sync);
            R.prof[2] = static_cast<unsigned long long>(c_reduce);
        }
    }
}

__global__ void de_init_adapt_kernel(double *F, double *CR, unsigned *variant, unsigned NP, DeConfig cfg, unsigned long long seed,
                                     unsigned generation)
{ // de1220.cpp:147-165 / sade.cpp:137-156
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NP) return;
    PhiloxStream rs(seed, kTagInit, generation, i);
    if (cfg.variant_adptv == 1u) {
        const double c = rs.next();
        const double f = rs.next();
        CR[i] = c;
        F[i] = f * 0.9 + 0.1;
    } else {
        const double c = normal01(rs);
        const double f = normal01(rs);
        CR[i] = c * 0.15 + 0.5;
        F[i] = f * 0.15 + 0.5;
    }
    if (variant) variant[i] = cfg.allowed[uint_below(rs, cfg.n_allowed)];
}

inline unsigned nblk(size_t n, unsigned t) { return static_cast<unsigned>((n + t - 1) / t); }

// ---- cached workspace + generation graph ---------------------------------------------------------------------------------
constexpr size_t kMaxWorkspaces = 4;            // per problem handle
constexpr unsigned kGraphMaxGenerations = 128;  // generations recorded into one graph
constexpr unsigned kGraphMaxPopulation = 65536; // larger populations are kernel-bound, not launch-bound

bool graphs_enabled()
{
    const char *e = std::getenv("PGC_GRAPHS"); // PGC_GRAPHS=0: plain launches only (tests compare the two paths)
    return !(e && e[0] == '0');
}

struct DeKey {
    unsigned NP, dim;
    DeConfig cfg;
    const void *d_x, *d_f, *d_F, *d_CR, *d_variant, *eval;
    unsigned long long seed;
    double xtol, ftol;
    unsigned config_epoch;
    bool operator==(const DeKey &o) const
    {
        bool same = NP == o.NP && dim == o.dim && cfg.algo == o.cfg.algo && cfg.variant == o.cfg.variant
                    && cfg.variant_adptv == o.cfg.variant_adptv && cfg.F == o.cfg.F && cfg.CR == o.cfg.CR && cfg.n_allowed == o.cfg.n_allowed
                    && d_x == o.d_x && d_f == o.d_f && d_F == o.d_F && d_CR == o.d_CR && d_variant == o.d_variant && eval == o.eval
                    && seed == o.seed && xtol == o.xtol && ftol == o.ftol && config_epoch == o.config_epoch;
        for (unsigned k = 0; same && k < cfg.n_allowed; ++k) same = cfg.allowed[k] == o.cfg.allowed[k];
        return same;
    }
};

struct DeWork : LoopWorkspace {
    DeKey key{};
    int device = 0;
    double *trial = nullptr, *ftrial = nullptr, *gbX = nullptr, *lb = nullptr, *ub = nullptr, *Ftry = nullptr, *CRtry = nullptr, *Fs = nullptr,
           *CRs = nullptr;
    unsigned *vars = nullptr, *vtry = nullptr;
    unsigned char *accepted = nullptr;
    DeGlobal *G = nullptr;
    DePartial *parts = nullptr;
    unsigned *ticket = nullptr; // de_finish_kernel: CTAs done in the current launch
    // resident loop: the second copy of the population state
    double *x2 = nullptr, *f2 = nullptr, *F2 = nullptr, *CR2 = nullptr;
    unsigned *var2 = nullptr;
    unsigned char *acc2[2] = {nullptr, nullptr};
    DePartial *res_parts = nullptr;
    unsigned nparts = 0;
    std::vector<void *> owned;
    cudaGraphExec_t exec = nullptr;
    unsigned graph_gens = 0, graph_launches = 0;
    bool no_graph = false;
    const void *scratch_at_capture = nullptr;
    size_t scratch_bytes_at_capture = 0;

    template <class T>
    int get(T **out, size_t bytes)
    {
        void *p = nullptr;
        PGC_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
        owned.push_back(p);
        *out = static_cast<T *>(p);
        return PGC_OK;
    }
    int allocate(pgc_problem *prob, unsigned algo)
    {
        device = prob->ctx->device;
        const unsigned NP = key.NP, dim = key.dim;
        const size_t nd = static_cast<size_t>(NP) * dim;
        int rc;
        if ((rc = get(&trial, 8 * nd)) || (rc = get(&ftrial, 8 * NP)) || (rc = get(&gbX, 8 * dim)) || (rc = get(&lb, 8 * dim))
            || (rc = get(&ub, 8 * dim)) || (rc = get(&accepted, NP)) || (rc = get(&G, sizeof(DeGlobal))))
            return rc;
        PGC_CUDA(cudaMemcpy(lb, prob->lb.data(), 8 * dim, cudaMemcpyHostToDevice));
        PGC_CUDA(cudaMemcpy(ub, prob->ub.data(), 8 * dim, cudaMemcpyHostToDevice));
        if (algo != 0u) {
            if ((rc = get(&Ftry, 8 * NP)) || (rc = get(&CRtry, 8 * NP))) return rc;
            if (algo == 2u && (rc = get(&vtry, 4 * NP))) return rc;
            if (!key.d_F && (rc = get(&Fs, 8 * NP))) return rc;
            if (!key.d_CR && (rc = get(&CRs, 8 * NP))) return rc;
            if (algo == 2u && !key.d_variant && (rc = get(&vars, 4 * NP))) return rc;
        }
        // large populations: the fitness scan of the global-best kernel is spread over `nparts` CTAs
        nparts = NP >= kFinishMaxNP ? std::min(1024u, NP / 4096u) : 0u;
        if ((rc = get(&parts, sizeof(DePartial) * (nparts ? nparts : nblk(NP, kFinishWarps))))) return rc;
        if ((rc = get(&ticket, sizeof(unsigned)))) return rc;
        PGC_CUDA(cudaMemset(ticket, 0, sizeof(unsigned)));
        return PGC_OK;
    }
    int allocate_resident(unsigned algo)
    {
        if (x2) return PGC_OK;
        const unsigned NP = key.NP, dim = key.dim;
        int rc;
        if ((rc = get(&x2, 8 * static_cast<size_t>(NP) * dim)) || (rc = get(&f2, 8 * NP)) || (rc = get(&acc2[0], NP)) || (rc = get(&acc2[1], NP)))
            return rc;
        if (algo != 0u && ((rc = get(&F2, 8 * NP)) || (rc = get(&CR2, 8 * NP)))) return rc;
        if (algo == 2u && (rc = get(&var2, 4 * NP))) return rc;
        if ((rc = get(&res_parts, 2 * sizeof(DePartial) * ((NP + kResWarps - 1) / kResWarps)))) return rc;
        return PGC_OK;
    }
    void drop_graph()
    {
        if (exec) cudaGraphExecDestroy(exec);
        exec = nullptr;
        graph_gens = graph_launches = 0;
    }
    ~DeWork() override
    {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(device);
        drop_graph();
        for (void *p : owned) cudaFree(p);
        cudaSetDevice(cur);
    }
};

} // namespace

// The self-adaptation state a first evolve() with memory draws (sade.cpp:137-156, de1220.cpp:147-165): F / CR per individual,
// and de1220's mutation variant.  algo: 1 sade, 2 de1220.
int de_init_adaptation_device(unsigned NP, unsigned algo, unsigned variant_adptv, const unsigned *allowed, unsigned n_allowed,
                              unsigned long long seed, unsigned generation, double *d_F, double *d_CR, unsigned *d_variant, cudaStream_t st)
{
    PGC_REQUIRE(algo == 1u || algo == 2u, "de adaptation state: only sade (1) and de1220 (2) keep one");
    PGC_REQUIRE(variant_adptv >= 1u && variant_adptv <= 2u, "The variant for self-adaptation must be in [1,2], while a value of %u was detected.",
                variant_adptv);
    PGC_REQUIRE(d_F && d_CR && (algo == 1u || d_variant), "de adaptation state: null array");
    DeConfig cfg{};
    cfg.algo = algo;
    cfg.variant_adptv = variant_adptv;
    if (algo == 2u) {
        PGC_REQUIRE(allowed && n_allowed >= 1u && n_allowed <= 18u, "de1220 needs between 1 and 18 allowed mutation variants");
        for (unsigned k = 0; k < n_allowed; ++k) cfg.allowed[k] = allowed[k];
        cfg.n_allowed = n_allowed;
    }
    de_init_adapt_kernel<<<nblk(NP, 128), 128, 0, st>>>(d_F, d_CR, algo == 2u ? d_variant : nullptr, NP, cfg, seed, generation);
    PGC_CUDA(cudaGetLastError());
    return PGC_OK;
}

namespace
{

bool resident_enabled()
{
    const char *e = std::getenv("PGC_DE_RESIDENT"); // PGC_DE_RESIDENT=0: one launch per phase (tests compare the two paths)
    return !(e && e[0] == '0');
}

constexpr unsigned kResidentMaxDim = 512;

template <int FAM> int launch_resident(pgc_ctx *ctx, const ResidentParams &R, cudaStream_t st)
{
    auto kern = de_resident_kernel<FAM>;
    const size_t smem = sizeof(double) * 3 * R.dim * kResWarps;
    if (smem > 48 * 1024) PGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    PGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kResWarps * 32, smem));
    PGC_REQUIRE(per_sm >= 1, "de_evolve: the resident loop does not fit on an SM (dimension %u)", R.dim);
    const unsigned blocks = std::min(nblk(R.NP, kResWarps), static_cast<unsigned>(per_sm * ctx->sm_count));
    void *args[] = {const_cast<ResidentParams *>(&R)};
    PGC_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(kern), dim3(blocks), dim3(kResWarps * 32), args, smem, st));
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

} // namespace

// gens generations of de / sade / de1220 on a device-resident population (d_x, d_f updated in place).
// d_F / d_CR / d_variant: per-individual self-adaptation state (sade, de1220); nullptr = initialise as the reference does when
// it has no memory.  *gens_done receives the generations actually run (the xtol / ftol exits of de.cpp:302-321).
int de_evolve_device(pgc_problem *prob, double *d_x, double *d_f, unsigned NP, unsigned gens, unsigned algo, unsigned variant,
                     unsigned variant_adptv, double F, double CR, const unsigned *allowed, unsigned n_allowed, double ftol, double xtol,
                     double *d_F, double *d_CR, unsigned *d_variant, unsigned long long seed, unsigned first_generation,
                     unsigned *gens_done, int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st)
{
    pgc_ctx *ctx = prob->ctx;
    const unsigned dim = static_cast<unsigned>(prob->nx);
    if (gens_done) *gens_done = 0;
    PGC_REQUIRE(algo <= 2u, "de_evolve: algo must be 0 (de), 1 (sade) or 2 (de1220)");
    PGC_REQUIRE(prob->nobj == 1, "Multiple objectives detected in %s instance. Differential evolution cannot deal with them", prob->name.c_str());
    if (gens == 0) return PGC_OK;
    const unsigned min_np = algo == 0 ? 5u : 7u; // de.cpp:107-110, sade.cpp:109-112, de1220.cpp:111-114
    PGC_REQUIRE(NP >= min_np, "differential evolution needs at least %u individuals in the population, %u detected", min_np, NP);
    DeConfig cfg{};
    cfg.algo = algo;
    cfg.variant = variant;
    cfg.variant_adptv = variant_adptv;
    cfg.F = F;
    cfg.CR = CR;
    if (algo == 0u) {
        PGC_REQUIRE(variant >= 1u && variant <= 10u, "The Differential Evolution variant must be in [1, .., 10], while a value of %u was detected.", variant);
        PGC_REQUIRE(F >= 0. && F <= 1. && CR >= 0. && CR <= 1., "The F and CR parameters must be in the [0,1] range");
    } else {
        PGC_REQUIRE(variant_adptv >= 1u && variant_adptv <= 2u, "The variant for self-adaptation must be in [1,2], while a value of %u was detected.", variant_adptv);
        if (algo == 1u) PGC_REQUIRE(variant >= 1u && variant <= 18u, "The Differential Evolution mutation variant must be in [1, .., 18], while a value of %u was detected.", variant);
        if (algo == 2u) {
            PGC_REQUIRE(allowed && n_allowed >= 1u && n_allowed <= 18u, "de1220 needs between 1 and 18 allowed mutation variants");
            for (unsigned k = 0; k < n_allowed; ++k) {
                PGC_REQUIRE(allowed[k] >= 1u && allowed[k] <= 18u, "All mutation variants considered must be in [1, .., 18], while a value of %u was detected.", allowed[k]);
                cfg.allowed[k] = allowed[k];
            }
            cfg.n_allowed = n_allowed;
        }
    }
    // ---- workspace: scratch buffers (and, from the second call on, the instantiated graph of a batch of generations) cached on the
    // problem handle, keyed on everything a launch parameter is derived from
    DeKey key{};
    key.NP = NP;
    key.dim = dim;
    key.cfg = cfg;
    key.d_x = d_x;
    key.d_f = d_f;
    key.d_F = d_F;
    key.d_CR = d_CR;
    key.d_variant = d_variant;
    key.eval = reinterpret_cast<const void *>(eval);
    key.seed = seed;
    key.xtol = xtol;
    key.ftol = ftol;
    key.config_epoch = prob->config_epoch;
    DeWork *W = nullptr;
    {
        std::lock_guard<std::mutex> lock(prob->work_mu);
        for (auto &w : prob->work) {
            auto *dw = dynamic_cast<DeWork *>(w.get());
            if (dw && dw->key == key) W = dw;
        }
        if (!W) {
            if (prob->work.size() >= kMaxWorkspaces) { // evict the least recently used idle one (its population is gone or resting)
                auto lru = prob->work.end();
                for (auto it = prob->work.begin(); it != prob->work.end(); ++it)
                    if ((*it)->users == 0 && (lru == prob->work.end() || (*it)->last_use < (*lru)->last_use)) lru = it;
                if (lru != prob->work.end()) { // (every workspace busy: other threads evolve other populations of this problem - grow)
                    if (int rc = ctx_sync(ctx)) return rc;
                    PGC_CUDA(cudaStreamSynchronize(st));
                    prob->work.erase(lru);
                }
            }
            auto fresh_ws = std::make_unique<DeWork>();
            fresh_ws->key = key;
            int rc = fresh_ws->allocate(prob, algo);
            if (rc != PGC_OK) return rc;
            W = fresh_ws.get();
            prob->work.push_back(std::move(fresh_ws));
        }
        W->last_use = ++prob->work_clock;
        ++W->users;
    }
    struct Release { // the workspace may be evicted again once this call has returned
        pgc_problem *p;
        DeWork *w;
        ~Release()
        {
            std::lock_guard<std::mutex> lock(p->work_mu);
            --w->users;
        }
    } release{prob, W};
    double *trial = W->trial, *ftrial = W->ftrial, *gbX = W->gbX, *lb = W->lb, *ub = W->ub, *Ftry = W->Ftry, *CRtry = W->CRtry;
    double *Fs = d_F ? d_F : W->Fs, *CRs = d_CR ? d_CR : W->CRs;
    unsigned *vars = d_variant ? d_variant : W->vars, *vtry = W->vtry;
    unsigned char *accepted = W->accepted;
    DeGlobal *G = W->G;
    DePartial *parts = W->parts;
    const unsigned nparts = W->nparts;
    const size_t nd = static_cast<size_t>(NP) * dim;
    if (algo != 0u && !(d_F && d_CR && (algo == 1u || d_variant))) // no memory: initialise as the reference does at every evolve()
        de_init_adapt_kernel<<<nblk(NP, 128), 128, 0, st>>>(Fs, CRs, algo == 2u ? vars : nullptr, NP, cfg, seed, first_generation);
    if (nparts) de_global_partial_kernel<<<nparts, 256, 0, st>>>(d_f, nullptr, NP, 1, &G->stopped, parts);
    de_global_kernel<<<1, 256, 0, st>>>(d_x, d_f, nullptr, NP, dim, Fs, CRs, algo == 2u ? vars : nullptr, gbX, G, 1, xtol, ftol,
                                        nparts ? parts : nullptr, nparts, first_generation);
    // ---- launch-bound populations of the simple UDPs: all generations in one cooperative launch (de_resident_kernel)
    const int fam = prob->desc.family;
    const bool simple_family = !prob->inner
                               && (fam == PGC_RASTRIGIN || fam == PGC_ACKLEY || fam == PGC_GRIEWANK || fam == PGC_SCHWEFEL || fam == PGC_ROSENBROCK);
    const bool logging = tls_log != nullptr; // log lines are appended between generations: one launch set per generation
    if (!logging && simple_family && eval == &problem_eval_device && NP < kFinishMaxNP && dim <= kResidentMaxDim && resident_enabled()) {
        if (int rc = W->allocate_resident(algo)) return rc;
        ResidentParams R{};
        R.x[0] = d_x; R.x[1] = W->x2;
        R.f[0] = d_f; R.f[1] = W->f2;
        R.F[0] = algo ? Fs : nullptr; R.F[1] = algo ? W->F2 : nullptr;
        R.CR[0] = algo ? CRs : nullptr; R.CR[1] = algo ? W->CR2 : nullptr;
        R.variant[0] = algo == 2u ? vars : nullptr; R.variant[1] = algo == 2u ? W->var2 : nullptr;
        R.accepted[0] = W->acc2[0]; R.accepted[1] = W->acc2[1];
        R.lb = lb; R.ub = ub;
        R.gbX = gbX;
        R.G = G;
        R.NP = NP; R.dim = dim; R.gens = gens;
        R.seed = seed;
        R.cfg = cfg;
        R.xtol = xtol; R.ftol = ftol;
        R.parts[0] = W->res_parts;
        R.parts[1] = W->res_parts + nblk(NP, kResWarps);
        const char *prof_env = std::getenv("PGC_DE_PROF");
        unsigned long long *d_prof = nullptr;
        if (prof_env && prof_env[0] == '1') PGC_CUDA(cudaMalloc(&d_prof, 3 * sizeof(unsigned long long)));
        R.prof = d_prof;
        int rc;
        switch (fam) {
            case PGC_RASTRIGIN: rc = launch_resident<PGC_RASTRIGIN>(ctx, R, st); break;
            case PGC_ACKLEY: rc = launch_resident<PGC_ACKLEY>(ctx, R, st); break;
            case PGC_GRIEWANK: rc = launch_resident<PGC_GRIEWANK>(ctx, R, st); break;
            case PGC_SCHWEFEL: rc = launch_resident<PGC_SCHWEFEL>(ctx, R, st); break;
            default: rc = launch_resident<PGC_ROSENBROCK>(ctx, R, st); break;
        }
        if (rc != PGC_OK) return rc;
        DeGlobal hres{};
        PGC_CUDA(cudaMemcpyAsync(&hres, G, sizeof(DeGlobal), cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        if (gens_done) *gens_done = hres.gens_done;
        if (d_prof) {
            unsigned long long hp[3] = {};
            cudaMemcpy(hp, d_prof, sizeof(hp), cudaMemcpyDeviceToHost);
            cudaFree(d_prof);
            const double g = hres.gens_done ? hres.gens_done : 1;
            std::fprintf(stderr, "[pgc de_resident] cycles per generation in CTA 0: work %.0f, grid barrier %.0f, reduction %.0f\n", hp[0] / g,
                         hp[1] / g, hp[2] / g);
        }
        return PGC_OK;
    }
    // one generation: no launch parameter depends on the generation index (the kernels read it from G), so the same launches can
    // be replayed from a graph
    auto generation = [&]() -> int {
        TrialParams tp{d_x, gbX, lb, ub, Fs, CRs, algo == 2u ? vars : nullptr, &G->gbF, &G->gbCR, &G->stopped, trial, Ftry, CRtry, vtry, NP, dim,
                       seed, &G->gen_base, &G->gens_done, cfg};
        // launch-bound populations and long rows: a warp per individual (coalesced rows, substream read by position)
        if (dim >= 16u || NP < 65536u) de_trial_warp_kernel<<<nblk(NP, kTrialWarps), kTrialWarps * 32, 0, st>>>(tp);
        else de_trial_kernel<<<nblk(NP, 64), 64, 0, st>>>(tp);
        if (int rc = eval(prob, trial, NP, ftrial, st)) return rc;
        unsigned launched = 2;
        if (!nparts) { // NP < 16384: selection, global best and exit quantities in one launch
            de_finish_kernel<<<nblk(NP, kFinishWarps), kFinishWarps * 32, 0, st>>>(trial, ftrial, d_x, d_f, NP, dim, Ftry, CRtry, vtry,
                                                                                   algo ? Fs : nullptr, algo ? CRs : nullptr,
                                                                                   algo == 2u ? vars : nullptr, gbX, G, xtol, ftol, parts,
                                                                                   W->ticket);
        } else {
            if (dim < 32u) { // large population of short rows: element-parallel selection
                de_accept_kernel<<<nblk(NP, 256), 256, 0, st>>>(ftrial, d_f, NP, accepted, Ftry, CRtry, vtry, algo ? Fs : nullptr,
                                                                algo ? CRs : nullptr, algo == 2u ? vars : nullptr, &G->stopped);
                de_copy_accepted_kernel<<<nblk(nd, 256), 256, 0, st>>>(trial, ftrial, d_x, d_f, accepted, NP, dim, &G->stopped);
                ++launched;
            } else {
                de_select_kernel<<<nblk(static_cast<size_t>(NP) * 32, 256), 256, 0, st>>>(trial, ftrial, d_x, d_f, NP, dim, accepted, Ftry, CRtry,
                                                                                          vtry, algo ? Fs : nullptr, algo ? CRs : nullptr,
                                                                                          algo == 2u ? vars : nullptr, &G->stopped);
            }
            de_global_partial_kernel<<<nparts, 256, 0, st>>>(d_f, accepted, NP, 0, &G->stopped, parts);
            de_global_kernel<<<1, 256, 0, st>>>(d_x, d_f, accepted, NP, dim, algo ? Fs : nullptr, algo ? CRs : nullptr,
                                                algo == 2u ? vars : nullptr, gbX, G, 0, xtol, ftol, parts, nparts, 0u);
            launched += 2;
        }
        ctx->launches.fetch_add(launched, std::memory_order_relaxed);
        return PGC_OK;
    };
    // The exit conditions live on the device (DeGlobal::stopped): once one fires, the generations already queued return
    // immediately, and the host only looks every kPoll generations - a sync per generation made small populations
    // latency-bound on the host round trip.
    constexpr unsigned kPoll = 16;
    unsigned done = 0, g = 0;
    DeGlobal h{};
    int rc = PGC_OK;
    // (a) replay the cached graph while whole batches remain.  Launch-bound populations only: the graph removes the per-launch
    // host cost and the gaps between dependent kernels, which is all a generation of a small population consists of.
    const bool scratch_same = W->scratch_at_capture == ctx->scratch && W->scratch_bytes_at_capture == ctx->scratch_bytes;
    if (W->exec && !scratch_same) W->drop_graph();
    while (!logging && W->exec && !h.stopped && gens - g >= W->graph_gens) {
        PGC_CUDA(cudaGraphLaunch(W->exec, st));
        ctx->launches.fetch_add(W->graph_launches, std::memory_order_relaxed);
        g += W->graph_gens;
        PGC_CUDA(cudaMemcpyAsync(&h, G, sizeof(DeGlobal), cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        done = h.gens_done;
    }
    // (b) plain launches for the rest (and for the whole first call on a new workspace)
    for (; g < gens && !h.stopped; ++g) {
        if ((rc = generation())) return rc;
        if (log_due(g + 1u)) { // de.cpp:324-347, sade.cpp:556-580, de1220.cpp:570-595
            de_log_kernel<<<1, 1, 0, st>>>(G, d_f, algo, static_cast<double>(g + 1u), static_cast<double>(g + 1u) * static_cast<double>(NP),
                                           tls_log->d_rows, tls_log->d_count, tls_log->max_rows, tls_log->row_len);
            ctx->launches.fetch_add(1, std::memory_order_relaxed);
        }
        if ((g + 1) % kPoll == 0 || g + 1 == gens) {
            PGC_CUDA(cudaMemcpyAsync(&h, G, sizeof(DeGlobal), cudaMemcpyDeviceToHost, st));
            PGC_CUDA(cudaStreamSynchronize(st));
            done = h.gens_done;
        }
    }
    if (gens_done) *gens_done = done;
    PGC_CUDA(cudaGetLastError());
    PGC_CUDA(cudaStreamSynchronize(st));
    // (c) first call on this workspace: record a batch of generations for the next call.  Every buffer the launches touch is now
    // allocated (the evaluators size the context's scratch area on first use), so the capture sees no allocation.
    if (!W->exec && !W->no_graph && graphs_enabled() && NP <= kGraphMaxPopulation) {
        const unsigned batch = std::min(gens, kGraphMaxGenerations);
        const unsigned long long l0 = ctx->launches.load(std::memory_order_relaxed);
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess;
        if (ok) {
            for (unsigned b = 0; b < batch && ok; ++b) ok = generation() == PGC_OK;
            ok = (cudaStreamEndCapture(st, &graph) == cudaSuccess) && ok && graph;
        }
        const unsigned long long recorded = ctx->launches.load(std::memory_order_relaxed) - l0;
        ctx->launches.fetch_sub(recorded, std::memory_order_relaxed); // recorded, not run
        if (ok) ok = cudaGraphInstantiate(&W->exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        if (ok) {
            W->graph_gens = batch;
            W->graph_launches = static_cast<unsigned>(recorded);
            W->scratch_at_capture = ctx->scratch;
            W->scratch_bytes_at_capture = ctx->scratch_bytes;
        } else {
            (void)cudaGetLastError(); // an evaluator that cannot be captured keeps the plain path
            W->exec = nullptr;
            W->no_graph = true;
        }
    }
    return PGC_OK;
}

} // namespace pgc
