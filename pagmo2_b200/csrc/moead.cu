// moead.cu - moead_gen (the reference's generational MOEA/D) on the device.
//
// Reference: src/algorithms/moead_gen.cpp:128-345 (evolve), :395-426 (select_parents); decompose_objectives
// (src/utils/multi_objective.cpp:582-638), polynomial_mutation_impl (src/utils/genetic_operators.cpp:148-197).
// moead_gen builds all NP candidates of a generation from the population the previous generation left, evaluates them as ONE batch
// (the bfe branch, :270-288) and only then inserts them one after the other.  On the device a generation is
//   1. the order of the generation: stable argsort of the Philox keys (seed, kTagMoeadOrder, generation, 0, slot = i) (the reference
//      shuffles one persistent index vector with std::shuffle; the restated oracle uses the same keys outside its mt19937 pin mode);
//   2. moead_candidate_kernel, one thread per individual n (Philox substream (seed, kTagMoead, generation, n), read in the
//      reference's order): diversity draw, two distinct parents (neighbourhood or whole population), DE/rand/1 with binomial
//      crossover and the reference's bound repair, polynomial mutation with p_m = 1 / dim;
//   3. one batch evaluation;
//   4. moead_insert_kernel, ONE CTA walking the candidates in the generation's order (the insertion is sequential by definition:
//      a replaced sub-problem is seen by every later candidate, and so is the ideal point): ideal point update, replacement of the
//      own sub-problem, then of the sub-problems of the shuffled neighbourhood (or of everybody) - at most `limit` replacements in
//      all when diversity is preserved.  The shuffle is the order of the keys (seed, kTagMoeadInsert, generation, position, slot =
//      element); all comparisons of a step are independent (the picks are distinct), so they are evaluated in parallel and the
//      `limit` cut takes the successes with the smallest keys.
// Weight vectors and neighbourhoods are inputs (pagmo::decomposition_weights / kNN are utilities outside evolve(); the C++ adapter
// calls pagmo's own).
#include <cmath>
#include <cstdlib>
#include <vector>

#include <cub/cub.cuh>

#include "pgc_internal.cuh"
#include "philox.cuh"

namespace pgc
{

namespace
{

inline unsigned nblk(size_t n, unsigned t) { return static_cast<unsigned>((n + t - 1) / t); }

struct Scratch { // stream-ordered scratch from the (warm) device memory pool
    cudaStream_t st;
    std::vector<void *> owned;
    explicit Scratch(cudaStream_t s) : st(s) {}
    ~Scratch()
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
        int rc = alloc_bytes(&p, sizeof(T) * (count ? count : 1));
        *out = static_cast<T *>(p);
        return rc;
    }
};

// decompose_objectives, multi_objective.cpp:582-638; method 0 weighted, 1 tchebycheff, 2 bi (pow(v, 2) == v * v exactly)
__device__ double decompose(const double *f, unsigned m, const double *weight, const double *ref_point, int method)
{
    double fd = 0.;
    if (method == 0) {
        for (unsigned i = 0; i < m; ++i) fd += weight[i] * f[i];
    } else if (method == 1) {
        for (unsigned i = 0; i < m; ++i) {
            const double fixed_weight = (weight[i] == 0.) ? 1e-4 : weight[i];
            const double tmp = fixed_weight * fabs(f[i] - ref_point[i]);
            if (tmp > fd) fd = tmp;
        }
    } else {
        const double THETA = 5.;
        double d1 = 0., weight_norm = 0., d2 = 0.;
        for (unsigned i = 0; i < m; ++i) {
            d1 += (f[i] - ref_point[i]) * weight[i];
            weight_norm += weight[i] * weight[i];
        }
        weight_norm = sqrt(weight_norm);
        d1 = d1 / weight_norm;
        for (unsigned i = 0; i < m; ++i) {
            const double t = f[i] - (ref_point[i] + d1 * weight[i] / weight_norm);
            d2 += t * t;
        }
        d2 = sqrt(d2);
        fd = d1 + THETA * d2;
    }
    return fd;
}

__global__ void moead_order_keys_kernel(unsigned n, unsigned long long seed, unsigned generation, unsigned long long *keys, unsigned *idx)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = philox_u64(seed, kTagMoeadOrder, generation, 0u, i);
    idx[i] = i;
}

struct CandParams {
    const double *x, *lb, *ub;
    const unsigned *order, *neigh;
    double *cand;
    unsigned char *whole;
    unsigned NP, dim, T;
    double CR, F, eta_m, realb;
    int preserve_diversity;
    unsigned long long seed;
    unsigned generation;
};

// :227-269, one thread per position q of the generation's order
__global__ void moead_candidate_kernel(const CandParams P)
{
    const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.NP) return;
    const unsigned n = P.order[q], NP = P.NP, dim = P.dim;
    PhiloxStream rs(P.seed, kTagMoead, P.generation, n);
    const double u = rs.next(); // the draw is taken whatever preserve_diversity says, :233
    const bool whole = !(u < P.realb || !P.preserve_diversity);
    P.whole[q] = whole;
    unsigned parents[2], np_ = 0;
    while (np_ < 2u) { // select_parents, :395-426
        unsigned r = static_cast<unsigned>(rs.next() * static_cast<double>(NP));
        if (r >= NP) r = NP - 1u;
        const unsigned p = whole ? r : P.neigh[static_cast<size_t>(n) * P.T + r % P.T];
        if (np_ == 1u && parents[0] == p) continue;
        parents[np_++] = p;
    }
    double *c = P.cand + static_cast<size_t>(q) * dim;
    const double *xn = P.x + static_cast<size_t>(n) * dim, *x0 = P.x + static_cast<size_t>(parents[0]) * dim,
                 *x1 = P.x + static_cast<size_t>(parents[1]) * dim;
    for (unsigned kk = 0; kk < dim; ++kk) { // DE/rand/1 + binomial crossover + bound repair, :243-259
        double v;
        if (rs.next() < P.CR) {
            v = xn[kk] + P.F * (x0[kk] - x1[kk]);
            if (v < P.lb[kk]) v = P.lb[kk] + rs.next() * (xn[kk] - P.lb[kk]);
            if (v > P.ub[kk]) v = P.ub[kk] - rs.next() * (P.ub[kk] - xn[kk]);
        } else {
            v = xn[kk];
        }
        c[kk] = v;
    }
    const double p_m = 1.0 / static_cast<double>(dim);
    for (unsigned j = 0; j < dim; ++j) { // polynomial_mutation_impl, genetic_operators.cpp:148-197
        const double yl = P.lb[j], yu = P.ub[j];
        if (rs.next() < p_m && yl != yu) {
            double y = c[j], deltaq, xy, val;
            const double delta1 = (y - yl) / (yu - yl), delta2 = (yu - y) / (yu - yl);
            const double rnd = rs.next(), mut_pow = 1. / (P.eta_m + 1.);
            if (rnd < 0.5) {
                xy = 1. - delta1;
                val = 2. * rnd + (1. - 2. * rnd) * (pow(xy, (P.eta_m + 1.)));
                deltaq = pow(val, mut_pow) - 1.;
            } else {
                xy = 1. - delta2;
                val = 2. * (1. - rnd) + 2. * (rnd - 0.5) * (pow(xy, (P.eta_m + 1.)));
                deltaq = 1. - (pow(val, mut_pow));
            }
            y = y + deltaq * (yu - yl);
            if (y < yl) y = yl;
            if (y > yu) y = yu;
            c[j] = y;
        }
    }
}

// ideal(pop.get_f()), :171: first minimum of every objective under less_than_f; one CTA
__global__ void moead_ideal_kernel(const double *f, unsigned n, unsigned m, double *ideal)
{
    for (unsigned k = threadIdx.x; k < m; k += blockDim.x) {
        unsigned b = 0;
        for (unsigned i = 1; i < n; ++i) {
            const double a = f[static_cast<size_t>(i) * m + k], c = f[static_cast<size_t>(b) * m + k];
            if (!(a != a) && ((c != c) || a < c)) b = i;
        }
        ideal[k] = f[static_cast<size_t>(b) * m + k];
    }
}

struct InsertParams {
    double *x, *f, *ideal;
    const double *cand, *fnew, *weights;
    const unsigned *order, *neigh;
    const unsigned char *whole;
    unsigned NP, dim, m, T, limit;
    int decomposition, preserve_diversity;
    unsigned long long seed;
    unsigned generation;
};

constexpr unsigned kInsertThreads = 1024;
constexpr unsigned kMaxObj = 64;

struct KeyIdx {
    unsigned long long key;
    unsigned k;
};
__device__ __forceinline__ KeyIdx key_min(KeyIdx a, KeyIdx b) // smaller key, ties -> smaller element index; k == ~0u: empty
{
    if (b.k == 0xffffffffu) return a;
    if (a.k == 0xffffffffu) return b;
    return (b.key < a.key || (b.key == a.key && b.k < a.k)) ? b : a;
}

// :297-344, one CTA, the candidates one after the other
__global__ void __launch_bounds__(kInsertThreads) moead_insert_kernel(const InsertParams P)
{
    __shared__ double s_ideal[kMaxObj], s_nf[kMaxObj];
    __shared__ KeyIdx s_red[32];
    __shared__ unsigned s_time, s_take;
    __shared__ unsigned s_sel[kInsertThreads]; // unlimited mode: the picks of a pass that are replaced
    const unsigned t = threadIdx.x, lane = t & 31u, warp = t >> 5, NP = P.NP, dim = P.dim, m = P.m;
    for (unsigned j = t; j < m; j += blockDim.x) s_ideal[j] = P.ideal[j];
    __syncthreads();
    for (unsigned q = 0; q < NP; ++q) {
        const unsigned n = P.order[q];
        const double *nf = P.fnew + static_cast<size_t>(q) * m, *c = P.cand + static_cast<size_t>(q) * dim;
        // 8 - the ideal point, :303-307
        for (unsigned j = t; j < m; j += blockDim.x) {
            const double v = nf[j];
            s_nf[j] = v;
            if (v < s_ideal[j]) s_ideal[j] = v;
        }
        __syncthreads();
        // 9 - the own sub-problem first, :311-316
        if (t == 0) {
            const double f1 = decompose(P.f + static_cast<size_t>(n) * m, m, P.weights + static_cast<size_t>(n) * m, s_ideal, P.decomposition);
            const double f2 = decompose(s_nf, m, P.weights + static_cast<size_t>(n) * m, s_ideal, P.decomposition);
            s_time = (f2 < f1) ? 1u : 0u;
        }
        __syncthreads();
        if (s_time) {
            for (unsigned j = t; j < dim; j += blockDim.x) P.x[static_cast<size_t>(n) * dim + j] = c[j];
            for (unsigned j = t; j < m; j += blockDim.x) P.f[static_cast<size_t>(n) * m + j] = s_nf[j];
        }
        __syncthreads(); // the own row is final before the neighbourhood is compared (it is its own neighbour's neighbour)
        const bool whole = P.whole[q] != 0;
        const unsigned size = whole ? NP : P.T;
        // every element k of the shuffled range: its key and whether the candidate beats the pick's sub-problem
        const bool limited = P.preserve_diversity != 0;
        const unsigned time0 = s_time;
        // passes over the range in chunks of blockDim elements are only needed for the whole-population case
        // (a budget that covers the whole range cannot cut it short: the same as no limit)
        if (!limited || (P.limit > time0 && P.limit - time0 >= size)) { // every success is applied (order irrelevant: the picks are distinct)
            for (unsigned k0 = 0; k0 < size; k0 += blockDim.x) {
                const unsigned k = k0 + t;
                bool win = false;
                unsigned pick = 0;
                if (k < size) {
                    pick = whole ? k : P.neigh[static_cast<size_t>(n) * P.T + k];
                    const double f1 = decompose(P.f + static_cast<size_t>(pick) * m, m, P.weights + static_cast<size_t>(pick) * m, s_ideal, P.decomposition);
                    const double f2 = decompose(s_nf, m, P.weights + static_cast<size_t>(pick) * m, s_ideal, P.decomposition);
                    win = f2 < f1;
                }
                s_sel[t] = win ? pick : 0xffffffffu;
                __syncthreads();
                for (unsigned w = 0; w < min(blockDim.x, size - k0); ++w) { // (uniform loop; rows are copied by the whole CTA)
                    const unsigned p = s_sel[w];
                    if (p == 0xffffffffu) continue;
                    for (unsigned j = t; j < dim; j += blockDim.x) P.x[static_cast<size_t>(p) * dim + j] = c[j];
                    for (unsigned j = t; j < m; j += blockDim.x) P.f[static_cast<size_t>(p) * m + j] = s_nf[j];
                }
                __syncthreads();
            }
            continue;
        }
        // limited: walk the shuffled range until `limit` replacements were made in all (the check follows every element, so with the
        // budget already spent by the own sub-problem exactly one element - the first of the shuffle - is still tried), :339-342
        const unsigned budget = time0 >= P.limit ? 0u : P.limit - time0; // successes still allowed (0: first element only)
        // round r selects, among the elements not selected before, the one with the smallest key that (budget > 0) is a success or
        // (budget == 0) is simply first
        for (unsigned round = 0; round < (budget ? budget : 1u); ++round) {
            KeyIdx best{0ull, 0xffffffffu};
            for (unsigned k = t; k < size; k += blockDim.x) {
                const unsigned pick = whole ? k : P.neigh[static_cast<size_t>(n) * P.T + k];
                bool eligible = true; // (a pick replaced in an earlier round now holds the candidate itself: it no longer wins)
                if (budget) {
                    const double f1 = decompose(P.f + static_cast<size_t>(pick) * m, m, P.weights + static_cast<size_t>(pick) * m, s_ideal, P.decomposition);
                    const double f2 = decompose(s_nf, m, P.weights + static_cast<size_t>(pick) * m, s_ideal, P.decomposition);
                    eligible = f2 < f1;
                }
                if (eligible) best = key_min(best, KeyIdx{philox_u64(P.seed, kTagMoeadInsert, P.generation, q, k), k});
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                KeyIdx o;
                o.key = __shfl_xor_sync(0xffffffffu, best.key, s);
                o.k = __shfl_xor_sync(0xffffffffu, best.k, s);
                best = key_min(best, o);
            }
            if (lane == 0) s_red[warp] = best;
            __syncthreads();
            if (t == 0) {
                KeyIdx b = s_red[0];
                for (unsigned w = 1; w < (blockDim.x >> 5); ++w) b = key_min(b, s_red[w]);
                unsigned take = 0xffffffffu;
                if (b.k != 0xffffffffu) {
                    const unsigned pick = whole ? b.k : P.neigh[static_cast<size_t>(n) * P.T + b.k];
                    bool win = true;
                    if (!budget) { // the one element tried after the budget was spent: replaced only if it wins
                        const double f1 = decompose(P.f + static_cast<size_t>(pick) * m, m, P.weights + static_cast<size_t>(pick) * m, s_ideal, P.decomposition);
                        const double f2 = decompose(s_nf, m, P.weights + static_cast<size_t>(pick) * m, s_ideal, P.decomposition);
                        win = f2 < f1;
                    }
                    if (win) take = pick;
                }
                s_take = take;
            }
            __syncthreads();
            const unsigned p = s_take;
            if (p == 0xffffffffu) break; // no (further) success in the range: uniform
            for (unsigned j = t; j < dim; j += blockDim.x) P.x[static_cast<size_t>(p) * dim + j] = c[j];
            for (unsigned j = t; j < m; j += blockDim.x) P.f[static_cast<size_t>(p) * m + j] = s_nf[j];
            __syncthreads();
        }
        __syncthreads();
    }
    for (unsigned j = t; j < m; j += blockDim.x) P.ideal[j] = s_ideal[j];
}

} // namespace

namespace
{
// moead_gen.cpp:180-211: (gen, fevals, ADF, ideal point); ADF = sum over the population of the decomposed fitness of individual i on
// its own weight vector against the current ideal point
__global__ void moead_log_kernel(const double *f, const double *w, const double *ideal, unsigned NP, unsigned m, int method, double gen,
                                 double fevals, double *rows, unsigned *count, unsigned max_rows, unsigned row_len)
{
    __shared__ double s[256];
    const unsigned r = *count;
    if (r >= max_rows || row_len < 3u + m) return;
    double a = 0.;
    for (unsigned i = threadIdx.x; i < NP; i += blockDim.x) a += decompose(f + static_cast<size_t>(i) * m, m, w + static_cast<size_t>(i) * m, ideal, method);
    s[threadIdx.x] = a;
    __syncthreads();
    for (unsigned k = blockDim.x / 2; k; k >>= 1) {
        if (threadIdx.x < k) s[threadIdx.x] += s[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double *o = rows + static_cast<size_t>(r) * row_len;
        o[0] = gen, o[1] = fevals, o[2] = s[0];
        for (unsigned c = 0; c < m; ++c) o[3u + c] = ideal[c];
        *count = r + 1u;
    }
}
} // namespace

int moead_gen_evolve_device(pgc_problem *prob, double *d_x, double *d_f, unsigned NP, unsigned gens, const double *h_weights,
                            const unsigned *h_neigh, unsigned T, int decomposition, double CR, double F, double eta_m, double realb, unsigned limit,
                            int preserve_diversity, unsigned long long seed, unsigned first_generation,
                            int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st)
{
    pgc_ctx *ctx = prob->ctx;
    const unsigned dim = static_cast<unsigned>(prob->nx), m = static_cast<unsigned>(prob->nobj);
    // the reference constructor's checks (moead_gen.cpp:60-104) and evolve()'s (:140-166)
    PGC_REQUIRE(decomposition >= 0 && decomposition <= 2, "moead_gen: decomposition must be 0 (weighted), 1 (tchebycheff) or 2 (bi)");
    PGC_REQUIRE(CR >= 0. && CR <= 1.,
                "The parameter CR (used by the differential evolution operator) needs to be in [0,1], while a value of %g was detected", CR);
    PGC_REQUIRE(F >= 0. && F <= 1., "The parameter F (used by the differential evolution operator) needs to be in [0,1], while a value of %g was detected",
                F);
    PGC_REQUIRE(eta_m >= 0., "The distribution index for the polynomial mutation (eta_m) needs to be positive, while a value of %g was detected", eta_m);
    PGC_REQUIRE(realb >= 0. && realb <= 1., "The chance of considering a neighbourhood (realb) needs to be in [0,1], while a value of %g was detected",
                realb);
    PGC_REQUIRE(T >= 2u, "The size of the weight's neighborhood needs to be >= 2, while a size of %u was detected", T);
    PGC_REQUIRE(NP >= 1u, "MOEAD-GEN cannot work on an empty population");
    PGC_REQUIRE(m >= 2u, "This is a multiobjective algorithm, while number of objectives detected in %s is %u", prob->name.c_str(), m);
    PGC_REQUIRE(m <= kMaxObj, "moead_gen on the device handles at most %u objectives, %u detected", kMaxObj, m);
    PGC_REQUIRE(T <= NP - 1u, "The neighbourhood size specified (T) is %u: too large for the input population having size %u", T, NP);
    PGC_REQUIRE(h_weights && h_neigh, "moead_gen: the weight vectors and their neighbourhoods are required");
    for (unsigned j = 0; j < dim; ++j)
        PGC_REQUIRE(prob->lb[j] != prob->ub[j],
                    "MOEAD-GEN cannot work on problems having a lower bound equal to an upper bound. Check your bounds.");
    for (size_t e = 0; e < static_cast<size_t>(NP) * T; ++e)
        PGC_REQUIRE(h_neigh[e] < NP, "moead_gen: neighbourhood entry %zu names individual %u of %u", e, h_neigh[e], NP);
    if (gens == 0) return PGC_OK;

    Scratch sc(st);
    double *lb, *ub, *w, *ideal, *cand, *fnew;
    unsigned *neigh, *order, *idx_in;
    unsigned long long *k_in, *k_out;
    unsigned char *whole;
    int rc;
    if ((rc = sc.alloc(&lb, dim)) || (rc = sc.alloc(&ub, dim)) || (rc = sc.alloc(&w, static_cast<size_t>(NP) * m)) || (rc = sc.alloc(&ideal, m))
        || (rc = sc.alloc(&cand, static_cast<size_t>(NP) * dim)) || (rc = sc.alloc(&fnew, static_cast<size_t>(NP) * m))
        || (rc = sc.alloc(&neigh, static_cast<size_t>(NP) * T)) || (rc = sc.alloc(&order, NP)) || (rc = sc.alloc(&idx_in, NP))
        || (rc = sc.alloc(&k_in, NP)) || (rc = sc.alloc(&k_out, NP)) || (rc = sc.alloc(&whole, NP)))
        return rc;
    PGC_CUDA(cudaMemcpyAsync(lb, prob->lb.data(), sizeof(double) * dim, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(ub, prob->ub.data(), sizeof(double) * dim, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(w, h_weights, sizeof(double) * NP * m, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(neigh, h_neigh, sizeof(unsigned) * NP * T, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaStreamSynchronize(st)); // the caller's arrays may go away
    moead_ideal_kernel<<<1, 64, 0, st>>>(d_f, NP, m, ideal);
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0;
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, k_in, k_out, idx_in, order, static_cast<int>(NP), 0, 64, st));
    if ((rc = sc.alloc_bytes(&cub_tmp, cub_bytes))) return rc;
    for (unsigned g = 0; g < gens; ++g) {
        const unsigned generation = first_generation + g;
        if (log_due(g + 1u))
            moead_log_kernel<<<1, 256, 0, st>>>(d_f, w, ideal, NP, m, decomposition, static_cast<double>(g + 1u), static_cast<double>(g) * NP,
                                                tls_log->d_rows, tls_log->d_count, tls_log->max_rows, tls_log->row_len);
        moead_order_keys_kernel<<<nblk(NP, 256), 256, 0, st>>>(NP, seed, generation, k_in, idx_in);
        size_t bytes = cub_bytes;
        PGC_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, bytes, k_in, k_out, idx_in, order, static_cast<int>(NP), 0, 64, st));
        CandParams cp{d_x, lb, ub, order, neigh, cand, whole, NP, dim, T, CR, F, eta_m, realb, preserve_diversity, seed, generation};
        moead_candidate_kernel<<<nblk(NP, 128), 128, 0, st>>>(cp);
        if ((rc = eval(prob, cand, NP, fnew, st))) return rc;
        InsertParams ip{d_x, d_f, ideal, cand, fnew, w, order, neigh, whole, NP, dim, m, T, limit, decomposition, preserve_diversity, seed, generation};
        moead_insert_kernel<<<1, kInsertThreads, 0, st>>>(ip);
        ctx->launches.fetch_add(4, std::memory_order_relaxed);
    }
    PGC_CUDA(cudaGetLastError());
    PGC_CUDA(cudaStreamSynchronize(st));
    return PGC_OK;
}

} // namespace pgc
