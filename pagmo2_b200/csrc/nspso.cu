// nspso.cu - non-dominated sorting particle swarm optimisation on the device (the first of the bfe-caller UDAs).
//
// Reference: src/algorithms/nspso.cpp:84-411 (evolve), :444-520 (minfit, compute_maxmin, euclidian_distance, compute_niche_count).
// nspso is generational in the reference itself: within a generation every particle moves on the state the previous generation
// left - positions, velocities, the archive (m_best_dvs / m_best_fit) and the leaders chosen at the top of the generation - the
// moved swarm is evaluated as ONE batch (the bfe branch, :342-359), and the archive becomes the best N of (moved swarm | archive).
// Per generation on the device:
//   1. fast_non_dominated_sorting of the swarm (fnds_device) and the leaders (:195-291):
//        crowding distance: the first max(|front 0|, 2) entries of sort_population_mo(fit);
//        niche count      : front 0 ordered by the number of front-0 decision vectors closer than the Fonseca-Fleming delta
//                           (all pairs of the front, one warp per member), or {front0[0], front1[0]} for a single-point front;
//        max min          : all particles ordered by max_j min_k (f_i[k] - f_j[k]), the negative ones kept (at least 2);
//      orders are stable sorts on order-preserving keys (detail::less_than_f; the reference's std::sort leaves ties in an
//      unspecified order, the restated oracle uses the same stable order outside its mt19937 pin mode).
//   2. move (:293-340): particle p owns the Philox substream (seed, kTagNspso, generation, p): leader index = floor(u * (ext + 1)),
//      repeated while it names the particle itself, then r1, r2; one thread per coordinate, the three draws recomputed per thread
//      (the rejection loop is a handful of draws).  Arithmetic: the reference's expressions, one rounding each (-fmad=false).
//   3. batch evaluation of the moved swarm.
//   4. archive = rows sort_population_mo(moved | archive)[0 .. N) (or the max-min order), population = moved swarm (:361-395).
// Initial velocities (memory-less start): (seed, kTagInit, generation, p, d), uniform_real_from_range(minv, maxv) (:145-152).
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

__device__ __forceinline__ bool less_f(double a, double b) { return !(a != a) && ((b != b) || a < b); } // detail::less_than_f

__global__ void nspso_init_velocity_kernel(double *V, unsigned n, unsigned dim, const double *minv, const double *maxv, unsigned long long seed,
                                           unsigned generation)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(n) * dim) return;
    const unsigned p = static_cast<unsigned>(e / dim), d = static_cast<unsigned>(e % dim);
    const double lo = minv[d], hi = maxv[d];
    V[e] = (lo == hi) ? lo : philox_u01(seed, kTagInit, generation, p, d) * (hi - lo) + lo; // uniform_real_from_range, generic.hpp:98-104
}

// order-preserving u64 keys of doubles, NaN last, -0 == +0 (less_than_f); idx = 0 .. n-1
__global__ void nspso_keys_kernel(const double *v, unsigned n, unsigned long long *keys, unsigned *idx)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = v[i] + 0.0;
    unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(x));
    b = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
    keys[i] = (x != x) ? 0xffffffffffffffffull : b;
    idx[i] = i;
}

// compute_maxmin, :464-484: maxmin[i] = max_{j != i} min_k (f_i[k] - f_j[k]); one warp per i
__global__ void nspso_maxmin_kernel(const double *f, unsigned n, unsigned m, double *maxmin)
{
    const unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (i >= n) return;
    const double *fi = f + static_cast<size_t>(i) * m;
    auto minfit = [&](unsigned j) { // :444-462
        const double *fj = f + static_cast<size_t>(j) * m;
        double mn = fi[0] - fj[0];
        for (unsigned k = 0; k < m; ++k) {
            const double t = fi[k] - fj[k];
            if (t < mn) mn = t;
        }
        return mn;
    };
    // the reference starts from j = (i + 1) % n and replaces on `tmp > maxmin[i]`: the maximum under operator> (NaN never wins, and a
    // NaN start value is never replaced)
    const double start = minfit((i + 1u) % n);
    double best = start;
    for (unsigned j = lane; j < n; j += 32u)
        if (j != i) {
            const double t = minfit(j);
            if (t > best) best = t;
        }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, best, s);
        if (o > best) best = o;
    }
    if (lane == 0) maxmin[i] = (start != start) ? start : best;
}

// ideal (min over all points) and nadir (max over the first front), multi_objective.cpp:480-541, and the Fonseca-Fleming delta
// (:222-252) -> out[0]; one CTA
__global__ void nspso_delta_kernel(const double *f, unsigned n, unsigned m, const unsigned *front0, unsigned n0, double *out)
{
    __shared__ double s_ideal[64], s_nadir[64];
    for (unsigned k = threadIdx.x; k < m; k += blockDim.x) {
        unsigned bi = 0, wi = 0;
        for (unsigned i = 1; i < n; ++i)
            if (less_f(f[static_cast<size_t>(i) * m + k], f[static_cast<size_t>(bi) * m + k])) bi = i;
        for (unsigned q = 1; q < n0; ++q)
            if (less_f(f[static_cast<size_t>(front0[wi]) * m + k], f[static_cast<size_t>(front0[q]) * m + k])) wi = q;
        s_ideal[k] = f[static_cast<size_t>(bi) * m + k];
        s_nadir[k] = f[static_cast<size_t>(front0[wi]) * m + k];
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double delta = 1.0;
    if (m == 2u) {
        const unsigned dd = n0 == 1u ? 2u : n0;
        delta = ((s_nadir[0] - s_ideal[0]) + (s_nadir[1] - s_ideal[1])) / (static_cast<double>(dd) - 1);
    } else if (m == 3u) {
        const double d1 = s_nadir[0] - s_ideal[0], d2 = s_nadir[1] - s_ideal[1], d3 = s_nadir[2] - s_ideal[2];
        double ns = static_cast<double>(n0);
        if (ns < 2.0) ns = 2.0;
        delta = sqrt(4 * d2 * d1 * ns + 4 * d3 * d1 * ns + 4 * d2 * d3 * ns + d1 * d1 + d2 * d2 + d3 * d3 - 2 * d2 * d1 - 2 * d3 * d1 - 2 * d2 * d3 + d1
                     + d2 + d3)
                / (2 * (ns - 1));
    } else {
        for (unsigned k = 0; k < m; ++k) delta *= s_nadir[k] - s_ideal[k];
        delta = pow(delta, 1.0 / static_cast<double>(m)) / static_cast<double>(n0);
    }
    out[0] = delta;
}

// compute_niche_count, :500-518: count[a] = |{b in front 0 : |x_a - x_b| < delta}| (the point itself counts); one warp per a,
// the lanes take the partners b, each distance summed in the reference's coordinate order
__global__ void nspso_niche_kernel(const double *x, unsigned dim, const unsigned *front0, unsigned n0, const double *delta, double *count)
{
    const unsigned a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (a >= n0) return;
    const double *xa = x + static_cast<size_t>(front0[a]) * dim;
    const double dl = delta[0];
    unsigned cnt = 0;
    for (unsigned b = lane; b < n0; b += 32u) {
        const double *xb = x + static_cast<size_t>(front0[b]) * dim;
        double sum = 0.0;
        for (unsigned j = 0; j < dim; ++j) {
            const double d = xa[j] - xb[j];
            sum += d * d;
        }
        if (sqrt(sum) < dl) ++cnt;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
    if (lane == 0) count[a] = static_cast<double>(cnt);
}

__global__ void nspso_gather_u32_kernel(const unsigned *src, const unsigned *idx, unsigned n, unsigned *dst)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// number of leading entries of the max-min order that are kept: the first one, then every following one with a negative key, at
// least two (:283-291); one thread (the scan stops at the first non-negative key)
__global__ void nspso_maxmin_cut_kernel(const double *maxmin, const unsigned *order, unsigned n, unsigned *nb)
{
    unsigned i = 1;
    for (; i < n && maxmin[order[i]] < 0; ++i) {
    }
    nb[0] = i < 2u ? 2u : i;
}

struct MoveParams {
    const double *x, *best_x, *lb, *ub, *minv, *maxv;
    double *V, *x_new;
    const unsigned *bnd; // the leaders (indices into the archive)
    const unsigned *nb;  // device: how many of them
    unsigned n, dim, leader_selection_range;
    double omega, c1, c2, chi;
    unsigned long long seed;
    unsigned generation;
};

// :293-340, one thread per (particle, coordinate)
__global__ void nspso_move_kernel(const MoveParams P)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(P.n) * P.dim) return;
    const unsigned idx = static_cast<unsigned>(e / P.dim), i = static_cast<unsigned>(e % P.dim);
    int ext = static_cast<int>(ceil(static_cast<double>(P.nb[0]) * static_cast<double>(P.leader_selection_range) / 100.0) - 1);
    if (ext < 1) ext = 1;
    PhiloxStream rs(P.seed, kTagNspso, P.generation, idx);
    unsigned leader_idx;
    do { // uniform_int_distribution<int>(0, ext)
        const unsigned v = static_cast<unsigned>(rs.next() * static_cast<double>(ext + 1));
        leader_idx = v < static_cast<unsigned>(ext + 1) ? v : static_cast<unsigned>(ext);
    } while (P.bnd[leader_idx] == idx);
    const double r1 = rs.next();
    const double r2 = rs.next();
    const double xi = P.x[e], leader = P.best_x[static_cast<size_t>(P.bnd[leader_idx]) * P.dim + i];
    double v = P.omega * P.V[e] + P.c1 * r1 * (P.best_x[e] - xi) + P.c2 * r2 * (leader - xi);
    if (v > P.maxv[i]) v = P.maxv[i];
    else if (v < P.minv[i]) v = P.minv[i];
    double xn = xi + P.chi * v;
    if (xn > P.ub[i]) {
        xn = P.ub[i];
        v = 0.0;
    } else if (xn < P.lb[i]) {
        xn = P.lb[i];
        v = 0.0;
    }
    P.V[e] = v;
    P.x_new[e] = xn;
}

__global__ void nspso_gather_rows_kernel(const double *src, const unsigned *idx, unsigned rows, unsigned width, double *dst)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e < static_cast<size_t>(rows) * width) dst[e] = src[static_cast<size_t>(idx[e / width]) * width + e % width];
}

// stable ascending argsort of n doubles under less_than_f: d_order_out[0..n)
int argsort_less_f(const double *d_v, unsigned n, unsigned *d_order_out, cudaStream_t st)
{
    Scratch sc(st); // freed in stream order when this call returns
    unsigned long long *k_in, *k_out;
    unsigned *i_in;
    int rc;
    if ((rc = sc.alloc(&k_in, n)) || (rc = sc.alloc(&k_out, n)) || (rc = sc.alloc(&i_in, n))) return rc;
    nspso_keys_kernel<<<nblk(n, 256), 256, 0, st>>>(d_v, n, k_in, i_in);
    void *tmp = nullptr;
    size_t bytes = 0;
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k_in, k_out, i_in, d_order_out, static_cast<int>(n), 0, 64, st));
    if ((rc = sc.alloc_bytes(&tmp, bytes))) return rc;
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, k_in, k_out, i_in, d_order_out, static_cast<int>(n), 0, 64, st));
    return PGC_OK;
}

} // namespace

// what a first evolve() with memory sets up (nspso.cpp:127-152): velocities drawn in +-(ub - lb) * v_coeff, archive = the population
int nspso_init_memory_device(pgc_problem *prob, const double *d_x, const double *d_f, unsigned NP, double v_coeff, unsigned long long seed,
                             unsigned generation, double *d_vel, double *d_best_x, double *d_best_f, cudaStream_t st)
{
    const unsigned dim = static_cast<unsigned>(prob->nx), m = static_cast<unsigned>(prob->nobj);
    PGC_REQUIRE(d_vel && d_best_x && d_best_f, "nspso memory: null array");
    PGC_REQUIRE(v_coeff > 0. && v_coeff <= 1., "velocity scaling factor should be in ]0,1] range, while a value of %g was detected", v_coeff);
    Scratch sc(st);
    double *minv, *maxv;
    int rc;
    if ((rc = sc.alloc(&minv, dim)) || (rc = sc.alloc(&maxv, dim))) return rc;
    std::vector<double> h_minv(dim), h_maxv(dim);
    for (unsigned j = 0; j < dim; ++j) {
        const double vwidth = (prob->ub[j] - prob->lb[j]) * v_coeff;
        h_minv[j] = -1. * vwidth;
        h_maxv[j] = vwidth;
    }
    PGC_CUDA(cudaMemcpyAsync(minv, h_minv.data(), sizeof(double) * dim, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(maxv, h_maxv.data(), sizeof(double) * dim, cudaMemcpyHostToDevice, st));
    nspso_init_velocity_kernel<<<nblk(static_cast<size_t>(NP) * dim, 256), 256, 0, st>>>(d_vel, NP, dim, minv, maxv, seed, generation);
    PGC_CUDA(cudaMemcpyAsync(d_best_x, d_x, sizeof(double) * NP * dim, cudaMemcpyDeviceToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(d_best_f, d_f, sizeof(double) * NP * m, cudaMemcpyDeviceToDevice, st));
    PGC_CUDA(cudaGetLastError());
    PGC_CUDA(cudaStreamSynchronize(st));
    return PGC_OK;
}

int nspso_evolve_device(pgc_problem *prob, double *d_x, double *d_f, unsigned NP, unsigned gens, double omega, double c1, double c2, double chi,
                        double v_coeff, unsigned leader_selection_range, unsigned diversity, unsigned long long seed, unsigned first_generation,
                        double *d_vel, double *d_best_x, double *d_best_f,
                        int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st)
{
    pgc_ctx *ctx = prob->ctx;
    const unsigned dim = static_cast<unsigned>(prob->nx), m = static_cast<unsigned>(prob->nobj);
    // the reference constructor's checks (nspso.cpp:50-82) and evolve()'s (:99-119)
    PGC_REQUIRE(omega >= 0. && omega <= 1., "The particles' inertia weight must be in the [0,1] range, while a value of %g was detected", omega);
    PGC_REQUIRE(c1 > 0. && c2 > 0. && chi > 0.,
                "first and second magnitude of the force coefficients and velocity scaling factor should be greater than 0");
    PGC_REQUIRE(v_coeff > 0. && v_coeff <= 1., "velocity scaling factor should be in ]0,1] range, while a value of %g was detected", v_coeff);
    PGC_REQUIRE(leader_selection_range <= 100u, "leader selection range coefficient should be in the ]0,100] range, while a value of %u was detected",
                leader_selection_range);
    PGC_REQUIRE(diversity <= 2u, "Non existing diversity mechanism method.");
    PGC_REQUIRE(m >= 2u, "This is a multi-objective algorithm, while number of objectives detected in %s is %u", prob->name.c_str(), m);
    PGC_REQUIRE(m <= 64u, "nspso on the device handles at most 64 objectives, %u detected", m);
    PGC_REQUIRE(NP >= 2u, "NSPSO can only work with population sizes >=2, whereas %u were detected.", NP);
    PGC_REQUIRE((d_best_x == nullptr) == (d_best_f == nullptr), "nspso: the archive's decision vectors and fitness come together");
    if (gens == 0) return PGC_OK;

    Scratch sc(st);
    const size_t nd = static_cast<size_t>(NP) * dim, nm = static_cast<size_t>(NP) * m;
    double *lb, *ub, *minv, *maxv, *V = d_vel, *bx = d_best_x, *bf = d_best_f, *x2, *f2, *keyv, *delta;
    unsigned *rank, *order, *foff, *sorted, *bnd, *nb, *order2, *sl, *fkey;
    int rc;
    if ((rc = sc.alloc(&lb, dim)) || (rc = sc.alloc(&ub, dim)) || (rc = sc.alloc(&minv, dim)) || (rc = sc.alloc(&maxv, dim))
        || (rc = sc.alloc(&x2, 2 * nd)) || (rc = sc.alloc(&f2, 2 * nm)) || (rc = sc.alloc(&keyv, 2 * static_cast<size_t>(NP)))
        || (rc = sc.alloc(&delta, 1)) || (rc = sc.alloc(&rank, NP)) || (rc = sc.alloc(&order, NP)) || (rc = sc.alloc(&foff, NP + 1))
        || (rc = sc.alloc(&sorted, NP)) || (rc = sc.alloc(&bnd, NP)) || (rc = sc.alloc(&nb, 1)) || (rc = sc.alloc(&order2, 2 * static_cast<size_t>(NP)))
        || (rc = sc.alloc(&sl, NP)) || (rc = sc.alloc(&fkey, NP)))
        return rc;
    if (!V && (rc = sc.alloc(&V, nd))) return rc;
    if (!bx && ((rc = sc.alloc(&bx, nd)) || (rc = sc.alloc(&bf, nm)))) return rc;
    std::vector<double> h_minv(dim), h_maxv(dim);
    for (unsigned j = 0; j < dim; ++j) { // :139-143
        const double vwidth = (prob->ub[j] - prob->lb[j]) * v_coeff;
        h_minv[j] = -1. * vwidth;
        h_maxv[j] = vwidth;
    }
    PGC_CUDA(cudaMemcpyAsync(lb, prob->lb.data(), sizeof(double) * dim, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(ub, prob->ub.data(), sizeof(double) * dim, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(minv, h_minv.data(), sizeof(double) * dim, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(maxv, h_maxv.data(), sizeof(double) * dim, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaStreamSynchronize(st)); // the host vectors go out of scope only at return, but keep the copies unambiguous
    if (!d_vel) { // memory-less start, :145-152
        nspso_init_velocity_kernel<<<nblk(nd, 256), 256, 0, st>>>(V, NP, dim, minv, maxv, seed, first_generation);
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
    }
    if (!d_best_x) { // :127-133
        PGC_CUDA(cudaMemcpyAsync(bx, d_x, sizeof(double) * nd, cudaMemcpyDeviceToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(bf, d_f, sizeof(double) * nm, cudaMemcpyDeviceToDevice, st));
    }
    for (unsigned g = 0; g < gens; ++g) {
        const unsigned generation = first_generation + g;
        if (log_due(g + 1u) && (rc = log_ideal_device(ctx, bf, NP, m, g + 1u, static_cast<unsigned long long>(g) * NP, st))) return rc; // nspso.cpp:163-192
        // 1 - the leaders
        // only the first front is needed (its size; for the niche count its members in the reference's order, and the first member
        // of the second front when it is a single point): the level loop stops after one (two) closed fronts
        unsigned nfronts = 0, h_foff[3] = {0, 0, 0};
        if ((rc = fnds_device(ctx, d_f, NP, m, rank, nullptr, order, foff, &nfronts, st, 1u, fkey))) return rc; // :160
        PGC_CUDA(cudaMemcpyAsync(h_foff, foff, sizeof(unsigned) * 2, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        const unsigned n0 = h_foff[1] - h_foff[0];
        if (diversity == 1u && n0 == 1u && (rc = fnds_device(ctx, d_f, NP, m, rank, nullptr, order, foff, &nfronts, st, 2u, fkey))) return rc;
        unsigned h_nb = 0;
        if (diversity == 0u) {
            h_nb = n0 > 1u ? n0 : 2u;
            if ((rc = sort_population_device(ctx, d_f, NP, m, sorted, st, h_nb))) return rc;
            PGC_CUDA(cudaMemcpyAsync(bnd, sorted, sizeof(unsigned) * h_nb, cudaMemcpyDeviceToDevice, st));
        } else if (diversity == 1u) {
            if (n0 > 1u) {
                nspso_delta_kernel<<<1, 64, 0, st>>>(d_f, NP, m, order, n0, delta);
                nspso_niche_kernel<<<nblk(static_cast<size_t>(n0) * 32, 256), 256, 0, st>>>(d_x, dim, order, n0, delta, keyv);
                if ((rc = argsort_less_f(keyv, n0, sl, st))) return rc;
                nspso_gather_u32_kernel<<<nblk(n0, 256), 256, 0, st>>>(order, sl, n0, bnd);
                ctx->launches.fetch_add(4, std::memory_order_relaxed);
                h_nb = n0;
            } else { // a single-point front: the point and the first member of the second front, :271-275
                PGC_CUDA(cudaMemcpyAsync(bnd, order, sizeof(unsigned) * 2, cudaMemcpyDeviceToDevice, st)); // order[foff[1]] == order[1]
                h_nb = 2;
            }
        } else {
            nspso_maxmin_kernel<<<nblk(static_cast<size_t>(NP) * 32, 256), 256, 0, st>>>(d_f, NP, m, keyv);
            if ((rc = argsort_less_f(keyv, NP, bnd, st))) return rc;
            nspso_maxmin_cut_kernel<<<1, 1, 0, st>>>(keyv, bnd, NP, nb);
            ctx->launches.fetch_add(3, std::memory_order_relaxed);
        }
        if (diversity != 2u) PGC_CUDA(cudaMemcpyAsync(nb, &h_nb, sizeof(unsigned), cudaMemcpyHostToDevice, st));
        // 2 - move
        MoveParams mp{d_x, bx, lb, ub, minv, maxv, V, x2, bnd, nb, NP, dim, leader_selection_range, omega, c1, c2, chi, seed, generation};
        nspso_move_kernel<<<nblk(nd, 256), 256, 0, st>>>(mp);
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        PGC_CUDA(cudaStreamSynchronize(st)); // h_nb (pageable) has been consumed
        // 3 - evaluate the moved swarm
        if ((rc = eval(prob, x2, NP, f2, st))) return rc;
        // 4 - archive = best N of (moved | archive); population = moved swarm
        PGC_CUDA(cudaMemcpyAsync(x2 + nd, bx, sizeof(double) * nd, cudaMemcpyDeviceToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(f2 + nm, bf, sizeof(double) * nm, cudaMemcpyDeviceToDevice, st));
        if (diversity != 2u) {
            if ((rc = sort_population_device(ctx, f2, 2 * static_cast<size_t>(NP), m, order2, st, NP))) return rc;
        } else {
            nspso_maxmin_kernel<<<nblk(static_cast<size_t>(2 * NP) * 32, 256), 256, 0, st>>>(f2, 2 * NP, m, keyv);
            if ((rc = argsort_less_f(keyv, 2 * NP, order2, st))) return rc;
            ctx->launches.fetch_add(2, std::memory_order_relaxed);
        }
        nspso_gather_rows_kernel<<<nblk(nd, 256), 256, 0, st>>>(x2, order2, NP, dim, bx);
        nspso_gather_rows_kernel<<<nblk(nm, 256), 256, 0, st>>>(f2, order2, NP, m, bf);
        PGC_CUDA(cudaMemcpyAsync(d_x, x2, sizeof(double) * nd, cudaMemcpyDeviceToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(d_f, f2, sizeof(double) * nm, cudaMemcpyDeviceToDevice, st));
        ctx->launches.fetch_add(2, std::memory_order_relaxed);
    }
    PGC_CUDA(cudaGetLastError());
    PGC_CUDA(cudaStreamSynchronize(st));
    return PGC_OK;
}

} // namespace pgc
