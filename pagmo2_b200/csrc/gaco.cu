// gaco.cu - pagmo::gaco::evolve (extended ant colony optimisation, reference src/algorithms/gaco.cpp:104-445) on a device-resident
// population of an unconstrained single-objective problem, memory = false.
//
// One generation (the reference's own batch structure, :207-320):
//   gaco_penalty_kernel   penalties against the oracle parameter (penalty_computation :506-549) + sort keys
//   cub radix sort        the population in order of penalty (:216-221; stable where the reference's std::sort leaves ties open)
//   gaco_archive_kernel   one CTA: first generation fills the solution archive with the best `ker` (:226-243); later ones merge the
//                         best `ker` of the population into it with the accuracy filter (update_sol_archive :563-675)
//   gaco_pheromone_kernel kernel weights + their cumulative sums (at generation 1 and at `threshold`), sigma per variable from
//                         the archive's extreme pairwise distances (pheromone_computation :690-796)
//   gaco_ants_kernel      one thread per ant: kernel choice, one normal deviate per variable with up to ten redraws, clamp,
//                         rounding of the integer tail (generate_new_ants :812-875); Philox substream (seed, generation, ant)
//   batch evaluation      the problem's evaluator (the bfe branch, :288-320)
//   gaco_finish_kernel    the population becomes the ants; champion / evalstop counter (:338-347); oracle update and the archive's
//                         penalties under the new oracle (:349-402)
// The scalar members of the algorithm live in GacoState on the device; a stopping criterion (:193-205) turns the remaining launches
// into no-ops, and the archive is written back into the population at the end unless one fired (:408-421).
#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <vector>

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
    template <class T> int alloc(T **out, size_t count)
    {
        void *p = nullptr;
        PGC_CUDA(cudaMallocAsync(&p, sizeof(T) * (count ? count : 1), st));
        owned.push_back(p);
        *out = static_cast<T *>(p);
        return PGC_OK;
    }
};

struct GacoState {
    double oracle, q;
    unsigned n_evalstop, n_impstop, gen_mark;
    unsigned long long fevals;
    double champ;
    unsigned stopped, gens_done;
};

__device__ __forceinline__ bool less_f(double a, double b) { return !(a != a) && ((b != b) || a < b); } // detail::less_than_f

__device__ __forceinline__ unsigned long long order_key(double v) // order-preserving, every NaN last (less_than_f)
{
    unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
    b = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
    return (v != v) ? 0xffffffffffffffffull : b;
}

// penalty_computation without constraints (m_res = 0), :506-549, written as the reference writes it
__device__ double gaco_penalty(double fitness, double oracle)
{
    const double res = 0.0;
    double alpha = 0.0;
    const double diff = fabs(fitness - oracle);
    double penalty = 0.0;
    if (fitness > oracle && res < diff / 3.0) {
        alpha = (diff * (6.0 * sqrt(3.0) - 2.0) / (6.0 * sqrt(3.0)) - res) / (diff - res);
    }
    if (fitness > oracle || res > 0.) {
        penalty = alpha * diff + (1 - alpha) * res;
    } else if (fitness <= oracle && res == 0.) {
        penalty = -diff;
    }
    return penalty;
}

__global__ void gaco_init_state_kernel(GacoState *S, const double *f, unsigned n)
{ // the population's champion: the first best individual
    __shared__ double s[256];
    double v = NAN;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x)
        if (less_f(f[i], v)) v = f[i];
    s[threadIdx.x] = v;
    __syncthreads();
    for (unsigned w = blockDim.x / 2; w; w >>= 1) {
        if (threadIdx.x < w && less_f(s[threadIdx.x + w], s[threadIdx.x])) s[threadIdx.x] = s[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        S->champ = s[0];
        S->stopped = 0;
        S->gens_done = 0;
    }
}

__global__ void gaco_penalty_kernel(GacoState *S, const double *f, unsigned n, unsigned impstop, unsigned evalstop, double *pen,
                                    unsigned long long *keys, unsigned *idx)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    // the stopping criteria, :193-205 (every thread reads the same counters; thread 0 of block 0 records the outcome)
    const bool stop = S->stopped || (impstop != 0u && S->n_impstop >= impstop) || (evalstop != 0u && S->n_evalstop >= evalstop);
    if (i < n) {
        const double p = gaco_penalty(f[i], S->oracle);
        pen[i] = p;
        keys[i] = order_key(p);
        idx[i] = i;
    }
    if (i == 0 && stop) S->stopped = 2; // 2: decided in this generation; becomes 1 in the archive kernel (other blocks may still read it)
}

// rows of the archive: [penalty | x (nx) | f]
__global__ void gaco_archive_kernel(GacoState *S, const double *x, const double *f, const double *pen, const unsigned *sl, unsigned nx,
                                    unsigned ker, int first, double acc, unsigned n_gen_mark, double *arch, double *tmp_arch, double *tp,
                                    unsigned *slp, unsigned *nsl, unsigned *n_new_out)
{
    const unsigned row = 1u + nx + 1u, t = threadIdx.x, T = blockDim.x;
    __shared__ int replace;
    if (t == 0 && S->stopped == 2u) S->stopped = 1u;
    __syncthreads();
    if (S->stopped) return;
    if (first) { // :226-243
        for (unsigned e = t; e < ker * row; e += T) {
            const unsigned i = e / row, c = e % row, src = sl[i];
            arch[e] = c == 0u ? pen[src] : (c <= nx ? x[static_cast<size_t>(src) * nx + (c - 1u)] : f[src]);
        }
        return;
    }
    // update_sol_archive, :563-675
    if (t == 0) {
        replace = pen[sl[0]] < arch[static_cast<size_t>(ker - 1u) * row];
    }
    __syncthreads();
    if (replace) {
        for (unsigned i = t; i < ker; i += T) {
            tp[i] = pen[sl[i]];
            tp[ker + i] = arch[static_cast<size_t>(i) * row];
        }
        __syncthreads();
        // stable order of the 2 ker penalties: the rank of every entry
        for (unsigned e = t; e < 2u * ker; e += T) {
            const double v = tp[e];
            unsigned r = 0;
            for (unsigned j = 0; j < 2u * ker; ++j) r += (less_f(tp[j], v) || (!less_f(v, tp[j]) && j < e)) ? 1u : 0u;
            slp[r] = e;
        }
        for (unsigned e = t; e < ker * row; e += T) tmp_arch[e] = arch[e];
        __syncthreads();
        if (t == 0) { // the accuracy filter, :617-633, as written
            unsigned count = 0, n_new = 0;
            if (!(slp[0] < ker)) ++count;
            nsl[n_new++] = 0;
            for (unsigned j = 1; j < 2u * ker; ++j) {
                if (fabs(tp[slp[j]] - tp[slp[count]]) < acc) {
                } else {
                    ++count;
                    nsl[n_new++] = j;
                }
            }
            *n_new_out = n_new;
            S->n_impstop = 1;
        }
        __syncthreads();
        const unsigned n_new = *n_new_out;
        // row 0 (:607-616) is rewritten by the loop below with the same source (nsl[0] = 0); rows ii < min(ker, n_new), :634-657
        const unsigned rows = ker < n_new ? ker : n_new;
        for (unsigned e = t; e < rows * row; e += T) {
            const unsigned ii = e / row, c = e % row, idx = slp[nsl[ii]];
            double v;
            if (idx < ker) {
                const unsigned src = sl[idx];
                v = c == 0u ? tp[idx] : (c <= nx ? x[static_cast<size_t>(src) * nx + (c - 1u)] : f[src]);
            } else {
                v = arch[static_cast<size_t>(idx - ker) * row + c];
            }
            tmp_arch[e] = v;
        }
        __syncthreads();
        for (unsigned e = t; e < ker * row; e += T) arch[e] = tmp_arch[e];
    } else if (t == 0) {
        ++S->n_impstop;
    }
    if (t == 0) { // :668-674
        if (S->n_evalstop == 1u || S->n_evalstop > 2u) ++S->gen_mark;
        if (S->gen_mark > n_gen_mark) S->gen_mark = 1;
    }
}

// block h < nx: sigma[h]; block nx: the kernel weights (only when gen == 1 or gen == threshold)
__global__ void gaco_pheromone_kernel(GacoState *S, const double *arch, const double *lb, const double *ub, unsigned nx, unsigned ncx,
                                      unsigned ker, unsigned gen, unsigned threshold, double focus, double *omega, double *pc, double *sigma)
{
    if (S->stopped) return;
    const unsigned row = 1u + nx + 1u, t = threadIdx.x, T = blockDim.x;
    if (blockIdx.x == nx) {
        if (t == 0 && (gen == 1u || gen == threshold)) { // :706-730
            if (gen == threshold) S->q = 0.01;
            const double q = S->q, k = static_cast<double>(ker);
            double sum_omega = 0;
            for (unsigned l = 1; l <= ker; ++l) {
                const double lm = l - 1.0;
                const double omega_new = 1.0 / (q * k * sqrt(2 * 3.141592653589793238462643383279502884)) * exp(-(lm * lm) / (2.0 * (q * q) * (k * k)));
                omega[l - 1u] = omega_new;
                sum_omega += omega_new;
            }
            double cumulative = 0;
            for (unsigned j = 0; j < ker; ++j) {
                cumulative += omega[j] / sum_omega;
                pc[j] = cumulative;
            }
        }
        return;
    }
    const unsigned h = blockIdx.x + 1u; // the archive column of variable h - 1
    __shared__ double smin[256], smax[256];
    double d_min = fabs(arch[h] - arch[row + h]), d_max = d_min; // :759-761
    const unsigned long long pairs = static_cast<unsigned long long>(ker) * ker;
    for (unsigned long long e = t; e < pairs; e += T) {
        const unsigned c = static_cast<unsigned>(e / ker), k = static_cast<unsigned>(e % ker);
        if (k > c) {
            const double d = fabs(arch[static_cast<size_t>(c) * row + h] - arch[static_cast<size_t>(k) * row + h]);
            d_min = fmin(d_min, d);
            d_max = fmax(d_max, d);
        }
    }
    smin[t] = d_min, smax[t] = d_max;
    __syncthreads();
    for (unsigned w = T / 2; w; w >>= 1) {
        if (t < w) {
            smin[t] = fmin(smin[t], smin[t + w]);
            smax[t] = fmax(smax[t], smax[t + w]);
        }
        __syncthreads();
    }
    if (t == 0) { // :778-795
        d_min = smin[0], d_max = smax[0];
        const double width = ub[h - 1u] - lb[h - 1u], gm = static_cast<double>(S->gen_mark);
        double s;
        if (focus != 0. && ((d_max - d_min) / gen > width / focus)) s = width / focus;
        else if (h <= ncx) s = (d_max - d_min) / gm;
        else s = fmax(fmax((d_max - d_min) / gm, 1.0 / gm), (1.0 - 1.0 / (sqrt(static_cast<double>(nx - ncx)))));
        sigma[h - 1u] = s;
    }
}

__device__ __forceinline__ double normal01(PhiloxStream &rs) // Box-Muller on two uniforms
{
    const double u1 = 1.0 - rs.next();
    const double u2 = rs.next();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

__global__ void gaco_ants_kernel(const GacoState *S, const double *arch, const double *pc, const double *sigma, const double *lb,
                                 const double *ub, unsigned n, unsigned nx, unsigned ncx, unsigned ker, unsigned long long seed,
                                 unsigned generation, double *ants)
{
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || S->stopped) return;
    const unsigned row = 1u + nx + 1u;
    PhiloxStream rs(seed, kTagGaco, generation, j);
    const double number = rs.next();
    unsigned k_omega = 0; // :833-846
    if (number <= pc[0]) k_omega = 0;
    else if (number > pc[ker - 2u]) k_omega = ker - 1u;
    else
        for (unsigned k = 1; k + 1u < ker; ++k)
            if (number > pc[k - 1u] && number <= pc[k]) k_omega = k;
    const double *mean = arch + static_cast<size_t>(k_omega) * row + 1u;
    for (unsigned h = 0; h < nx; ++h) { // :847-868
        const double l = lb[h], u = ub[h];
        double g_h = mean[h] + sigma[h] * normal01(rs);
        if (g_h < l || g_h > u) {
            int iter_while = 0;
            while ((g_h < l || g_h > u) && iter_while < 10) {
                g_h = mean[h] + sigma[h] * normal01(rs);
                ++iter_while;
            }
            if (g_h < l) g_h = l;
            if (g_h > u) g_h = u;
        }
        ants[static_cast<size_t>(j) * nx + h] = (h >= ncx) ? round(g_h) : g_h;
    }
}

__global__ void gaco_commit_kernel(const GacoState *S, const double *ants, const double *fnew, unsigned n, unsigned nx, double *x, double *f)
{
    if (S->stopped) return;
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e < static_cast<size_t>(n) * nx) x[e] = ants[e];
    if (e < n) f[e] = fnew[e];
}

__global__ void gaco_finish_kernel(GacoState *S, const double *fnew, unsigned n, unsigned nx, unsigned ker, double *arch)
{
    if (S->stopped) return;
    __shared__ double s[256];
    __shared__ int update;
    const unsigned row = 1u + nx + 1u, t = threadIdx.x;
    double v = NAN;
    for (unsigned i = t; i < n; i += blockDim.x)
        if (less_f(fnew[i], v)) v = fnew[i];
    s[t] = v;
    __syncthreads();
    for (unsigned w = blockDim.x / 2; w; w >>= 1) {
        if (t < w && less_f(s[t + w], s[t])) s[t] = s[t + w];
        __syncthreads();
    }
    if (t == 0) {
        const double champ_old = S->champ;
        if (less_f(s[0], S->champ)) S->champ = s[0];
        if (!less_f(S->champ, champ_old)) ++S->n_evalstop; // :338-347
        else S->n_evalstop = 1u;
        S->fevals += n;
        S->gens_done += 1u;
        update = arch[1u + nx] < S->oracle; // :349
        if (update) S->oracle = arch[1u + nx];
    }
    __syncthreads();
    if (update)
        for (unsigned r = t; r < ker; r += blockDim.x) arch[static_cast<size_t>(r) * row] = gaco_penalty(arch[static_cast<size_t>(r) * row + 1u + nx], S->oracle);
}

__global__ void gaco_writeback_kernel(const GacoState *S, const double *arch, unsigned nx, unsigned ker, double *x, double *f)
{ // :408-421; a stopping criterion returned the population as it was
    if (S->stopped) return;
    const unsigned row = 1u + nx + 1u;
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(ker) * nx) return;
    const unsigned i = static_cast<unsigned>(e / nx), c = static_cast<unsigned>(e % nx);
    x[e] = arch[static_cast<size_t>(i) * row + 1u + c];
    if (c == 0u) f[i] = arch[static_cast<size_t>(i) * row + 1u + nx];
}

} // namespace

int gaco_evolve_device(pgc_problem *prob, double *d_x, double *d_f, unsigned n, unsigned gens, unsigned ker, double acc, unsigned threshold,
                       unsigned n_gen_mark, unsigned impstop, unsigned evalstop, double focus, unsigned long long seed, unsigned first_generation,
                       pgc_gaco_state *state, unsigned *gens_done,
                       int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st)
{
    pgc_ctx *ctx = prob->ctx;
    const unsigned nx = static_cast<unsigned>(prob->nx), ncx = nx - static_cast<unsigned>(prob->nix);
    if (gens_done) *gens_done = 0;
    if (n == 0u || gens == 0u) return PGC_OK; // :147-156
    // constructor and evolve checks, gaco.cpp:62-94,157-171
    PGC_REQUIRE(acc >= 0., "The accuracy parameter must be >=0, while a value of %g was detected", acc);
    PGC_REQUIRE(focus >= 0., "The focus parameter must be >=0  while a value of %g was detected", focus);
    PGC_REQUIRE(threshold >= 1u && threshold <= gens, "If memory is inactive, the threshold parameter must be either in [1,m_gen] while a value of %u was detected", threshold);
    PGC_REQUIRE(state->q >= 0., "The convergence speed parameter must be >=0  while a value of %g was detected", state->q);
    PGC_REQUIRE(ker >= 2u, "The ker size parameter must be >=2  while a value of %u was detected", ker);
    PGC_REQUIRE(n >= 2u, "GACO: Ant Colony Optimization needs at least 2 individuals in the population, %u detected", n);
    PGC_REQUIRE(ker <= n, "GACO: Ant Colony Optimization cannot work with a solution archive bigger than the population size");
    PGC_REQUIRE(prob->nobj == 1, "Multiple objectives detected in %s instance. GACO: Ant Colony Optimization cannot deal with them", prob->name.c_str());
    PGC_REQUIRE(n_gen_mark >= 1u, "gaco: n_gen_mark must be at least 1");
    const unsigned row = 1u + nx + 1u;
    Scratch sc(st);
    GacoState *S;
    double *arch, *tmp_arch, *pen, *tp, *omega, *pc, *sigma, *ants, *fnew, *lb, *ub;
    unsigned long long *k0, *k1;
    unsigned *i0, *sl, *slp, *nsl, *n_new;
    unsigned char *ws = nullptr;
    size_t ws_bytes = 0;
    int rc;
    if ((rc = sc.alloc(&S, 1)) || (rc = sc.alloc(&arch, static_cast<size_t>(ker) * row)) || (rc = sc.alloc(&tmp_arch, static_cast<size_t>(ker) * row))
        || (rc = sc.alloc(&pen, n)) || (rc = sc.alloc(&tp, 2 * static_cast<size_t>(ker))) || (rc = sc.alloc(&omega, ker)) || (rc = sc.alloc(&pc, ker))
        || (rc = sc.alloc(&sigma, nx)) || (rc = sc.alloc(&ants, static_cast<size_t>(n) * nx)) || (rc = sc.alloc(&fnew, n)) || (rc = sc.alloc(&lb, nx))
        || (rc = sc.alloc(&ub, nx)) || (rc = sc.alloc(&k0, n)) || (rc = sc.alloc(&k1, n)) || (rc = sc.alloc(&i0, n)) || (rc = sc.alloc(&sl, n))
        || (rc = sc.alloc(&slp, 2 * static_cast<size_t>(ker))) || (rc = sc.alloc(&nsl, 2 * static_cast<size_t>(ker))) || (rc = sc.alloc(&n_new, 1)))
        return rc;
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, ws_bytes, k0, k1, i0, sl, static_cast<int>(n), 0, 64, st));
    if ((rc = sc.alloc(&ws, ws_bytes))) return rc;
    GacoState h{};
    h.oracle = state->oracle, h.q = state->q, h.n_evalstop = state->n_evalstop, h.n_impstop = state->n_impstop, h.gen_mark = state->gen_mark;
    h.fevals = state->fevals;
    PGC_CUDA(cudaMemcpyAsync(S, &h, sizeof(GacoState), cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(lb, prob->lb.data(), sizeof(double) * nx, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(ub, prob->ub.data(), sizeof(double) * nx, cudaMemcpyHostToDevice, st));
    gaco_init_state_kernel<<<1, 256, 0, st>>>(S, d_f, n);
    for (unsigned gen = 1; gen <= gens; ++gen) {
        const unsigned generation = first_generation + (gen - 1u);
        gaco_penalty_kernel<<<nblk(n, 256), 256, 0, st>>>(S, d_f, n, impstop, evalstop, pen, k0, i0);
        PGC_CUDA(cub::DeviceRadixSort::SortPairs(ws, ws_bytes, k0, k1, i0, sl, static_cast<int>(n), 0, 64, st));
        gaco_archive_kernel<<<1, 256, 0, st>>>(S, d_x, d_f, pen, sl, nx, ker, gen == 1u ? 1 : 0, acc, n_gen_mark, arch, tmp_arch, tp, slp, nsl, n_new);
        gaco_pheromone_kernel<<<nx + 1u, 256, 0, st>>>(S, arch, lb, ub, nx, ncx, ker, gen, threshold, focus, omega, pc, sigma);
        gaco_ants_kernel<<<nblk(n, 128), 128, 0, st>>>(S, arch, pc, sigma, lb, ub, n, nx, ncx, ker, seed, generation, ants);
        PGC_CUDA(cudaGetLastError());
        // (a stopped run evaluates stale ants into fnew; nothing reads them)
        if ((rc = eval(prob, ants, n, fnew, st))) return rc;
        gaco_commit_kernel<<<nblk(static_cast<size_t>(n) * nx, 256), 256, 0, st>>>(S, ants, fnew, n, nx, d_x, d_f);
        gaco_finish_kernel<<<1, 256, 0, st>>>(S, fnew, n, nx, ker, arch);
        ctx->launches.fetch_add(7, std::memory_order_relaxed);
    }
    gaco_writeback_kernel<<<nblk(static_cast<size_t>(ker) * nx, 256), 256, 0, st>>>(S, arch, nx, ker, d_x, d_f);
    PGC_CUDA(cudaGetLastError());
    PGC_CUDA(cudaMemcpyAsync(&h, S, sizeof(GacoState), cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    state->oracle = h.oracle, state->q = h.q, state->n_evalstop = h.n_evalstop, state->n_impstop = h.n_impstop, state->gen_mark = h.gen_mark;
    state->fevals = h.fevals;
    if (gens_done) *gens_done = h.gens_done;
    return PGC_OK;
}

} // namespace pgc
