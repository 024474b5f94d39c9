// gaco.cu - pagmo::gaco::evolve (extended ant colony optimisation, reference src/algorithms/gaco.cpp:104-445) on a device-resident
// population of an unconstrained single-objective problem, memory = false; and pagmo::maco::evolve (maco.cpp:88-533), its
// multi-objective sibling, at the end of the file (same pheromone values and ants).
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

#include <algorithm>
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

__global__ void gaco_init_state_kernel(GacoState *S, const double *f, unsigned n, int has_champion, double champion)
{ // the population's champion: the best individual it ever held (handed in), else the first best of the ones it holds
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
        S->champ = (has_champion && less_f(champion, s[0])) ? champion : s[0];
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
// (archive rows of `row` doubles with the decision vector at column `xoff`: [penalty | x | f] for gaco, [x | f] for maco)
__global__ void gaco_pheromone_kernel(GacoState *S, const double *arch, const double *lb, const double *ub, unsigned nx, unsigned ncx,
                                      unsigned ker, unsigned gen, unsigned threshold, double focus, double *omega, double *pc, double *sigma,
                                      unsigned row, unsigned xoff, unsigned memory = 0u, unsigned counter = 0u)
{
    if (S->stopped) return;
    const unsigned t = threadIdx.x, T = blockDim.x;
    if (blockIdx.x == nx) {
        // memory = false: at generation 1 and at `threshold` (:706-730); memory = true: every generation, the switch on the call counter (:732-752)
        if (t == 0 && (memory || gen == 1u || gen == threshold)) {
            if (memory ? counter == threshold : gen == threshold) S->q = 0.01;
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
    const unsigned h = blockIdx.x + xoff, v = blockIdx.x; // the archive column of variable v
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
        const double width = ub[v] - lb[v], gm = static_cast<double>(S->gen_mark);
        double s;
        if (focus != 0. && ((d_max - d_min) / (memory ? counter : gen) > width / focus)) s = width / focus; // :778-784
        else if (v < ncx) s = (d_max - d_min) / gm;
        else s = fmax(fmax((d_max - d_min) / gm, 1.0 / gm), (1.0 - 1.0 / (sqrt(static_cast<double>(nx - ncx)))));
        sigma[v] = s;
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
                                 unsigned generation, double *ants, unsigned row, unsigned xoff)
{
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || S->stopped) return;
    PhiloxStream rs(seed, kTagGaco, generation, j);
    const double number = rs.next();
    unsigned k_omega = 0; // :833-846
    if (number <= pc[0]) k_omega = 0;
    else if (number > pc[ker - 2u]) k_omega = ker - 1u;
    else
        for (unsigned k = 1; k + 1u < ker; ++k)
            if (number > pc[k - 1u] && number <= pc[k]) k_omega = k;
    const double *mean = arch + static_cast<size_t>(k_omega) * row + xoff;
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

// the reference's log line (gaco.cpp:254-287 inside the loop, :405-445 after it): (gen, fevals, best, kernel, oracle, dx, dp) with dx / dp
// the spread of the archive between its first and last row.  final != 0: the line after the loop, whose "best" is the population's
// champion.  A stopped run logs nothing more.
__global__ void gaco_log_kernel(const GacoState *S, const double *arch, unsigned nx, unsigned ker, unsigned gen, int final, double *rows,
                                unsigned *count, unsigned max_rows, unsigned row_len)
{
    if (S->stopped || threadIdx.x || blockIdx.x) return;
    const unsigned row = 1u + nx + 1u, r = *count;
    if (r >= max_rows || row_len < 7u) return;
    double dx = 0.;
    for (unsigned i = 0; i < nx; ++i) dx += fabs(arch[static_cast<size_t>(ker - 1u) * row + 1u + i] - arch[1u + i]);
    double *o = rows + static_cast<size_t>(r) * row_len;
    o[0] = gen, o[1] = static_cast<double>(S->fevals), o[2] = final ? S->champ : arch[1u + nx], o[3] = ker, o[4] = S->oracle, o[5] = dx;
    o[6] = fabs(arch[static_cast<size_t>(ker - 1u) * row] - arch[0]);
    *count = r + 1u;
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
    const bool memory = state->memory != 0u;
    if (memory) {
        PGC_REQUIRE(threshold >= 1u, "If memory is active, the threshold parameter must be >=1 while a value of %u was detected", threshold);
        PGC_REQUIRE(state->h_archive && state->h_archive_len >= static_cast<size_t>(ker) * (nx + 2u),
                    "gaco with memory keeps its archive in h_archive: %zu doubles are needed", static_cast<size_t>(ker) * (nx + 2u));
        ++state->counter; // :106-108
    } else {
        PGC_REQUIRE(threshold >= 1u && threshold <= gens, "If memory is inactive, the threshold parameter must be either in [1,m_gen] while a value of %u was detected", threshold);
    }
    const unsigned counter = state->counter;
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
    gaco_init_state_kernel<<<1, 256, 0, st>>>(S, d_f, n, static_cast<int>(state->has_champion), state->champion_f);
    if (memory && counter > 1u) // sol_archive = m_sol_archive, :223-225
        PGC_CUDA(cudaMemcpyAsync(arch, state->h_archive, sizeof(double) * ker * row, cudaMemcpyHostToDevice, st));
    for (unsigned gen = 1; gen <= gens; ++gen) {
        const unsigned generation = first_generation + (gen - 1u);
        gaco_penalty_kernel<<<nblk(n, 256), 256, 0, st>>>(S, d_f, n, impstop, evalstop, pen, k0, i0);
        PGC_CUDA(cub::DeviceRadixSort::SortPairs(ws, ws_bytes, k0, k1, i0, sl, static_cast<int>(n), 0, 64, st));
        gaco_archive_kernel<<<1, 256, 0, st>>>(S, d_x, d_f, pen, sl, nx, ker, (gen == 1u && counter < 2u) ? 1 : 0, acc, n_gen_mark, arch, tmp_arch, tp, slp, nsl,
                                               n_new);
        if ((gen != gens || memory) && log_due(gen)) // 3 - :254-287 (memory = false: every due generation but the last)
            gaco_log_kernel<<<1, 32, 0, st>>>(S, arch, nx, ker, gen, 0, tls_log->d_rows, tls_log->d_count, tls_log->max_rows, tls_log->row_len);
        gaco_pheromone_kernel<<<nx + 1u, 256, 0, st>>>(S, arch, lb, ub, nx, ncx, ker, gen, threshold, focus, omega, pc, sigma, row, 1u, memory ? 1u : 0u, counter);
        gaco_ants_kernel<<<nblk(n, 128), 128, 0, st>>>(S, arch, pc, sigma, lb, ub, n, nx, ncx, ker, seed, generation, ants, row, 1u);
        PGC_CUDA(cudaGetLastError());
        // (a stopped run evaluates stale ants into fnew; nothing reads them)
        if ((rc = eval(prob, ants, n, fnew, st))) return rc;
        gaco_commit_kernel<<<nblk(static_cast<size_t>(n) * nx, 256), 256, 0, st>>>(S, ants, fnew, n, nx, d_x, d_f);
        gaco_finish_kernel<<<1, 256, 0, st>>>(S, fnew, n, nx, ker, arch);
        ctx->launches.fetch_add(7, std::memory_order_relaxed);
    }
    if (!memory) gaco_writeback_kernel<<<nblk(static_cast<size_t>(ker) * nx, 256), 256, 0, st>>>(S, arch, nx, ker, d_x, d_f); // :408-421
    else PGC_CUDA(cudaMemcpyAsync(state->h_archive, arch, sizeof(double) * ker * row, cudaMemcpyDeviceToHost, st));            // m_sol_archive
    if (!memory && tls_log && tls_log->verbosity && (gens % tls_log->verbosity == 1u || tls_log->verbosity == 1u)) // :405-445
        gaco_log_kernel<<<1, 32, 0, st>>>(S, arch, nx, ker, gens, 1, tls_log->d_rows, tls_log->d_count, tls_log->max_rows, tls_log->row_len);
    PGC_CUDA(cudaGetLastError());
    PGC_CUDA(cudaMemcpyAsync(&h, S, sizeof(GacoState), cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    state->oracle = h.oracle, state->q = h.q, state->n_evalstop = h.n_evalstop, state->n_impstop = h.n_impstop, state->gen_mark = h.gen_mark;
    state->fevals = h.fevals;
    state->champion_f = h.champ, state->has_champion = 1u;
    if (gens_done) *gens_done = h.gens_done;
    return PGC_OK;
}


// ---- maco (reference src/algorithms/maco.cpp:88-533) ---------------------------------------------------------------------------
// Multi-objective hypervolume-based ant colony optimisation: gaco's pheromone values and ants on an archive that is rebuilt every
// generation from the non-dominated fronts of (archive + population), each front ordered by DEcreasing exclusive hypervolume
// contribution.  The device does the quadratic and the batched parts - non-dominated sorting (fnds_device), the contributions of
// a front (hv_device: sweeps for 2 / 3 objectives, WFG above), the ants, the evaluation; the host walks the fronts: it needs
// their sizes, a reference point per front (nadir + offset) and the order of a handful of contributions, and keeps the three
// counters of the algorithm object.
namespace
{

__global__ void maco_gather_rows_kernel(const double *src, const unsigned *idx, unsigned k, unsigned width, double *dst)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(k) * width) return;
    dst[e] = src[static_cast<size_t>(idx[e / width]) * width + e % width];
}

// archive row i <- merged individual src[i]: [x | f]
__global__ void maco_fill_archive_kernel(const double *mx, const double *mf, const unsigned *src, unsigned ker, unsigned nx, unsigned m, double *arch)
{
    const unsigned row = nx + m;
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(ker) * row) return;
    const unsigned i = static_cast<unsigned>(e / row), c = static_cast<unsigned>(e % row), s = src[i];
    arch[e] = c < nx ? mx[static_cast<size_t>(s) * nx + c] : mf[static_cast<size_t>(s) * m + (c - nx)];
}

// merged <- [archive rows ; population]
__global__ void maco_merge_kernel(const double *arch, const double *x, const double *f, unsigned ker, unsigned n, unsigned nx, unsigned m,
                                  double *mx, double *mf)
{
    const unsigned row = nx + m;
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t np = static_cast<size_t>(ker) + n;
    if (e < np * nx) {
        const size_t j = e / nx, c = e % nx;
        mx[e] = j < ker ? arch[j * row + c] : x[(j - ker) * nx + c];
    }
    if (e < np * m) {
        const size_t j = e / m, c = e % m;
        mf[e] = j < ker ? arch[j * row + nx + c] : f[(j - ker) * m + c];
    }
}

__global__ void maco_writeback_kernel(const double *arch, unsigned ker, unsigned nx, unsigned m, double *x, double *f)
{ // :535-545
    const unsigned row = nx + m;
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(ker) * row) return;
    const unsigned i = static_cast<unsigned>(e / row), c = static_cast<unsigned>(e % row);
    if (c < nx) x[static_cast<size_t>(i) * nx + c] = arch[e];
    else f[static_cast<size_t>(i) * m + (c - nx)] = arch[e];
}

inline bool greater_than_f(double a, double b) // detail::greater_than_f
{
    if (!std::isnan(a)) return !std::isnan(b) ? a > b : false;
    return !std::isnan(b);
}

} // namespace

int maco_evolve_device(pgc_problem *prob, double *d_x, double *d_f, unsigned n, unsigned gens, unsigned ker, unsigned threshold,
                       unsigned n_gen_mark, unsigned evalstop, double focus, unsigned long long seed, unsigned first_generation,
                       pgc_maco_state *state, unsigned *gens_done,
                       int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st)
{
    pgc_ctx *ctx = prob->ctx;
    const unsigned nx = static_cast<unsigned>(prob->nx), ncx = nx - static_cast<unsigned>(prob->nix), m = static_cast<unsigned>(prob->nobj);
    if (gens_done) *gens_done = 0;
    // constructor and evolve checks, maco.cpp:64-83,126-152
    PGC_REQUIRE(focus >= 0., "The focus parameter must be >=0  while a value of %g was detected", focus);
    PGC_REQUIRE(n > 0u, "MHACO: Multi-objective Hypervolume-based Ant Colony Optimization cannot work on an empty population");
    if (gens == 0u) return PGC_OK;
    PGC_REQUIRE(threshold >= 1u && threshold <= gens, "If memory is inactive, the threshold parameter must be either in [1,m_gen] while a value of %u was detected", threshold);
    PGC_REQUIRE(ker <= n, "MHACO: Multi-objective Hypervolume-based Ant Colony Optimization cannot work with a solution archive bigger than the population size");
    PGC_REQUIRE(ker >= 2u, "maco: the pheromone values need an archive of at least two solutions (ker = %u)", ker);
    PGC_REQUIRE(m >= 2u, "This is a multiobjective algorithm, while number of objectives detected in %s is %u", prob->name.c_str(), m);
    PGC_REQUIRE(n_gen_mark >= 1u, "maco: n_gen_mark must be at least 1");
    const unsigned row = nx + m, np_max = ker + n;
    Scratch sc(st);
    GacoState *S;
    double *arch, *mx, *mf, *lf, *contrib, *omega, *pc, *sigma, *ants, *fnew, *lb, *ub;
    unsigned *rank, *order, *foff, *src;
    int rc;
    if ((rc = sc.alloc(&S, 1)) || (rc = sc.alloc(&arch, static_cast<size_t>(ker) * row)) || (rc = sc.alloc(&mx, static_cast<size_t>(np_max) * nx))
        || (rc = sc.alloc(&mf, static_cast<size_t>(np_max) * m)) || (rc = sc.alloc(&lf, static_cast<size_t>(np_max) * m))
        || (rc = sc.alloc(&contrib, np_max)) || (rc = sc.alloc(&omega, ker)) || (rc = sc.alloc(&pc, ker)) || (rc = sc.alloc(&sigma, nx))
        || (rc = sc.alloc(&ants, static_cast<size_t>(n) * nx)) || (rc = sc.alloc(&fnew, static_cast<size_t>(n) * m)) || (rc = sc.alloc(&lb, nx))
        || (rc = sc.alloc(&ub, nx)) || (rc = sc.alloc(&rank, np_max)) || (rc = sc.alloc(&order, np_max)) || (rc = sc.alloc(&foff, np_max + 1u))
        || (rc = sc.alloc(&src, np_max)))
        return rc;
    PGC_CUDA(cudaMemcpyAsync(lb, prob->lb.data(), sizeof(double) * nx, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(ub, prob->ub.data(), sizeof(double) * nx, cudaMemcpyHostToDevice, st));
    std::vector<double> h_mf(static_cast<size_t>(np_max) * m), h_arch_fit(static_cast<size_t>(ker) * m, 1.0), h_contrib(np_max), ref(m), idp(m);
    std::vector<unsigned> h_order(np_max), h_foff(np_max + 1u), h_src(ker), sl(np_max);
    bool stopped = false;
    unsigned done = 0;
    for (unsigned gen = 1; gen <= gens; ++gen) {
        const unsigned generation = first_generation + (gen - 1u);
        // the individuals the archive is rebuilt from: the population (first generation) or archive + population (:264-284)
        const unsigned np = gen == 1u ? n : np_max;
        if (gen == 1u) {
            PGC_CUDA(cudaMemcpyAsync(mx, d_x, sizeof(double) * n * nx, cudaMemcpyDeviceToDevice, st));
            PGC_CUDA(cudaMemcpyAsync(mf, d_f, sizeof(double) * n * m, cudaMemcpyDeviceToDevice, st));
        } else {
            maco_merge_kernel<<<nblk(static_cast<size_t>(np) * std::max(nx, m), 256), 256, 0, st>>>(arch, d_x, d_f, ker, n, nx, m, mx, mf);
            PGC_CUDA(cudaGetLastError());
        }
        PGC_CUDA(cudaMemcpyAsync(h_mf.data(), mf, sizeof(double) * np * m, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        // rebuild the archive from the fronts (:177-262 first generation, offset 0.1; :318-410 afterwards, offset 0.01).  The first
        // generation does it BEFORE the ideal-point test below, later ones after it - and a stop leaves the archive as it was.
        const auto rebuild = [&](double offset, bool first) -> int {
            unsigned nfronts = 0;
            if (np < 2u) { // fast_non_dominated_sorting needs two points; one point is one front
                h_order[0] = 0, h_foff[0] = 0, h_foff[1] = 1, nfronts = 1;
            } else {
                if (int r = fnds_device(ctx, mf, np, m, rank, nullptr, order, foff, &nfronts, st)) return r;
                PGC_CUDA(cudaMemcpyAsync(h_order.data(), order, sizeof(unsigned) * np, cudaMemcpyDeviceToHost, st));
                PGC_CUDA(cudaMemcpyAsync(h_foff.data(), foff, sizeof(unsigned) * (nfronts + 1u), cudaMemcpyDeviceToHost, st));
                PGC_CUDA(cudaStreamSynchronize(st));
            }
            unsigned i_arch = 0;
            for (unsigned fr = 0; fr < nfronts && i_arch < ker; ++fr) {
                const unsigned *idxs = h_order.data() + h_foff[fr], k = h_foff[fr + 1u] - h_foff[fr];
                for (unsigned c = 0; c < m; ++c) { // hypervolume::refpoint(offset), hypervolume.cpp:160-181
                    double v = h_mf[static_cast<size_t>(idxs[0]) * m + c];
                    for (unsigned i = 1; i < k; ++i) v = std::max(v, h_mf[static_cast<size_t>(idxs[i]) * m + c]);
                    ref[c] = v + offset;
                }
                if (k == 1u) { // hypervolume::contributions' trivial case, hypervolume.cpp:292-297
                    double v = 1.0;
                    for (unsigned c = 0; c < m; ++c) v *= (h_mf[static_cast<size_t>(idxs[0]) * m + c] - ref[c]);
                    h_contrib[0] = std::fabs(v);
                } else {
                    maco_gather_rows_kernel<<<nblk(static_cast<size_t>(k) * m, 256), 256, 0, st>>>(mf, order + h_foff[fr], k, m, lf);
                    PGC_CUDA(cudaGetLastError());
                    if (int r = hv_device(ctx, lf, k, m, ref.data(), 0, contrib, st)) return r;
                    PGC_CUDA(cudaMemcpyAsync(h_contrib.data(), contrib, sizeof(double) * k, cudaMemcpyDeviceToHost, st));
                    PGC_CUDA(cudaStreamSynchronize(st));
                    ctx->launches.fetch_add(1, std::memory_order_relaxed);
                }
                for (unsigned i = 0; i < k; ++i) sl[i] = i;
                std::stable_sort(sl.begin(), sl.begin() + k, [&](unsigned a, unsigned b) { return greater_than_f(h_contrib[a], h_contrib[b]); });
                for (unsigned i = 0; i < k && i_arch < ker; ++i, ++i_arch) {
                    h_src[i_arch] = idxs[sl[i]];
                    if (first) std::copy_n(h_mf.begin() + static_cast<std::ptrdiff_t>(idxs[sl[i]]) * m, m, h_arch_fit.begin() + static_cast<std::ptrdiff_t>(i_arch) * m);
                }
                if (i_arch >= ker && fr == 0u) { // the extremities of an overflowing first front, :231-259, as written
                    for (unsigned c = 0; c < m; ++c) {
                        double v = h_mf[static_cast<size_t>(idxs[0]) * m + c];
                        for (unsigned i = 1; i < k; ++i) v = std::min(v, h_mf[static_cast<size_t>(idxs[i]) * m + c]);
                        idp[c] = v;
                    }
                    std::vector<unsigned> border;
                    for (unsigned c = 0; c < m; ++c)
                        for (unsigned i = 0; i < k; ++i)
                            if (h_mf[static_cast<size_t>(idxs[i]) * m + c] == idp[c]) {
                                border.push_back(idxs[i]);
                                break;
                            }
                    for (unsigned c = 0; c < m && c < ker; ++c) {
                        h_src[ker - 1u - c] = border[c];
                        // sol_archive_fit[ker - 1 - c] = the fitness in archive row c AT THAT MOMENT (rows ker-1-c' with c' < c are
                        // already overwritten by their border points)
                        const unsigned from = h_src[c];
                        std::copy_n(h_mf.begin() + static_cast<std::ptrdiff_t>(from) * m, m, h_arch_fit.begin() + static_cast<std::ptrdiff_t>(ker - 1u - c) * m);
                    }
                }
            }
            PGC_CUDA(cudaMemcpyAsync(src, h_src.data(), sizeof(unsigned) * ker, cudaMemcpyHostToDevice, st));
            maco_fill_archive_kernel<<<nblk(static_cast<size_t>(ker) * row, 256), 256, 0, st>>>(mx, mf, src, ker, nx, m, arch);
            PGC_CUDA(cudaGetLastError());
            PGC_CUDA(cudaStreamSynchronize(st)); // h_src is reused by the next call
            ctx->launches.fetch_add(2, std::memory_order_relaxed);
            return PGC_OK;
        };
        if (gen == 1u) {
            if ((rc = rebuild(0.1, true))) return rc;
        } else {
            std::copy_n(h_mf.begin(), static_cast<size_t>(ker) * m, h_arch_fit.begin()); // sol_archive_fit = the archive's fitness, :267-270
        }
        // the ideal point of the archive against the ideal point of everything, :289-318
        bool check = false;
        std::vector<double> ideal_arch(m);
        for (unsigned c = 0; c < m; ++c) {
            double a = h_arch_fit[c], b = h_mf[c];
            for (unsigned j = 1; j < ker; ++j) a = std::min(a, h_arch_fit[static_cast<size_t>(j) * m + c]);
            if (gen == 1u) b = a;
            else
                for (unsigned j = 1; j < np; ++j) b = std::min(b, h_mf[static_cast<size_t>(j) * m + c]);
            ideal_arch[c] = a;
            if (a != b && !check) check = true;
        }
        if (check) ++state->n_evalstop;
        else state->n_evalstop = 0;
        if (state->n_evalstop == 0u || state->n_evalstop > 2u) ++state->gen_mark;
        if (state->gen_mark > n_gen_mark) state->gen_mark = 1;
        if (evalstop != 0u && state->n_evalstop >= evalstop) { // `return pop`, :312-317
            stopped = true;
            break;
        }
        if (gen > 1u && (rc = rebuild(0.01, false))) return rc;
        if (log_due(gen)) { // 2 - maco.cpp:415-463: (gen, fevals, the ideal point of the archive as the test above saw it)
            tls_log->host_rows.push_back(gen);
            tls_log->host_rows.push_back(static_cast<double>(gen - 1u) * n);
            tls_log->host_rows.insert(tls_log->host_rows.end(), ideal_arch.begin(), ideal_arch.end());
        }
        // pheromone values and ants: gaco's kernels on rows [x | f]
        if (gen == threshold) state->q = 0.01;
        GacoState h{};
        h.q = state->q, h.gen_mark = state->gen_mark;
        PGC_CUDA(cudaMemcpyAsync(S, &h, sizeof(GacoState), cudaMemcpyHostToDevice, st));
        gaco_pheromone_kernel<<<nx + 1u, 256, 0, st>>>(S, arch, lb, ub, nx, ncx, ker, gen, threshold, focus, omega, pc, sigma, row, 0u);
        gaco_ants_kernel<<<nblk(n, 128), 128, 0, st>>>(S, arch, pc, sigma, lb, ub, n, nx, ncx, ker, seed, generation, ants, row, 0u);
        PGC_CUDA(cudaGetLastError());
        PGC_CUDA(cudaStreamSynchronize(st)); // `h` leaves scope
        if ((rc = eval(prob, ants, n, fnew, st))) return rc;
        PGC_CUDA(cudaMemcpyAsync(d_x, ants, sizeof(double) * n * nx, cudaMemcpyDeviceToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(d_f, fnew, sizeof(double) * n * m, cudaMemcpyDeviceToDevice, st));
        ctx->launches.fetch_add(3, std::memory_order_relaxed);
        ++done;
    }
    if (!stopped) {
        maco_writeback_kernel<<<nblk(static_cast<size_t>(ker) * row, 256), 256, 0, st>>>(arch, ker, nx, m, d_x, d_f);
        PGC_CUDA(cudaGetLastError());
    }
    PGC_CUDA(cudaStreamSynchronize(st));
    if (gens_done) *gens_done = done;
    return PGC_OK;
}

} // namespace pgc
