// algo_log.cu - the log lines a reference UDA records when its verbosity is > 0 (get_log()), computed on the device from the
// resident population at the generations the reference logs (gen % verbosity == 1, or every generation for verbosity 1):
//   de      (gen, fevals, best, dx, df)                         de.cpp:324-347
//   sade    (gen, fevals, best, F, CR, dx, df)                  sade.cpp:556-580
//   de1220  (gen, fevals, best, F, CR, variant, dx, df)         de1220.cpp:570-595
//   pso_gen (gen, fevals, gbest, mean velocity, mean lbest, average distance)   pso_gen.cpp:464-518
//   sga     (gen, fevals, best, improvement)                    sga.cpp:252-274 (verbosity 1: only the generations that improve)
//   cmaes   (gen, fevals, best, dx, df, sigma)                  cmaes.cpp:276-296 (appended on the host: cmaes.cu runs its loop there)
//   nsga2   (gen, fevals, ideal point)                          nsga2.cpp:144-173 (before the generation's variation)
//   nspso   (gen, fevals, ideal point of the archive)           nspso.cpp:163-192
// Rows are doubles, `row_len` per line, appended in the stream's order at *d_count.  The generation loops call the hooks through
// pgc::tls_log (set by pgc_algo_evolve_logged_device for the duration of one call on the calling thread).
#include "pgc_internal.cuh"

namespace pgc
{

thread_local LogSink *tls_log = nullptr;

namespace
{

inline unsigned nblk(size_t n, unsigned t) { return static_cast<unsigned>((n + t - 1) / t); }

// column-wise minimum (pagmo::ideal, multi_objective.cpp:351-371) appended to (gen, fevals)
__global__ void log_ideal_kernel(const double *f, unsigned n, unsigned m, double gen, double fevals, double *rows, unsigned *count,
                                 unsigned max_rows, unsigned row_len)
{
    __shared__ double s[256];
    __shared__ unsigned row;
    if (threadIdx.x == 0) row = *count;
    __syncthreads();
    if (row >= max_rows) return;
    double *out = rows + static_cast<size_t>(row) * row_len;
    for (unsigned k = 0; k < m; ++k) {
        double v = INFINITY;
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) v = fmin(v, f[static_cast<size_t>(i) * m + k]);
        s[threadIdx.x] = v;
        __syncthreads();
        for (unsigned w = blockDim.x / 2; w; w >>= 1) {
            if (threadIdx.x < w) s[threadIdx.x] = fmin(s[threadIdx.x], s[threadIdx.x + w]);
            __syncthreads();
        }
        if (threadIdx.x == 0) out[2 + k] = s[0];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = gen;
        out[1] = fevals;
        *count = row + 1u;
    }
}

// pso_gen.cpp:470-505.  One CTA: best and mean of lbfit; the reference's running "mean velocity" (it divides the running sum by the
// row length after every particle, :487-494 - reproduced as written); the mean pairwise distance of the current positions in units
// of the box (:496-510) is accumulated by log_pso_dist_kernel.
__global__ void log_pso_dist_kernel(const double *X, const double *lb, const double *ub, unsigned n, unsigned dim, double *acc)
{
    __shared__ double s[256];
    const unsigned i = blockIdx.x;
    double local = 0.;
    for (unsigned j = i + 1u + threadIdx.x; j < n; j += blockDim.x) {
        double a = 0.;
        for (unsigned k = 0; k < dim; ++k) {
            const double w = ub[k] - lb[k];
            if (ub[k] > lb[k]) {
                const double d = X[static_cast<size_t>(i) * dim + k] - X[static_cast<size_t>(j) * dim + k];
                a += d * d / w / w;
            }
        }
        local += sqrt(a);
    }
    s[threadIdx.x] = local;
    __syncthreads();
    for (unsigned w = blockDim.x / 2; w; w >>= 1) {
        if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0 && s[0] != 0.) atomicAdd(acc, s[0]);
}

__global__ void log_pso_row_sums_kernel(const double *V, const double *lb, const double *ub, unsigned n, unsigned dim, double *row_sum)
{
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    double s = 0.;
    for (unsigned j = 0; j < dim; ++j)
        if (ub[j] > lb[j]) s += fabs(V[static_cast<size_t>(p) * dim + j] / (ub[j] - lb[j]));
    row_sum[p] = s;
}

__global__ void log_pso_kernel(const double *lbfit, const double *row_sum, const double *dist_acc, unsigned n, unsigned dim, double gen,
                               double fevals, double *rows, unsigned *count, unsigned max_rows, unsigned row_len)
{
    __shared__ double smin[256], ssum[256];
    const unsigned row = *count;
    if (row >= max_rows) return;
    double mn = INFINITY, sm = 0.;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        mn = fmin(mn, lbfit[i]);
        sm += lbfit[i];
    }
    smin[threadIdx.x] = mn, ssum[threadIdx.x] = sm;
    __syncthreads();
    for (unsigned w = blockDim.x / 2; w; w >>= 1) {
        if (threadIdx.x < w) {
            smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + w]);
            ssum[threadIdx.x] += ssum[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // m_i = (m_{i-1} + s_i) / dim: particles more than ~40 places from the end no longer reach the double's last bit for dim >= 3,
        // but the recurrence is run in full as the reference runs it
        double mean_velocity = 0.;
        for (unsigned i = 0; i < n; ++i) mean_velocity = (mean_velocity + row_sum[i]) / static_cast<double>(dim);
        double *out = rows + static_cast<size_t>(row) * row_len;
        out[0] = gen;
        out[1] = fevals;
        out[2] = smin[0];
        out[3] = mean_velocity;
        out[4] = ssum[0] / static_cast<double>(n);
        out[5] = *dist_acc / (((static_cast<double>(n) - 1.) * static_cast<double>(n)) / 2.);
        *count = row + 1u;
    }
}

// sga.cpp:252-274: best of the children (plain <, from DBL_MAX), best of the parents (population::best_idx), their difference
__global__ void log_sga_kernel(const double *fp, const double *fc, unsigned n, unsigned verbosity, unsigned gen, double fevals, double *rows,
                               unsigned *count, unsigned max_rows, unsigned row_len)
{
    __shared__ double sp[256], sc[256];
    double bp = NAN, bc = 1.7976931348623157e308;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        const double p = fp[i], c = fc[i];
        if (!(p != p) && ((bp != bp) || p < bp)) bp = p; // less_than_f
        if (c < bc) bc = c;
    }
    sp[threadIdx.x] = bp, sc[threadIdx.x] = bc;
    __syncthreads();
    for (unsigned w = blockDim.x / 2; w; w >>= 1) {
        if (threadIdx.x < w) {
            const double p = sp[threadIdx.x + w];
            if (!(p != p) && ((sp[threadIdx.x] != sp[threadIdx.x]) || p < sp[threadIdx.x])) sp[threadIdx.x] = p;
            if (sc[threadIdx.x + w] < sc[threadIdx.x]) sc[threadIdx.x] = sc[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double improvement = sp[0] - sc[0];
        const bool due = ((gen % verbosity == 1u) && (verbosity > 1u)) || ((improvement > 0) && (verbosity == 1u));
        const unsigned row = *count;
        if (due && row < max_rows) {
            double *out = rows + static_cast<size_t>(row) * row_len;
            out[0] = gen, out[1] = fevals, out[2] = sp[0], out[3] = improvement;
            *count = row + 1u;
        }
    }
}

} // namespace

int log_sga_device(pgc_ctx *ctx, const double *d_f_parents, const double *d_f_children, unsigned n, unsigned gen, unsigned long long fevals,
                   cudaStream_t st)
{
    LogSink *L = tls_log;
    if (!L || !L->verbosity) return PGC_OK;
    log_sga_kernel<<<1, 256, 0, st>>>(d_f_parents, d_f_children, n, L->verbosity, gen, static_cast<double>(fevals), L->d_rows, L->d_count,
                                      L->max_rows, L->row_len);
    PGC_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

int log_ideal_device(pgc_ctx *ctx, const double *d_f, unsigned n, unsigned m, unsigned gen, unsigned long long fevals, cudaStream_t st)
{
    LogSink *L = tls_log;
    if (!L) return PGC_OK;
    log_ideal_kernel<<<1, 256, 0, st>>>(d_f, n, m, static_cast<double>(gen), static_cast<double>(fevals), L->d_rows, L->d_count, L->max_rows,
                                        L->row_len);
    PGC_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

int log_pso_device(pgc_ctx *ctx, const double *d_X, const double *d_V, const double *d_lbfit, const double *d_lb, const double *d_ub, unsigned n,
                   unsigned dim, unsigned gen, unsigned long long fevals, cudaStream_t st)
{
    LogSink *L = tls_log;
    if (!L) return PGC_OK;
    double *tmp = nullptr; // [n] row sums + the distance accumulator
    PGC_CUDA(cudaMallocAsync(&tmp, sizeof(double) * (static_cast<size_t>(n) + 1), st));
    cudaMemsetAsync(tmp + n, 0, sizeof(double), st);
    log_pso_row_sums_kernel<<<nblk(n, 256), 256, 0, st>>>(d_V, d_lb, d_ub, n, dim, tmp);
    if (n > 1u) log_pso_dist_kernel<<<n - 1u, 256, 0, st>>>(d_X, d_lb, d_ub, n, dim, tmp + n);
    log_pso_kernel<<<1, 256, 0, st>>>(d_lbfit, tmp, tmp + n, n, dim, static_cast<double>(gen), static_cast<double>(fevals), L->d_rows, L->d_count,
                                      L->max_rows, L->row_len);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(tmp, st);
    PGC_CUDA(e);
    ctx->launches.fetch_add(3, std::memory_order_relaxed);
    return PGC_OK;
}

} // namespace pgc
