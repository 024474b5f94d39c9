// capi.cu - the extern "C" boundary of libpgc.so (declared in include/pagmo_cuda/pgc.h).
// Contexts, problem handles, device/host batch evaluation, memory helpers, the FP64 ceiling probe.
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "pgc_internal.cuh"
#include "philox.cuh"

namespace pgc
{

static thread_local std::string g_last_error;

void set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) in `%s` at %s:%d", static_cast<int>(e), cudaGetErrorString(e), what, file, line);
    return e == cudaErrorMemoryAllocation ? PGC_ERR_OUT_OF_MEMORY : PGC_ERR_CUDA;
}

// Wait for the context's own streams.  Not cudaDeviceSynchronize(): a device-wide wait is an error while ANY stream of the device
// is being captured (another island's thread recording its generation graph, de.cu), whatever the capture mode.
int ctx_sync(pgc_ctx *ctx)
{
    if (ctx->stream) PGC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (cudaStream_t s : ctx->copy_stream)
        if (s) PGC_CUDA(cudaStreamSynchronize(s));
    return PGC_OK;
}

int ensure_scratch(pgc_ctx *ctx, size_t bytes)
{
    if (ctx->scratch_bytes >= bytes) return PGC_OK;
    // growing the scratch area must not race with kernels still using the old one
    if (int rc = ctx_sync(ctx)) return rc;
    if (ctx->scratch) PGC_CUDA(cudaFree(ctx->scratch));
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    PGC_CUDA(cudaMalloc(&ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
    return PGC_OK;
}

// ---- FP64 ceiling probe: 8 independent DFMA chains per thread, every SM saturated ------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
            r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
}

// Same probe through the legacy FP64 tensor path (mma.sync m8n8k4, "DMMA"): 8 independent accumulator chains
// per warp.  Only used to decide, with numbers, whether DMMA could beat the SIMT DFMA rotation on this part.
__global__ void __launch_bounds__(256) fp64_mma_peak_kernel(double *out, int iters, double a, double b)
{
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c0[i] = threadIdx.x + i;
        c1[i] = threadIdx.x - i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i])
                         : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int fp64_mma_peak(pgc_ctx *ctx, int iters, double *tflops)
{
    PGC_CUDA(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256;
    double *d = nullptr;
    PGC_CUDA(cudaMalloc(&d, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    PGC_CUDA(cudaEventCreate(&e0));
    PGC_CUDA(cudaEventCreate(&e1));
    fp64_mma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters / 4 + 1, 1e-3, 1e-3);
    PGC_CUDA(cudaEventRecord(e0, ctx->stream));
    fp64_mma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters, 1e-3, 1e-3);
    PGC_CUDA(cudaEventRecord(e1, ctx->stream));
    PGC_CUDA(cudaEventSynchronize(e1));
    ctx->launches.fetch_add(2, std::memory_order_relaxed);
    float ms = 0;
    PGC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 512.0 * 8.0 * static_cast<double>(iters) * blocks * (threads / 32);
    *tflops = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return PGC_OK;
}

// Probe kernel for design decisions: `dmma_warps` warps of each CTA run NACC independent DMMA chains,
// the remaining warps run 8 independent DFMA chains.  One CTA per SM (big dynamic smem keeps others out).
template <int NACC> __global__ void __launch_bounds__(512) fp64_mix_kernel(double *out, int iters, int dmma_warps, double a, double b)
{
    const int warp = threadIdx.x >> 5;
    double s = 0;
    if (warp < dmma_warps) {
        double c0[NACC], c1[NACC];
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            c0[i] = threadIdx.x + i;
            c1[i] = threadIdx.x - i;
        }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NACC; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0[i]), "+d"(c1[i])
                             : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
    } else {
        double r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
        // same FMA count per warp-iteration as the DMMA warps: NACC * 256 FMA / 32 lanes = 8*NACC per lane
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < NACC; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = fma(r[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s += r[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// tflops_out[0] = DMMA-warps' TFLOP/s, [1] = DFMA-warps' TFLOP/s, measured in the same launch
int fp64_mix_probe(pgc_ctx *ctx, int iters, int total_warps, int dmma_warps, double *tflops_out)
{
    PGC_CUDA(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count, threads = total_warps * 32;
    auto kern = fp64_mix_kernel<26>;
    const size_t smem = 120 * 1024; // one CTA per SM
    PGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    double *d = nullptr;
    PGC_CUDA(cudaMalloc(&d, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    PGC_CUDA(cudaEventCreate(&e0));
    PGC_CUDA(cudaEventCreate(&e1));
    kern<<<blocks, threads, smem, ctx->stream>>>(d, iters / 4 + 1, dmma_warps, 1e-3, 1e-3);
    PGC_CUDA(cudaEventRecord(e0, ctx->stream));
    kern<<<blocks, threads, smem, ctx->stream>>>(d, iters, dmma_warps, 1e-3, 1e-3);
    PGC_CUDA(cudaEventRecord(e1, ctx->stream));
    PGC_CUDA(cudaEventSynchronize(e1));
    PGC_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(2, std::memory_order_relaxed);
    float ms = 0;
    PGC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double per_warp = 2.0 * 256.0 * 26.0 * static_cast<double>(iters);
    tflops_out[0] = per_warp * dmma_warps * blocks / (ms * 1e-3) / 1e12;
    tflops_out[1] = per_warp * (total_warps - dmma_warps) * blocks / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return PGC_OK;
}

int fp64_peak(pgc_ctx *ctx, int iters, double *tflops)
{
    PGC_CUDA(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256;
    double *d = nullptr;
    PGC_CUDA(cudaMalloc(&d, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    PGC_CUDA(cudaEventCreate(&e0));
    PGC_CUDA(cudaEventCreate(&e1));
    fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters / 4 + 1, 0.999999, 1e-9); // warm-up
    PGC_CUDA(cudaEventRecord(e0, ctx->stream));
    fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters, 0.999999, 1e-9);
    PGC_CUDA(cudaEventRecord(e1, ctx->stream));
    PGC_CUDA(cudaEventSynchronize(e1));
    ctx->launches.fetch_add(2, std::memory_order_relaxed);
    float ms = 0;
    PGC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * static_cast<double>(iters) * blocks * threads;
    *tflops = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return PGC_OK;
}

int problem_eval_device(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t s)
{
    switch (p->desc.family) {
        case PGC_RASTRIGIN:
        case PGC_ACKLEY:
        case PGC_GRIEWANK:
        case PGC_SCHWEFEL:
        case PGC_ROSENBROCK: return simple_eval(p, d_dvs, n, d_fvs, s);
        case PGC_CEC2014: return cec2014_eval(p, d_dvs, n, d_fvs, s);
        case PGC_CEC2013: return cec2013_eval(p, d_dvs, n, d_fvs, s);
        case PGC_ZDT:
        case PGC_DTLZ:
        case PGC_WFG: return mo_eval(p, d_dvs, n, d_fvs, s);
        case PGC_LENNARD_JONES: return lj_eval(p, d_dvs, n, d_fvs, s);
        case PGC_HOCK_SCHITTKOWSKI_71:
        case PGC_LUKSAN_VLCEK1: return constrained_eval(p, d_dvs, n, d_fvs, s);
        case PGC_TRANSLATE:
        case PGC_DECOMPOSE:
        case PGC_UNCONSTRAIN: return meta_eval(p, d_dvs, n, d_fvs, s);
        default: set_error("family %d has no device evaluator in this build", p->desc.family); return PGC_ERR_UNSUPPORTED;
    }
}

} // namespace pgc

using namespace pgc;

namespace
{
struct DevBuf { // RAII device buffer for the host-vector convenience entry points
    void *p = nullptr;
    ~DevBuf() { cudaFree(p); }
    int alloc(size_t bytes)
    {
        PGC_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
        return PGC_OK;
    }
    template <class T> T *as() { return static_cast<T *>(p); }
};
} // namespace

static int indices_out(DevBuf &b, size_t cnt, size_t *dst)
{
    std::vector<unsigned> tmp(cnt ? cnt : 1);
    PGC_CUDA(cudaMemcpy(tmp.data(), b.p, 4 * cnt, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < cnt; ++i) dst[i] = tmp[i];
    return PGC_OK;
}

// The device generation operators treat every gene as continuous (the integer tails of sbx / polynomial mutation,
// genetic_operators.cpp:125-142,187-195, and sga's integer mutation are not built): refuse instead of silently running another algorithm.
#define PGC_NO_INTEGER_GENES(prob, who)                                                                                \
    do {                                                                                                               \
        if ((prob)->nix != 0) {                                                                                        \
            set_error("%s: '%s' has %zu integer decision variables; the device generation operators handle continuous " \
                      "genes only (no CPU fallback)",                                                                  \
                      who, (prob)->name.c_str(), (prob)->nix);                                                         \
            return PGC_ERR_UNSUPPORTED;                                                                                \
        }                                                                                                              \
    } while (0)

// The generation operators work on [n x nobj] fitness rows and rank by objectives alone.  The reference's algorithms refuse
// constrained problems too ("Non linear constraints detected in ... instance. de cannot deal with them", de.cpp:96-99 and the same
// check in every UDA of this path; gaco's penalty route is not built): wrap the problem in unconstrain (pgc_problem_unconstrain).
#define PGC_NO_CONSTRAINTS(prob, who)                                                                                  \
    do {                                                                                                               \
        if ((prob)->nec + (prob)->nic != 0) {                                                                          \
            set_error("Non linear constraints detected in %s instance. %s cannot deal with them", (prob)->name.c_str(), who); \
            return PGC_ERR_INVALID_ARGUMENT;                                                                           \
        }                                                                                                              \
    } while (0)

extern "C" {

const char *pgc_version(void) { return "pagmo2_b200 0.1.0 (sm_100a)"; }

const char *pgc_last_error(void) { return g_last_error.c_str(); }

int pgc_device_count(int *count)
{
    PGC_REQUIRE(count, "pgc_device_count: null output");
    *count = 0;
    PGC_CUDA(cudaGetDeviceCount(count));
    return PGC_OK;
}

int pgc_ctx_create(int device, pgc_ctx **out)
{
    PGC_REQUIRE(out, "pgc_ctx_create: null output");
    *out = nullptr;
    int count = 0;
    PGC_CUDA(cudaGetDeviceCount(&count));
    PGC_REQUIRE(device >= 0 && device < count, "pgc_ctx_create: device %d out of range (%d CUDA devices visible)", device, count);
    PGC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PGC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("pgc_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return PGC_ERR_UNSUPPORTED;
    }
    pgc_ctx *ctx = new (std::nothrow) pgc_ctx;
    if (!ctx) return PGC_ERR_OUT_OF_MEMORY;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    PGC_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    { // keep the stream-ordered pool warm: scratch of the level loops comes from cudaMallocAsync
        cudaMemPool_t pool;
        PGC_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = ~0ull;
        PGC_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    for (auto &s : ctx->copy_stream) PGC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (int i = 0; i < pgc_ctx::kRing; ++i) {
        PGC_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
        PGC_CUDA(cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming));
        PGC_CUDA(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
    }
    *out = ctx;
    return PGC_OK;
}

int pgc_ctx_set_sharers(pgc_ctx *ctx, int n)
{
    PGC_REQUIRE(ctx && n >= 1, "pgc_ctx_set_sharers: a context and a count >= 1 are needed");
    ctx->sharers = n;
    return PGC_OK;
}

int pgc_ctx_destroy(pgc_ctx *ctx)
{
    if (!ctx) return PGC_OK;
    cudaSetDevice(ctx->device);
    // no error reporting here: at process exit the runtime (and the thread-local error string) may already be gone
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (cudaStream_t s : ctx->copy_stream)
        if (s) cudaStreamSynchronize(s);
    (void)cudaGetLastError();
    for (int i = 0; i < pgc_ctx::kRing; ++i) {
        if (ctx->h_in[i]) cudaFreeHost(ctx->h_in[i]);
        if (ctx->h_out[i]) cudaFreeHost(ctx->h_out[i]);
        if (ctx->d_in[i]) cudaFree(ctx->d_in[i]);
        if (ctx->d_out[i]) cudaFree(ctx->d_out[i]);
        cudaEventDestroy(ctx->ev_in[i]);
        cudaEventDestroy(ctx->ev_k[i]);
        cudaEventDestroy(ctx->ev_out[i]);
    }
    if (ctx->scratch) cudaFree(ctx->scratch);
    delete ctx->copy_pool;
    for (auto &s : ctx->copy_stream) cudaStreamDestroy(s);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PGC_OK;
}

int pgc_ctx_device(const pgc_ctx *ctx, int *device)
{
    PGC_REQUIRE(ctx && device, "pgc_ctx_device: null argument");
    *device = ctx->device;
    return PGC_OK;
}

int pgc_ctx_stream(const pgc_ctx *ctx, void **stream)
{
    PGC_REQUIRE(ctx && stream, "pgc_ctx_stream: null argument");
    *stream = ctx->stream;
    return PGC_OK;
}

int pgc_ctx_synchronize(pgc_ctx *ctx)
{
    PGC_REQUIRE(ctx, "pgc_ctx_synchronize: null context");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return ctx_sync(ctx);
}

int pgc_ctx_launch_count(const pgc_ctx *ctx, uint64_t *count)
{
    PGC_REQUIRE(ctx && count, "pgc_ctx_launch_count: null argument");
    *count = ctx->launches.load(std::memory_order_relaxed);
    return PGC_OK;
}

int pgc_problem_create(pgc_ctx *ctx, const pgc_problem_desc *desc, pgc_problem **out)
{
    PGC_REQUIRE(ctx && desc && out, "pgc_problem_create: null argument");
    *out = nullptr;
    pgc_problem *p = new (std::nothrow) pgc_problem;
    if (!p) return PGC_ERR_OUT_OF_MEMORY;
    p->ctx = ctx;
    p->desc = *desc;
    p->desc.rotation = nullptr;
    p->desc.shift = nullptr;
    p->desc.shuffle = nullptr;
    int rc;
    switch (desc->family) {
        case PGC_RASTRIGIN:
        case PGC_ACKLEY:
        case PGC_GRIEWANK:
        case PGC_SCHWEFEL:
        case PGC_ROSENBROCK: rc = simple_create(p); break;
        case PGC_CEC2014: rc = cec2014_create(p, desc); break;
        case PGC_CEC2013: rc = cec2013_create(p, desc); break;
        case PGC_ZDT:
        case PGC_DTLZ:
        case PGC_WFG: rc = mo_create(p); break;
        case PGC_LENNARD_JONES: rc = lj_create(p); break;
        case PGC_HOCK_SCHITTKOWSKI_71:
        case PGC_LUKSAN_VLCEK1: rc = constrained_create(p); break;
        default:
            set_error("pgc_problem_create: family %d is not supported by this build (no CPU fallback)", desc->family);
            rc = PGC_ERR_UNSUPPORTED;
    }
    if (rc != PGC_OK) {
        cec2013_destroy(p);
        cec2014_destroy(p);
        delete p;
        return rc;
    }
    *out = p;
    return PGC_OK;
}

int pgc_problem_destroy(pgc_problem *p)
{
    if (!p) return PGC_OK;
    cudaSetDevice(p->ctx->device);
    // device-wide on purpose: the handle may outlive its context's streams at process exit (static destruction order in a host
    // application).  While another thread records a generation graph this returns an error, which is harmless here: the
    // cudaFree calls below order themselves after the work that uses the tables.
    if (cudaDeviceSynchronize() != cudaSuccess) (void)cudaGetLastError();
    p->work.clear(); // cached generation-loop workspaces (graphs, scratch buffers)
    if (p->inner) { // meta-problem: its tables are the wrapped problem's
        meta_destroy(p);
        delete p;
        return PGC_OK;
    }
    cec2013_destroy(p);
    cec2014_destroy(p);
    delete p;
    return PGC_OK;
}

int pgc_problem_translate(pgc_problem *inner, const double *translation, size_t len, pgc_problem **out)
{
    return meta_create(inner, PGC_TRANSLATE, translation, nullptr, len, 0, out);
}

int pgc_problem_unconstrain(pgc_problem *inner, int method, const double *weights, size_t len, pgc_problem **out)
{
    return meta_create(inner, PGC_UNCONSTRAIN, weights, nullptr, len, method, out);
}

int pgc_problem_decompose(pgc_problem *inner, const double *weight, const double *z, size_t len, int method, int adapt_ideal,
                          pgc_problem **out)
{
    if (adapt_ideal) {
        set_error("pgc_problem_decompose: ideal-point adaptation updates z after every single fitness call in call order "
                  "(decompose.cpp:143-149); the batch path does not reproduce that and there is no CPU fallback");
        return PGC_ERR_UNSUPPORTED;
    }
    return meta_create(inner, PGC_DECOMPOSE, weight, z, len, method, out);
}

int pgc_problem_nx(const pgc_problem *p, size_t *nx)
{
    PGC_REQUIRE(p && nx, "pgc_problem_nx: null argument");
    *nx = p->nx;
    return PGC_OK;
}

int pgc_problem_nix(const pgc_problem *p, size_t *nix)
{
    PGC_REQUIRE(p && nix, "pgc_problem_nix: null argument");
    *nix = p->nix;
    return PGC_OK;
}

int pgc_problem_set_strict(pgc_problem *p, int on)
{
    PGC_REQUIRE(p, "pgc_problem_set_strict: null problem");
    if (p->desc.family != PGC_CEC2013) {
        set_error("pgc_problem_set_strict: the strict summation order exists for the cec2013 suite only ('%s' given)", p->name.c_str());
        return PGC_ERR_UNSUPPORTED;
    }
    p->strict = on ? 1 : 0;
    ++p->config_epoch; // cached generation graphs were captured with the old switch
    return PGC_OK;
}

int pgc_problem_nobj(const pgc_problem *p, size_t *nobj)
{
    PGC_REQUIRE(p && nobj, "pgc_problem_nobj: null argument");
    *nobj = p->nobj;
    return PGC_OK;
}

int pgc_problem_nf(const pgc_problem *p, size_t *nf)
{
    PGC_REQUIRE(p && nf, "pgc_problem_nf: null argument");
    *nf = p->nf();
    return PGC_OK;
}

int pgc_problem_nec(const pgc_problem *p, size_t *nec)
{
    PGC_REQUIRE(p && nec, "pgc_problem_nec: null argument");
    *nec = p->nec;
    return PGC_OK;
}

int pgc_problem_nic(const pgc_problem *p, size_t *nic)
{
    PGC_REQUIRE(p && nic, "pgc_problem_nic: null argument");
    *nic = p->nic;
    return PGC_OK;
}

int pgc_problem_set_c_tol(pgc_problem *p, const double *c_tol, size_t len)
{
    // problem::set_c_tol, src/problem.cpp:620-644
    PGC_REQUIRE(p && (c_tol || len == 0), "pgc_problem_set_c_tol: null argument");
    const size_t nc = p->nec + p->nic;
    PGC_REQUIRE(len == nc, "The tolerance vector size should be: %zu, while a size of: %zu was detected.", nc, len);
    for (size_t i = 0; i < len; ++i) {
        PGC_REQUIRE(!std::isnan(c_tol[i]), "The tolerance vector has a NaN value at the index %zu", i);
        PGC_REQUIRE(!(c_tol[i] < 0.), "The tolerance vector has a negative value at the index %zu", i);
    }
    p->c_tol.assign(c_tol, c_tol + len);
    return PGC_OK;
}

int pgc_problem_c_tol(const pgc_problem *p, double *c_tol)
{
    PGC_REQUIRE(p && (c_tol || p->c_tol.empty()), "pgc_problem_c_tol: null argument");
    std::copy(p->c_tol.begin(), p->c_tol.end(), c_tol);
    return PGC_OK;
}

int pgc_feasibility_device(pgc_problem *p, const double *d_f, size_t n, uint8_t *d_feasible, void *stream)
{
    PGC_REQUIRE(p, "pgc_feasibility_device: null problem");
    PGC_REQUIRE(n == 0 || (d_f && d_feasible), "pgc_feasibility_device: null device buffer");
    PGC_CUDA(cudaSetDevice(p->ctx->device));
    return feasibility_rows(p, d_f, n, d_feasible, stream ? static_cast<cudaStream_t>(stream) : p->ctx->stream);
}

int pgc_problem_bounds(const pgc_problem *p, double *lb, double *ub)
{
    PGC_REQUIRE(p && lb && ub, "pgc_problem_bounds: null argument");
    std::copy(p->lb.begin(), p->lb.end(), lb);
    std::copy(p->ub.begin(), p->ub.end(), ub);
    return PGC_OK;
}

int pgc_problem_name(const pgc_problem *p, char *buf, size_t buflen)
{
    PGC_REQUIRE(p && buf && buflen, "pgc_problem_name: null argument");
    std::strncpy(buf, p->name.c_str(), buflen - 1);
    buf[buflen - 1] = 0;
    return PGC_OK;
}

int pgc_problem_work(const pgc_problem *p, double *flops, double *transc, double *bytes)
{
    PGC_REQUIRE(p, "pgc_problem_work: null problem");
    if (flops) *flops = p->flops_per_eval;
    if (transc) *transc = p->transc_per_eval;
    if (bytes) *bytes = 8.0 * static_cast<double>(p->nx + p->nf());
    return PGC_OK;
}

int pgc_eval_device(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, void *stream)
{
    PGC_REQUIRE(p, "pgc_eval_device: null problem");
    PGC_REQUIRE(n == 0 || (d_dvs && d_fvs), "pgc_eval_device: null device buffer");
    PGC_CUDA(cudaSetDevice(p->ctx->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : p->ctx->stream;
    return problem_eval_device(p, d_dvs, n, d_fvs, s);
}

// Host-vector path = the pagmo::bfe contract.  The batch is cut into chunks that cycle through a ring of
// pinned+device staging buffers: memcpy(host->pinned) | H2D | kernels | D2H | memcpy(pinned->host) overlap
// across chunks on three streams.  PCIe-bound by construction.
int pgc_eval_host(pgc_problem *p, const double *dvs, size_t n, double *fvs)
{
    PGC_REQUIRE(p, "pgc_eval_host: null problem");
    if (n == 0) return PGC_OK;
    PGC_REQUIRE(dvs && fvs, "pgc_eval_host: null host buffer");
    pgc_ctx *ctx = p->ctx;
    PGC_CUDA(cudaSetDevice(ctx->device));
    const size_t nx = p->nx, nf = p->nf();
    // chunk: ~32 MiB of decision vectors, a multiple of 16 individuals
    size_t chunk = (32u << 20) / (nx * sizeof(double));
    chunk = std::max<size_t>(16, chunk / 16 * 16);
    if (chunk > n) chunk = (n + 15) / 16 * 16;
    const size_t in_bytes = chunk * nx * sizeof(double), out_bytes = chunk * nf * sizeof(double);
    if (ctx->ring_in_bytes < in_bytes || ctx->ring_out_bytes < out_bytes) {
        if (int rc = ctx_sync(ctx)) return rc;
        for (int i = 0; i < pgc_ctx::kRing; ++i) {
            if (ctx->h_in[i]) cudaFreeHost(ctx->h_in[i]);
            if (ctx->h_out[i]) cudaFreeHost(ctx->h_out[i]);
            if (ctx->d_in[i]) cudaFree(ctx->d_in[i]);
            if (ctx->d_out[i]) cudaFree(ctx->d_out[i]);
            ctx->h_in[i] = ctx->h_out[i] = ctx->d_in[i] = ctx->d_out[i] = nullptr;
        }
        ctx->ring_in_bytes = ctx->ring_out_bytes = 0;
        for (int i = 0; i < pgc_ctx::kRing; ++i) {
            PGC_CUDA(cudaMallocHost(&ctx->h_in[i], in_bytes));
            PGC_CUDA(cudaMallocHost(&ctx->h_out[i], out_bytes));
            PGC_CUDA(cudaMalloc(&ctx->d_in[i], in_bytes));
            PGC_CUDA(cudaMalloc(&ctx->d_out[i], out_bytes));
        }
        ctx->ring_in_bytes = in_bytes;
        ctx->ring_out_bytes = out_bytes;
    }
    cudaPointerAttributes attr;
    const bool in_pinned = cudaPointerGetAttributes(&attr, dvs) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    const bool out_pinned = cudaPointerGetAttributes(&attr, fvs) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError(); // clear a possible "invalid value" from probing pageable memory
    if (!in_pinned && !ctx->copy_pool && n * nx * sizeof(double) >= (64u << 20)) {
        // PGC_COPY_THREADS: helpers for staging pageable input (default: up to 7, leaving cores for the other devices' contexts)
        const char *env = std::getenv("PGC_COPY_THREADS");
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        unsigned want = env ? static_cast<unsigned>(std::max(0, std::atoi(env))) : std::min(7u, hw > 2u ? hw / 2u - 1u : 0u);
        ctx->copy_pool = new pgc::CopyPool(want);
    }

    const size_t nchunks = (n + chunk - 1) / chunk;
    cudaStream_t s_in = ctx->copy_stream[0], s_out = ctx->copy_stream[1], s_k = ctx->stream;
    struct Pending { size_t first, count; };
    Pending pend[pgc_ctx::kRing];
    auto drain = [&](int slot) -> int { // finish the D2H of the chunk living in `slot`
        PGC_CUDA(cudaEventSynchronize(ctx->ev_out[slot]));
        if (!out_pinned)
            std::memcpy(fvs + pend[slot].first * nf, ctx->h_out[slot], pend[slot].count * nf * sizeof(double));
        return PGC_OK;
    };
    for (size_t c = 0; c < nchunks; ++c) {
        const int slot = static_cast<int>(c % pgc_ctx::kRing);
        if (c >= static_cast<size_t>(pgc_ctx::kRing)) {
            int rc = drain(slot);
            if (rc != PGC_OK) return rc;
        }
        const size_t first = c * chunk, count = std::min(chunk, n - first);
        pend[slot] = {first, count};
        const double *src = dvs + first * nx;
        if (!in_pinned) {
            if (ctx->copy_pool) ctx->copy_pool->copy(ctx->h_in[slot], src, count * nx * sizeof(double));
            else std::memcpy(ctx->h_in[slot], src, count * nx * sizeof(double));
            src = static_cast<const double *>(ctx->h_in[slot]);
        }
        PGC_CUDA(cudaMemcpyAsync(ctx->d_in[slot], src, count * nx * sizeof(double), cudaMemcpyHostToDevice, s_in));
        PGC_CUDA(cudaEventRecord(ctx->ev_in[slot], s_in));
        PGC_CUDA(cudaStreamWaitEvent(s_k, ctx->ev_in[slot], 0));
        int rc = problem_eval_device(p, static_cast<const double *>(ctx->d_in[slot]), count,
                                     static_cast<double *>(ctx->d_out[slot]), s_k);
        if (rc != PGC_OK) return rc;
        PGC_CUDA(cudaEventRecord(ctx->ev_k[slot], s_k));
        PGC_CUDA(cudaStreamWaitEvent(s_out, ctx->ev_k[slot], 0));
        double *dst = out_pinned ? fvs + first * nf : static_cast<double *>(ctx->h_out[slot]);
        PGC_CUDA(cudaMemcpyAsync(dst, ctx->d_out[slot], count * nf * sizeof(double), cudaMemcpyDeviceToHost, s_out));
        PGC_CUDA(cudaEventRecord(ctx->ev_out[slot], s_out));
    }
    const size_t tail = std::min<size_t>(nchunks, pgc_ctx::kRing);
    for (size_t c = nchunks - tail; c < nchunks; ++c) {
        int rc = drain(static_cast<int>(c % pgc_ctx::kRing));
        if (rc != PGC_OK) return rc;
    }
    return PGC_OK;
}

int pgc_debug_cec2014_phase_cycles(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, uint64_t *out7)
{
    PGC_REQUIRE(p && d_dvs && d_fvs && out7, "pgc_debug_cec2014_phase_cycles: null argument");
    PGC_REQUIRE(p->desc.family == PGC_CEC2014, "pgc_debug_cec2014_phase_cycles: not a cec2014 problem");
    PGC_CUDA(cudaSetDevice(p->ctx->device));
    unsigned long long tmp[8] = {};
    int rc = cec2014_phase_cycles(p, d_dvs, n, d_fvs, tmp);
    for (int i = 0; i < 7; ++i) out7[i] = tmp[i];
    return rc;
}

// ---- multi-objective utilities -----------------------------------------------------------------------------------
int pgc_fnds_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t m, uint32_t *d_rank, uint32_t *d_dom_count,
                    uint32_t *d_front_idx, uint32_t *d_front_off, uint32_t *nfronts, void *stream)
{
    PGC_REQUIRE(ctx && d_f && d_rank && d_front_idx && d_front_off && nfronts, "pgc_fnds_device: null argument");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return fnds_device(ctx, d_f, n, m, d_rank, d_dom_count, d_front_idx, d_front_off, nfronts,
                       stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int pgc_crowding_fronts_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t m, const uint32_t *d_front_idx,
                               const uint32_t *d_front_off, uint32_t nfronts, int small_front_rule, double *d_cd, void *stream)
{
    PGC_REQUIRE(ctx && d_f && d_front_idx && d_front_off && d_cd, "pgc_crowding_fronts_device: null argument");
    PGC_REQUIRE(m >= 2, "Points in the non dominated front must contain at least two objectives: %zu detected.", m);
    PGC_CUDA(cudaSetDevice(ctx->device));
    return crowding_device(ctx, d_f, n, m, d_front_idx, d_front_off, nfronts, small_front_rule, d_cd,
                           stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int pgc_select_best_N_mo_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t m, size_t N, uint32_t *d_out, uint32_t *nout,
                                void *stream)
{
    PGC_REQUIRE(ctx && nout && (n == 0 || (d_f && d_out)), "pgc_select_best_N_mo_device: null argument");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return select_best_device(ctx, d_f, n, m, N, d_out, nout, stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int pgc_fnds_host(pgc_ctx *ctx, const double *f, size_t n, size_t m, size_t *rank, size_t *dom_count, size_t *front_idx,
                  size_t *front_off, size_t *nfronts)
{
    PGC_REQUIRE(ctx && nfronts, "pgc_fnds_host: null argument");
    PGC_REQUIRE(n >= 2, "At least two points are needed for fast_non_dominated_sorting: %zu detected.", n);
    PGC_REQUIRE(f || m == 0, "pgc_fnds_host: null fitness buffer");
    PGC_CUDA(cudaSetDevice(ctx->device));
    DevBuf df, dr, dc, doo, dfo;
    int rc;
    if ((rc = df.alloc(sizeof(double) * n * m)) || (rc = dr.alloc(4 * n)) || (rc = dc.alloc(4 * n)) || (rc = doo.alloc(4 * n))
        || (rc = dfo.alloc(4 * (n + 1))))
        return rc;
    if (m) PGC_CUDA(cudaMemcpyAsync(df.p, f, sizeof(double) * n * m, cudaMemcpyHostToDevice, ctx->stream));
    unsigned nf = 0;
    if ((rc = fnds_device(ctx, df.as<double>(), n, m, dr.as<unsigned>(), dc.as<unsigned>(), doo.as<unsigned>(), dfo.as<unsigned>(), &nf,
                          ctx->stream)))
        return rc;
    std::vector<unsigned> tmp(n + 1);
    auto fetch = [&](DevBuf &b, size_t cnt, size_t *dst) -> int {
        if (!dst) return PGC_OK;
        PGC_CUDA(cudaMemcpy(tmp.data(), b.p, 4 * cnt, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < cnt; ++i) dst[i] = tmp[i];
        return PGC_OK;
    };
    if ((rc = fetch(dr, n, rank)) || (rc = fetch(dc, n, dom_count)) || (rc = fetch(doo, n, front_idx)) || (rc = fetch(dfo, nf + 1, front_off)))
        return rc;
    *nfronts = nf;
    return PGC_OK;
}

int pgc_crowding_distance_host(pgc_ctx *ctx, const double *f, size_t n, size_t m, double *out)
{
    PGC_REQUIRE(ctx && out, "pgc_crowding_distance_host: null argument");
    PGC_REQUIRE(n >= 2, "A non dominated front must contain at least two points: %zu detected.", n);            // :283-286
    PGC_REQUIRE(m >= 2, "Points in the non dominated front must contain at least two objectives: %zu detected.", m); // :289-292
    PGC_REQUIRE(f, "pgc_crowding_distance_host: null fitness buffer");
    PGC_CUDA(cudaSetDevice(ctx->device));
    DevBuf df, dcd;
    int rc;
    if ((rc = df.alloc(sizeof(double) * n * m)) || (rc = dcd.alloc(sizeof(double) * n))) return rc;
    PGC_CUDA(cudaMemcpyAsync(df.p, f, sizeof(double) * n * m, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = crowding_device(ctx, df.as<double>(), n, m, nullptr, nullptr, 1, 0, dcd.as<double>(), ctx->stream))) return rc;
    PGC_CUDA(cudaMemcpy(out, dcd.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return PGC_OK;
}

int pgc_select_best_N_mo_host(pgc_ctx *ctx, const double *f, size_t n, size_t m, size_t N, size_t *out, size_t *nout)
{
    PGC_REQUIRE(ctx && nout, "pgc_select_best_N_mo_host: null argument");
    *nout = 0;
    if (N == 0 || n == 0) return PGC_OK;
    PGC_REQUIRE(out && (f || m == 0), "pgc_select_best_N_mo_host: null buffer");
    PGC_CUDA(cudaSetDevice(ctx->device));
    DevBuf df, dout;
    int rc;
    if ((rc = df.alloc(sizeof(double) * n * m)) || (rc = dout.alloc(4 * n))) return rc;
    if (m) PGC_CUDA(cudaMemcpyAsync(df.p, f, sizeof(double) * n * m, cudaMemcpyHostToDevice, ctx->stream));
    unsigned cnt = 0;
    if ((rc = select_best_device(ctx, df.as<double>(), n, m, N, dout.as<unsigned>(), &cnt, ctx->stream))) return rc;
    PGC_CUDA(cudaStreamSynchronize(ctx->stream));
    if ((rc = indices_out(dout, cnt, out))) return rc;
    *nout = cnt;
    return PGC_OK;
}

int pgc_sort_population_mo_host(pgc_ctx *ctx, const double *f, size_t n, size_t m, size_t *out)
{
    PGC_REQUIRE(ctx, "pgc_sort_population_mo_host: null context");
    if (n == 0) return PGC_OK;
    PGC_REQUIRE(out && (f || m == 0), "pgc_sort_population_mo_host: null buffer");
    PGC_CUDA(cudaSetDevice(ctx->device));
    DevBuf df, dout;
    int rc;
    if ((rc = df.alloc(sizeof(double) * n * m)) || (rc = dout.alloc(4 * n))) return rc;
    if (m) PGC_CUDA(cudaMemcpyAsync(df.p, f, sizeof(double) * n * m, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = sort_population_device(ctx, df.as<double>(), n, m, dout.as<unsigned>(), ctx->stream))) return rc;
    PGC_CUDA(cudaStreamSynchronize(ctx->stream));
    return indices_out(dout, n, out);
}

// ---- generation operators ------------------------------------------------------------------------------------------
int pgc_philox_u01(uint64_t seed, uint32_t tag, uint32_t generation, uint32_t index, uint32_t slot, double *out)
{
    PGC_REQUIRE(out, "pgc_philox_u01: null output");
    *out = philox_u01(seed, tag, generation, index, slot);
    return PGC_OK;
}

int pgc_philox_permutation_device(pgc_ctx *ctx, size_t n, uint64_t seed, uint32_t tag, uint32_t generation, uint32_t *d_perm, void *stream)
{
    PGC_REQUIRE(ctx && (d_perm || n == 0), "pgc_philox_permutation_device: null argument");
    if (n == 0) return PGC_OK;
    PGC_CUDA(cudaSetDevice(ctx->device));
    return philox_permutation_device(ctx, static_cast<unsigned>(n), seed, tag, generation, d_perm,
                                     stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int pgc_nsga2_variation_device(pgc_ctx *ctx, const double *d_x, const uint32_t *d_rank, const double *d_cd, size_t NP, size_t nx,
                               const double *d_lb, const double *d_ub, const uint32_t *d_shuffle1, const uint32_t *d_shuffle2,
                               double cr, double eta_c, double m, double eta_m, uint64_t seed, uint32_t generation,
                               double *d_children, void *stream)
{
    PGC_REQUIRE(ctx && d_x && d_rank && d_cd && d_lb && d_ub && d_shuffle1 && d_shuffle2 && d_children, "pgc_nsga2_variation_device: null argument");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return nsga2_variation_device(ctx, d_x, d_rank, d_cd, static_cast<unsigned>(NP), static_cast<unsigned>(nx), d_lb, d_ub, d_shuffle1,
                                  d_shuffle2, cr, eta_c, m, eta_m, seed, generation, d_children,
                                  stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int pgc_nsga2_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t NP, unsigned gens, double cr, double eta_c, double m,
                            double eta_m, uint64_t seed, uint32_t first_generation, void *stream)
{
    PGC_REQUIRE(prob && d_x && d_f, "pgc_nsga2_evolve_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_nsga2_evolve_device");
    // integer alleles (the last nix genes, e.g. ZDT5) are handled: two-point crossover + uniform integer mutation, genetic_operators.cpp:125-137, :187-195
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    return nsga2_evolve_device(prob, d_x, d_f, static_cast<unsigned>(NP), gens, cr, eta_c, m, eta_m, seed, first_generation,
                               problem_eval_device, stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream);
}

int pgc_pso_evolve_device(pgc_problem *prob, double *d_x, double *d_f, double *d_v, double *d_xcur, size_t n, unsigned gens, double omega,
                          double eta1, double eta2, double max_vel, unsigned variant, unsigned neighb_type, unsigned neighb_param,
                          uint64_t seed, uint32_t first_generation, void *stream)
{
    PGC_REQUIRE(prob && d_x && d_f, "pgc_pso_evolve_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_pso_evolve_device");
    PGC_NO_INTEGER_GENES(prob, "pgc_pso_evolve_device");
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    return pso_evolve_device(prob, d_x, d_f, d_v, d_xcur, static_cast<unsigned>(n), gens, omega, eta1, eta2, max_vel, variant, neighb_type,
                             neighb_param, seed, first_generation, problem_eval_device,
                             stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream);
}

int pgc_pso_shard_step_device(pgc_problem *prob, double *d_X, double *d_V, double *d_lbX_ext, double *d_lbfit_ext, size_t n_loc, unsigned radius,
                              unsigned index_offset, double omega, double eta1, double eta2, double max_vel, unsigned variant, uint64_t seed,
                              uint32_t generation, int init_velocity, void *stream)
{
    PGC_REQUIRE(prob && d_V && (init_velocity || (d_X && d_lbX_ext && d_lbfit_ext)), "pgc_pso_shard_step_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_pso_shard_step_device");
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    return pso_shard_step_device(prob, d_X, d_V, d_lbX_ext, d_lbfit_ext, static_cast<unsigned>(n_loc), radius, index_offset, omega, eta1, eta2,
                                 max_vel, variant, seed, generation, init_velocity, problem_eval_device,
                                 stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream);
}

int pgc_pso_shard_step_gbest_device(pgc_problem *prob, double *d_X, double *d_V, double *d_lbX_ext, double *d_lbfit_ext, size_t n_loc,
                                    unsigned index_offset, double omega, double eta1, double eta2, double max_vel, unsigned variant,
                                    uint64_t seed, uint32_t generation, int init_velocity, double *d_cand, void *stream)
{
    PGC_REQUIRE(prob && d_V && d_cand && (init_velocity || (d_X && d_lbX_ext && d_lbfit_ext)), "pgc_pso_shard_step_gbest_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_pso_shard_step_gbest_device");
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    return pso_shard_step_device(prob, d_X, d_V, d_lbX_ext, d_lbfit_ext, static_cast<unsigned>(n_loc), 1u, index_offset, omega, eta1, eta2,
                                 max_vel, variant, seed, generation, init_velocity, problem_eval_device,
                                 stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream, d_cand);
}

int pgc_de_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t NP, unsigned gens, unsigned algo, unsigned variant,
                         unsigned variant_adptv, double F, double CR, const uint32_t *allowed_variants, unsigned n_allowed, double ftol,
                         double xtol, double *d_F, double *d_CR, uint32_t *d_variant, uint64_t seed, uint32_t first_generation,
                         unsigned *gens_done, void *stream)
{
    PGC_REQUIRE(prob && d_x && d_f, "pgc_de_evolve_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_de_evolve_device");
    PGC_NO_INTEGER_GENES(prob, "pgc_de_evolve_device");
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    return de_evolve_device(prob, d_x, d_f, static_cast<unsigned>(NP), gens, algo, variant, variant_adptv, F, CR, allowed_variants, n_allowed,
                            ftol, xtol, d_F, d_CR, d_variant, seed, first_generation, gens_done, problem_eval_device,
                            stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream);
}

int pgc_moead_gen_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t n, unsigned gens, const double *weights, const uint32_t *neigh,
                                unsigned T, int decomposition, double CR, double F, double eta_m, double realb, unsigned limit,
                                int preserve_diversity, uint64_t seed, uint32_t first_generation, void *stream)
{
    PGC_REQUIRE(prob && d_x && d_f, "pgc_moead_gen_evolve_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_moead_gen_evolve_device");
    PGC_NO_INTEGER_GENES(prob, "pgc_moead_gen_evolve_device");
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    return moead_gen_evolve_device(prob, d_x, d_f, static_cast<unsigned>(n), gens, weights, neigh, T, decomposition, CR, F, eta_m, realb, limit,
                                   preserve_diversity, seed, first_generation, problem_eval_device,
                                   stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream);
}

int pgc_nspso_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t n, unsigned gens, double omega, double c1, double c2, double chi,
                            double v_coeff, unsigned leader_selection_range, unsigned diversity, uint64_t seed, uint32_t first_generation,
                            double *d_vel, double *d_best_x, double *d_best_f, void *stream)
{
    PGC_REQUIRE(prob && d_x && d_f, "pgc_nspso_evolve_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_nspso_evolve_device");
    PGC_NO_INTEGER_GENES(prob, "pgc_nspso_evolve_device");
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    return nspso_evolve_device(prob, d_x, d_f, static_cast<unsigned>(n), gens, omega, c1, c2, chi, v_coeff, leader_selection_range, diversity, seed,
                               first_generation, d_vel, d_best_x, d_best_f, problem_eval_device,
                               stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream);
}

int pgc_gaco_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t n, unsigned gens, unsigned ker, double q, double oracle,
                           double acc, unsigned threshold, unsigned n_gen_mark, unsigned impstop, unsigned evalstop, double focus, uint64_t seed,
                           uint32_t first_generation, pgc_gaco_state *state, unsigned *gens_done, void *stream)
{
    PGC_REQUIRE(prob && (n == 0 || (d_x && d_f)), "pgc_gaco_evolve_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_gaco_evolve_device");
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    pgc_gaco_state local{};
    pgc_gaco_state *s = state ? state : &local;
    if (!s->initialized) { // a freshly constructed pagmo::gaco, gaco.cpp:60-64
        s->oracle = oracle, s->q = q;
        s->n_evalstop = 1, s->n_impstop = 1, s->gen_mark = 1, s->fevals = 0;
        s->initialized = 1;
    }
    return gaco_evolve_device(prob, d_x, d_f, static_cast<unsigned>(n), gens, ker, acc, threshold, n_gen_mark, impstop, evalstop, focus, seed,
                              first_generation, s, gens_done, problem_eval_device,
                              stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream);
}

int pgc_maco_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t n, unsigned gens, unsigned ker, double q, unsigned threshold,
                           unsigned n_gen_mark, unsigned evalstop, double focus, uint64_t seed, uint32_t first_generation,
                           pgc_maco_state *state, unsigned *gens_done, void *stream)
{
    PGC_REQUIRE(prob && (n == 0 || (d_x && d_f)), "pgc_maco_evolve_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_maco_evolve_device");
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    pgc_maco_state local{};
    pgc_maco_state *s = state ? state : &local;
    if (!s->initialized) { // a freshly constructed pagmo::maco, maco.cpp:60-63
        s->q = q;
        s->n_evalstop = 0, s->gen_mark = 1;
        s->initialized = 1;
    }
    return maco_evolve_device(prob, d_x, d_f, static_cast<unsigned>(n), gens, ker, threshold, n_gen_mark, evalstop, focus, seed, first_generation,
                              s, gens_done, problem_eval_device, stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream);
}

int pgc_xnes_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t lambda, unsigned gens, double eta_mu, double eta_sigma, double eta_b,
                           double sigma0, double ftol, double xtol, int force_bounds, uint64_t seed, uint32_t first_generation,
                           unsigned *gens_done, double *sigma_out, void *stream)
{
    PGC_REQUIRE(prob && (lambda == 0 || (d_x && d_f)), "pgc_xnes_evolve_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_xnes_evolve_device");
    PGC_NO_INTEGER_GENES(prob, "pgc_xnes_evolve_device");
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    return xnes_evolve_device(prob, d_x, d_f, lambda, gens, eta_mu, eta_sigma, eta_b, sigma0, ftol, xtol, force_bounds, seed, first_generation,
                              gens_done, sigma_out, problem_eval_device, stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream);
}

int pgc_hv_device(pgc_ctx *ctx, const double *d_points, size_t n, size_t m, const double *r_point, int compute, double *d_out, void *stream)
{
    PGC_REQUIRE(ctx && r_point && d_out && (d_points || n == 0), "pgc_hv_device: null argument");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return hv_device(ctx, d_points, n, m, r_point, compute, d_out, stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int pgc_hv_fpras_host(pgc_ctx *ctx, const double *points, size_t n, size_t m, const double *r_point, double eps, double delta, uint64_t seed,
                      double *hv)
{
    PGC_REQUIRE(ctx && r_point && hv && (points || n == 0), "pgc_hv_fpras_host: null argument");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return hv_fpras_host(ctx, points, n, m, r_point, eps, delta, seed, hv);
}

int pgc_hv_approx_extreme_host(pgc_ctx *ctx, const double *points, size_t n, size_t m, const double *r_point, int greatest, int use_exact,
                               unsigned trivial_subcase_size, double eps, double delta, double delta_multiplier, double alpha,
                               double initial_delta_coeff, double gamma, uint64_t seed, size_t *idx)
{
    PGC_REQUIRE(ctx && r_point && idx && points, "pgc_hv_approx_extreme_host: null argument");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return hv_approx_extreme_host(ctx, points, n, m, r_point, greatest, use_exact, trivial_subcase_size, eps, delta, delta_multiplier, alpha,
                                  initial_delta_coeff, gamma, seed, idx);
}

static int hv_host(pgc_ctx *ctx, const double *points, size_t n, size_t m, const double *r, int compute, double *out)
{
    PGC_REQUIRE(ctx && r && out && (points || n == 0), "pgc_hv_*_host: null argument");
    PGC_CUDA(cudaSetDevice(ctx->device));
    double *d_f = nullptr, *d_o = nullptr;
    PGC_CUDA(cudaMalloc(&d_f, (n * m ? n * m : 1) * sizeof(double)));
    cudaError_t e = cudaMalloc(&d_o, (n ? n : 1) * sizeof(double));
    int rc = PGC_OK;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_f, points, n * m * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) rc = hv_device(ctx, d_f, n, m, r, compute, d_o, ctx->stream);
    if (e == cudaSuccess && rc == PGC_OK)
        e = cudaMemcpyAsync(out, d_o, (compute ? 1 : n) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && rc == PGC_OK) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_f);
    cudaFree(d_o);
    if (e != cudaSuccess) return cuda_fail(e, "pgc_hv_*_host", __FILE__, __LINE__);
    return rc;
}

int pgc_hv_compute_host(pgc_ctx *ctx, const double *points, size_t n, size_t m, const double *r_point, double *hv)
{
    return hv_host(ctx, points, n, m, r_point, 1, hv);
}

int pgc_hv_contributions_host(pgc_ctx *ctx, const double *points, size_t n, size_t m, const double *r_point, double *out)
{
    return hv_host(ctx, points, n, m, r_point, 0, out);
}

int pgc_cmaes_sample_device(pgc_ctx *ctx, const double *d_mean, const double *d_bd, double sigma, size_t lambda, size_t D, uint64_t seed,
                            uint32_t generation, double *d_z, double *d_x, void *stream)
{
    PGC_REQUIRE(ctx && d_mean && d_bd && (d_x || lambda == 0), "pgc_cmaes_sample_device: null argument");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return cmaes_sample_device(ctx, d_mean, d_bd, sigma, lambda, D, seed, generation, d_z, d_x, stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int pgc_weighted_gram_device(pgc_ctx *ctx, const double *d_rows, const uint32_t *d_idx, const double *d_center, const double *d_w, size_t k,
                             size_t D, double scale_div, double *d_out, void *stream)
{
    PGC_REQUIRE(ctx && d_rows && d_w && d_out, "pgc_weighted_gram_device: null argument");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return weighted_gram_device(ctx, d_rows, d_idx, d_center, d_w, k, D, scale_div, d_out, stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int pgc_weighted_mean_device(pgc_ctx *ctx, const double *d_rows, const uint32_t *d_idx, const double *d_w, size_t k, size_t D, double *d_out,
                             void *stream)
{
    PGC_REQUIRE(ctx && d_rows && d_w && d_out, "pgc_weighted_mean_device: null argument");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return weighted_mean_device(ctx, d_rows, d_idx, d_w, k, D, d_out, stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int pgc_sga_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t NP, unsigned gens, double cr, double eta_c, double m,
                          double param_m, unsigned param_s, unsigned crossover, unsigned mutation, unsigned selection, uint64_t seed,
                          uint32_t first_generation, void *stream)
{
    PGC_REQUIRE(prob && d_x && d_f, "pgc_sga_evolve_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_sga_evolve_device");
    PGC_NO_INTEGER_GENES(prob, "pgc_sga_evolve_device");
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    return sga_evolve_device(prob, d_x, d_f, static_cast<unsigned>(NP), gens, cr, eta_c, m, param_m, param_s, crossover, mutation, selection,
                             seed, first_generation, problem_eval_device, stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream);
}

int pgc_cmaes_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t lambda, unsigned gens, double cc, double cs, double c1, double cmu,
                            double sigma0, double ftol, double xtol, int force_bounds, uint64_t seed, uint32_t first_generation,
                            unsigned *gens_done, double *sigma_out, void *stream)
{
    PGC_REQUIRE(prob && d_x && d_f, "pgc_cmaes_evolve_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_cmaes_evolve_device");
    PGC_NO_INTEGER_GENES(prob, "pgc_cmaes_evolve_device");
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    return cmaes_evolve_device(prob, d_x, d_f, lambda, gens, cc, cs, c1, cmu, sigma0, ftol, xtol, force_bounds, seed, first_generation, gens_done,
                               sigma_out, problem_eval_device, stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream);
}

int pgc_algo_defaults(int algo, unsigned gens, uint64_t seed, pgc_algo_desc *out)
{
    PGC_REQUIRE(out, "pgc_algo_defaults: null argument");
    pgc_algo_desc d{};
    d.algo = algo;
    d.gens = gens;
    d.seed = seed;
    d.ftol = d.xtol = 1e-6;
    d.variant_adptv = 1;
    switch (algo) {
        case PGC_ALGO_DE: d.F = 0.8, d.CR = 0.9, d.variant = 2; break;            // de.hpp:119
        case PGC_ALGO_SADE: d.variant = 2; break;                                 // sade.hpp:138
        case PGC_ALGO_DE1220: {                                                   // de1220.hpp:158, :52-53
            static const uint32_t allowed[8] = {2u, 3u, 7u, 10u, 13u, 14u, 15u, 16u};
            d.n_allowed = 8;
            for (int i = 0; i < 8; ++i) d.allowed_variants[i] = allowed[i];
            break;
        }
        case PGC_ALGO_PSO_GEN: // pso_gen.hpp:127
            d.omega = 0.7298, d.eta1 = d.eta2 = 2.05, d.max_vel = 0.5, d.variant = 5, d.neighb_type = 2, d.neighb_param = 4;
            break;
        case PGC_ALGO_NSGA2: d.cr = 0.95, d.eta_c = 10., d.m = 0.01, d.eta_m = 50.; break; // nsga2.hpp:103
        case PGC_ALGO_SGA: // sga.hpp:166: exponential crossover, polynomial mutation, tournament selection
            d.cr = 0.9, d.eta_c = 1., d.m = 0.02, d.param_m = 1., d.param_s = 2, d.crossover = 0, d.mutation = 2, d.selection = 0;
            break;
        case PGC_ALGO_CMAES: d.cma_cc = d.cma_cs = d.cma_c1 = d.cma_cmu = -1., d.sigma0 = 0.5; break; // cmaes.hpp:110
        case PGC_ALGO_XNES: d.cma_cc = d.cma_cs = d.cma_c1 = -1., d.sigma0 = -1.; break; // xnes.hpp:107-108 (eta_mu, eta_sigma, eta_b, sigma0)
        case PGC_ALGO_NSPSO: // nspso.hpp:59-62
            d.omega = 0.6, d.nspso_c1 = 2.0, d.nspso_c2 = 2.0, d.nspso_chi = 1.0, d.nspso_v_coeff = 0.5, d.leader_selection_range = 60, d.diversity = 0;
            break;
        default: set_error("pgc_algo_defaults: unknown algorithm %d", algo); return PGC_ERR_INVALID_ARGUMENT;
    }
    *out = d;
    return PGC_OK;
}

int pgc_algo_evolve_device(pgc_problem *prob, const pgc_algo_desc *a, double *d_x, double *d_f, size_t n, uint32_t first_generation,
                           unsigned *gens_done, void *stream)
{
    PGC_REQUIRE(prob && a && d_x && d_f, "pgc_algo_evolve_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_algo_evolve_device");
    if (gens_done) *gens_done = a->gens;
    switch (a->algo) {
        case PGC_ALGO_DE:
        case PGC_ALGO_SADE:
        case PGC_ALGO_DE1220:
            return pgc_de_evolve_device(prob, d_x, d_f, n, a->gens, static_cast<unsigned>(a->algo - PGC_ALGO_DE), a->variant, a->variant_adptv,
                                        a->F, a->CR, a->allowed_variants, a->n_allowed, a->ftol, a->xtol, nullptr, nullptr, nullptr, a->seed,
                                        first_generation, gens_done, stream);
        case PGC_ALGO_PSO_GEN:
            return pgc_pso_evolve_device(prob, d_x, d_f, nullptr, nullptr, n, a->gens, a->omega, a->eta1, a->eta2, a->max_vel, a->variant,
                                         a->neighb_type, a->neighb_param, a->seed, first_generation, stream);
        case PGC_ALGO_NSGA2:
            return pgc_nsga2_evolve_device(prob, d_x, d_f, n, a->gens, a->cr, a->eta_c, a->m, a->eta_m, a->seed, first_generation, stream);
        case PGC_ALGO_SGA:
            return pgc_sga_evolve_device(prob, d_x, d_f, n, a->gens, a->cr, a->eta_c, a->m, a->param_m, a->param_s, a->crossover, a->mutation,
                                         a->selection, a->seed, first_generation, stream);
        case PGC_ALGO_CMAES:
            return pgc_cmaes_evolve_device(prob, d_x, d_f, n, a->gens, a->cma_cc, a->cma_cs, a->cma_c1, a->cma_cmu, a->sigma0, a->ftol, a->xtol,
                                           static_cast<int>(a->force_bounds), a->seed, first_generation, gens_done, nullptr, stream);
        case PGC_ALGO_XNES:
            return pgc_xnes_evolve_device(prob, d_x, d_f, n, a->gens, a->cma_cc, a->cma_cs, a->cma_c1, a->sigma0, a->ftol, a->xtol,
                                          static_cast<int>(a->force_bounds), a->seed, first_generation, gens_done, nullptr, stream);
        case PGC_ALGO_NSPSO:
            return pgc_nspso_evolve_device(prob, d_x, d_f, n, a->gens, a->omega, a->nspso_c1, a->nspso_c2, a->nspso_chi, a->nspso_v_coeff,
                                           a->leader_selection_range, a->diversity, a->seed, first_generation, nullptr, nullptr, nullptr, stream);
        default: set_error("pgc_algo_evolve_device: unknown algorithm %d", a->algo); return PGC_ERR_INVALID_ARGUMENT;
    }
}

int pgc_es_state_len(int algo, size_t nx, size_t *len)
{
    PGC_REQUIRE(len && (algo == PGC_ALGO_CMAES || algo == PGC_ALGO_XNES), "pgc_es_state_len: cmaes or xnes, and somewhere to put the answer");
    *len = es_state_doubles(algo, nx);
    return PGC_OK;
}

int pgc_algo_evolve_memory_device(pgc_problem *prob, const pgc_algo_desc *a, double *d_x, double *d_f, size_t n, uint32_t first_generation,
                                  unsigned *gens_done, pgc_algo_memory *mem, void *stream)
{
    PGC_REQUIRE(prob && a && d_x && d_f && mem, "pgc_algo_evolve_memory_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_algo_evolve_memory_device");
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream;
    const unsigned NP = static_cast<unsigned>(n);
    if (gens_done) *gens_done = a->gens;
    int rc;
    switch (a->algo) {
        case PGC_ALGO_SADE:
        case PGC_ALGO_DE1220: {
            const unsigned algo = static_cast<unsigned>(a->algo - PGC_ALGO_DE);
            PGC_REQUIRE(mem->a && mem->b && (algo == 1u || mem->u), "pgc_algo_evolve_memory_device: sade / de1220 keep F, CR (and the variant)");
            if (!mem->initialized) {
                if ((rc = de_init_adaptation_device(NP, algo, a->variant_adptv, a->allowed_variants, a->n_allowed, a->seed, first_generation, mem->a,
                                                    mem->b, mem->u, st)))
                    return rc;
                mem->initialized = 1;
            }
            return pgc_de_evolve_device(prob, d_x, d_f, n, a->gens, algo, a->variant, a->variant_adptv, a->F, a->CR, a->allowed_variants,
                                        a->n_allowed, a->ftol, a->xtol, mem->a, mem->b, algo == 2u ? mem->u : nullptr, a->seed, first_generation,
                                        gens_done, stream);
        }
        case PGC_ALGO_PSO_GEN:
            PGC_REQUIRE(mem->a, "pgc_algo_evolve_memory_device: pso_gen keeps the velocities");
            PGC_NO_INTEGER_GENES(prob, "pgc_algo_evolve_memory_device");
            if (!mem->initialized) {
                if ((rc = pso_init_velocity_device(prob, NP, a->max_vel, a->seed, first_generation, mem->a, st))) return rc;
                mem->initialized = 1;
            }
            return pgc_pso_evolve_device(prob, d_x, d_f, mem->a, nullptr, n, a->gens, a->omega, a->eta1, a->eta2, a->max_vel, a->variant,
                                         a->neighb_type, a->neighb_param, a->seed, first_generation, stream);
        case PGC_ALGO_NSPSO:
            PGC_REQUIRE(mem->a && mem->b && mem->c, "pgc_algo_evolve_memory_device: nspso keeps the velocities and the archive");
            PGC_NO_INTEGER_GENES(prob, "pgc_algo_evolve_memory_device");
            if (!mem->initialized) {
                if ((rc = nspso_init_memory_device(prob, d_x, d_f, NP, a->nspso_v_coeff, a->seed, first_generation, mem->a, mem->b, mem->c, st)))
                    return rc;
                mem->initialized = 1;
            }
            return pgc_nspso_evolve_device(prob, d_x, d_f, n, a->gens, a->omega, a->nspso_c1, a->nspso_c2, a->nspso_chi, a->nspso_v_coeff,
                                           a->leader_selection_range, a->diversity, a->seed, first_generation, mem->a, mem->b, mem->c, stream);
        case PGC_ALGO_CMAES: // the state says itself whether it is usable (dimension, population size): `initialized` is not consulted
            PGC_REQUIRE(mem->h_state, "pgc_algo_evolve_memory_device: cmaes keeps its state in h_state (pgc_es_state_len doubles)");
            PGC_NO_INTEGER_GENES(prob, "pgc_algo_evolve_memory_device");
            if (!mem->initialized) mem->h_state[0] = 0.;
            mem->initialized = 1;
            return cmaes_evolve_device(prob, d_x, d_f, n, a->gens, a->cma_cc, a->cma_cs, a->cma_c1, a->cma_cmu, a->sigma0, a->ftol, a->xtol,
                                       static_cast<int>(a->force_bounds), a->seed, first_generation, gens_done, nullptr, problem_eval_device, st,
                                       mem->h_state, mem->h_state_len);
        case PGC_ALGO_XNES:
            PGC_REQUIRE(mem->h_state, "pgc_algo_evolve_memory_device: xnes keeps its state in h_state (pgc_es_state_len doubles)");
            PGC_NO_INTEGER_GENES(prob, "pgc_algo_evolve_memory_device");
            if (!mem->initialized) mem->h_state[0] = 0.;
            mem->initialized = 1;
            return xnes_evolve_device(prob, d_x, d_f, n, a->gens, a->cma_cc, a->cma_cs, a->cma_c1, a->sigma0, a->ftol, a->xtol,
                                      static_cast<int>(a->force_bounds), a->seed, first_generation, gens_done, nullptr, problem_eval_device, st,
                                      mem->h_state, mem->h_state_len);
        default:
            set_error("pgc_algo_evolve_memory_device: algorithm %d keeps no state between evolve() calls", a->algo);
            return PGC_ERR_INVALID_ARGUMENT;
    }
}

int pgc_algo_log_row_len(const pgc_problem *prob, int algo, size_t *row_len)
{
    PGC_REQUIRE(prob && row_len, "pgc_algo_log_row_len: null argument");
    switch (algo) {
        case PGC_ALGO_DE: *row_len = 5; return PGC_OK;
        case PGC_ALGO_SADE: *row_len = 7; return PGC_OK;
        case PGC_ALGO_DE1220: *row_len = 8; return PGC_OK;
        case PGC_ALGO_PSO_GEN: *row_len = 6; return PGC_OK;
        case PGC_ALGO_SGA: *row_len = 4; return PGC_OK;
        case PGC_ALGO_CMAES:
        case PGC_ALGO_XNES: *row_len = 6; return PGC_OK;
        case PGC_ALGO_NSGA2:
        case PGC_ALGO_NSPSO: *row_len = 2 + prob->nobj; return PGC_OK;
        default:
            set_error("pgc_algo_log_row_len: algorithm %d records no log on the device", algo);
            return PGC_ERR_UNSUPPORTED;
    }
}

int pgc_algo_evolve_logged_device(pgc_problem *prob, const pgc_algo_desc *a, double *d_x, double *d_f, size_t n, uint32_t first_generation,
                                  unsigned *gens_done, pgc_algo_memory *memory, unsigned verbosity, double *log_rows, size_t max_rows,
                                  size_t *n_rows, void *stream)
{
    PGC_REQUIRE(prob && a, "pgc_algo_evolve_logged_device: null argument");
    PGC_NO_CONSTRAINTS(prob, "pgc_algo_evolve_logged_device");
    if (n_rows) *n_rows = 0;
    const auto run = [&]() {
        return memory ? pgc_algo_evolve_memory_device(prob, a, d_x, d_f, n, first_generation, gens_done, memory, stream)
                      : pgc_algo_evolve_device(prob, a, d_x, d_f, n, first_generation, gens_done, stream);
    };
    if (verbosity == 0u) return run();
    PGC_REQUIRE(log_rows && n_rows, "pgc_algo_evolve_logged_device: verbosity > 0 needs somewhere to put the log");
    size_t row_len = 0;
    if (int rc = pgc_algo_log_row_len(prob, a->algo, &row_len)) return rc;
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream;
    const size_t due = a->gens ? (a->gens - 1u) / verbosity + 1u : 0u; // generations 1, 1 + v, ... <= gens
    const size_t rows = std::min(due, max_rows);
    struct Dev {
        void *p = nullptr;
        ~Dev() { if (p) cudaFree(p); }
    } d_rows, d_count;
    PGC_CUDA(cudaMalloc(&d_rows.p, sizeof(double) * std::max<size_t>(rows * row_len, 1)));
    PGC_CUDA(cudaMalloc(&d_count.p, sizeof(unsigned)));
    PGC_CUDA(cudaMemsetAsync(d_count.p, 0, sizeof(unsigned), st));
    LogSink sink;
    sink.d_rows = static_cast<double *>(d_rows.p);
    sink.d_count = static_cast<unsigned *>(d_count.p);
    sink.verbosity = verbosity;
    sink.max_rows = static_cast<unsigned>(rows);
    sink.row_len = static_cast<unsigned>(row_len);
    tls_log = &sink;
    const int rc = run();
    tls_log = nullptr;
    if (rc != PGC_OK) return rc;
    unsigned written = 0;
    PGC_CUDA(cudaMemcpyAsync(&written, d_count.p, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    if (written) PGC_CUDA(cudaMemcpy(log_rows, d_rows.p, sizeof(double) * written * row_len, cudaMemcpyDeviceToHost));
    if (!written && !sink.host_rows.empty()) { // a loop that runs on the host (cmaes) kept its lines there
        written = static_cast<unsigned>(std::min(sink.host_rows.size() / row_len, rows));
        std::copy(sink.host_rows.begin(), sink.host_rows.begin() + static_cast<std::ptrdiff_t>(written * row_len), log_rows);
    }
    *n_rows = written;
    return PGC_OK;
}

namespace
{
struct Capture { // one per thread: the sink of pgc_log_capture_begin / _end
    LogSink sink;
    bool active = false;
};
thread_local Capture tls_capture;
} // namespace

int pgc_log_capture_begin(pgc_ctx *ctx, unsigned verbosity, size_t max_rows, size_t row_len)
{
    PGC_REQUIRE(ctx && verbosity >= 1u && row_len >= 1u, "pgc_log_capture_begin: a context, a verbosity >= 1 and a row length are needed");
    PGC_REQUIRE(!tls_capture.active && !tls_log, "pgc_log_capture_begin: a capture is already running on this thread");
    PGC_CUDA(cudaSetDevice(ctx->device));
    Capture &c = tls_capture;
    c.sink = LogSink{};
    PGC_CUDA(cudaMalloc(&c.sink.d_rows, sizeof(double) * std::max<size_t>(max_rows * row_len, 1)));
    if (cudaMalloc(&c.sink.d_count, sizeof(unsigned)) != cudaSuccess) {
        cudaFree(c.sink.d_rows);
        set_error("pgc_log_capture_begin: out of device memory");
        return PGC_ERR_OUT_OF_MEMORY;
    }
    PGC_CUDA(cudaMemsetAsync(c.sink.d_count, 0, sizeof(unsigned), ctx->stream));
    c.sink.verbosity = verbosity, c.sink.max_rows = static_cast<unsigned>(max_rows), c.sink.row_len = static_cast<unsigned>(row_len);
    c.active = true;
    tls_log = &c.sink;
    return PGC_OK;
}

int pgc_log_capture_end(pgc_ctx *ctx, double *rows_out, size_t *n_rows)
{
    PGC_REQUIRE(ctx && n_rows, "pgc_log_capture_end: null argument");
    PGC_REQUIRE(tls_capture.active, "pgc_log_capture_end: no capture is running on this thread");
    Capture &c = tls_capture;
    tls_log = nullptr;
    c.active = false;
    struct Release {
        LogSink &s;
        ~Release()
        {
            cudaFree(s.d_rows);
            cudaFree(s.d_count);
            s = LogSink{};
        }
    } release{c.sink};
    PGC_CUDA(cudaSetDevice(ctx->device));
    unsigned written = 0;
    PGC_CUDA(cudaMemcpyAsync(&written, c.sink.d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    PGC_CUDA(cudaStreamSynchronize(ctx->stream));
    const size_t rl = c.sink.row_len;
    if (written && rows_out) PGC_CUDA(cudaMemcpy(rows_out, c.sink.d_rows, sizeof(double) * written * rl, cudaMemcpyDeviceToHost));
    if (!written && !c.sink.host_rows.empty()) {
        written = static_cast<unsigned>(std::min<size_t>(c.sink.host_rows.size() / rl, c.sink.max_rows));
        if (rows_out) std::copy(c.sink.host_rows.begin(), c.sink.host_rows.begin() + static_cast<std::ptrdiff_t>(written * rl), rows_out);
    }
    *n_rows = written;
    return PGC_OK;
}

int pgc_population_init_device(pgc_problem *prob, size_t n, uint64_t seed, double *d_x, double *d_f, uint64_t *d_ids, void *stream)
{
    PGC_REQUIRE(prob && (d_x || n == 0), "pgc_population_init_device: null argument");
    PGC_CUDA(cudaSetDevice(prob->ctx->device));
    return population_init_device(prob, n, seed, d_x, d_f, reinterpret_cast<unsigned long long *>(d_ids), problem_eval_device,
                                  stream ? static_cast<cudaStream_t>(stream) : prob->ctx->stream);
}

int pgc_select_best_device(pgc_ctx *ctx, const uint64_t *d_ids, const double *d_x, const double *d_f, size_t n, size_t nx, size_t nobj,
                           int rate_is_frac, double rate, uint64_t *d_ids_out, double *d_x_out, double *d_f_out, size_t *n_out, void *stream)
{
    PGC_REQUIRE(ctx && n_out && (n == 0 || (d_ids && d_x && d_f && d_ids_out && d_x_out && d_f_out)), "pgc_select_best_device: null argument");
    PGC_REQUIRE(nobj >= 1 && nx >= 1 && n < 0x7fffffffull, "pgc_select_best_device: invalid sizes");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return select_best_policy_device(ctx, reinterpret_cast<const unsigned long long *>(d_ids), d_x, d_f, n, nx, nobj, rate_is_frac, rate,
                                     reinterpret_cast<unsigned long long *>(d_ids_out), d_x_out, d_f_out, n_out,
                                     stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int pgc_fair_replace_device(pgc_ctx *ctx, uint64_t *d_ids, double *d_x, double *d_f, size_t n, size_t nx, size_t nobj, int rate_is_frac,
                            double rate, const uint64_t *d_mids, const double *d_mx, const double *d_mf, size_t nm, void *stream)
{
    PGC_REQUIRE(ctx && (n == 0 || (d_ids && d_x && d_f)) && (nm == 0 || (d_mids && d_mx && d_mf)), "pgc_fair_replace_device: null argument");
    PGC_REQUIRE(nobj >= 1 && nx >= 1 && n + nm < 0x7fffffffull, "pgc_fair_replace_device: invalid sizes");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return fair_replace_policy_device(ctx, reinterpret_cast<unsigned long long *>(d_ids), d_x, d_f, n, nx, nobj, rate_is_frac, rate,
                                      reinterpret_cast<const unsigned long long *>(d_mids), d_mx, d_mf, nm,
                                      stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int pgc_sort_population_con_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t nec, size_t nic, const double *tol, uint32_t *d_order,
                                   void *stream)
{
    PGC_REQUIRE(ctx && (n == 0 || (d_f && d_order)), "pgc_sort_population_con_device: null argument");
    PGC_REQUIRE(n < 0x7fffffffull, "pgc_sort_population_con_device: too many individuals");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return sort_population_con_device(ctx, d_f, n, nec, nic, tol, d_order, stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int pgc_select_best_con_device(pgc_ctx *ctx, const uint64_t *d_ids, const double *d_x, const double *d_f, size_t n, size_t nx, size_t nec,
                               size_t nic, const double *tol, int rate_is_frac, double rate, uint64_t *d_ids_out, double *d_x_out,
                               double *d_f_out, size_t *n_out, void *stream)
{
    PGC_REQUIRE(ctx && n_out && (n == 0 || (d_ids && d_x && d_f && d_ids_out && d_x_out && d_f_out)), "pgc_select_best_con_device: null argument");
    PGC_REQUIRE(nx >= 1 && n < 0x7fffffffull, "pgc_select_best_con_device: invalid sizes");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return select_best_con_policy_device(ctx, reinterpret_cast<const unsigned long long *>(d_ids), d_x, d_f, n, nx, nec, nic, tol, rate_is_frac, rate,
                                         reinterpret_cast<unsigned long long *>(d_ids_out), d_x_out, d_f_out, n_out,
                                         stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int pgc_fair_replace_con_device(pgc_ctx *ctx, uint64_t *d_ids, double *d_x, double *d_f, size_t n, size_t nx, size_t nec, size_t nic,
                                const double *tol, int rate_is_frac, double rate, const uint64_t *d_mids, const double *d_mx, const double *d_mf,
                                size_t nm, void *stream)
{
    PGC_REQUIRE(ctx && (n == 0 || (d_ids && d_x && d_f)) && (nm == 0 || (d_mids && d_mx && d_mf)), "pgc_fair_replace_con_device: null argument");
    PGC_REQUIRE(nx >= 1 && n + nm < 0x7fffffffull, "pgc_fair_replace_con_device: invalid sizes");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return fair_replace_con_policy_device(ctx, reinterpret_cast<unsigned long long *>(d_ids), d_x, d_f, n, nx, nec, nic, tol, rate_is_frac, rate,
                                          reinterpret_cast<const unsigned long long *>(d_mids), d_mx, d_mf, nm,
                                          stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int pgc_topology_connections(int kind, size_t n, size_t i, double weight, size_t *idx_out, double *w_out, size_t *count)
{
    PGC_REQUIRE(idx_out && w_out && count, "pgc_topology_connections: null argument");
    PGC_REQUIRE(weight >= 0.0 && weight <= 1.0, "invalid weight for the edge of a topology: the value %g is not in the [0., 1.] range", weight);
    std::vector<size_t> src;
    switch (kind) {
        case 0: PGC_REQUIRE(i < n, "pgc_topology_connections: vertex %zu of %zu", i, n); break; // unconnected.cpp: no edges
        case 1: {
            int rc = ring_connections(n, i, src);
            if (rc != PGC_OK) return rc;
            break;
        }
        case 2: // fully_connected.cpp:86-115
            PGC_REQUIRE(i < n,
                        "Cannot get the connections to the vertex at index %zu in a fully connected topology: the number of vertices in the "
                        "topology is only %zu",
                        i, n);
            for (size_t j = 0; j < n; ++j)
                if (j != i) src.push_back(j);
            break;
        default: set_error("pgc_topology_connections: unknown topology kind %d", kind); return PGC_ERR_INVALID_ARGUMENT;
    }
    *count = src.size();
    for (size_t q = 0; q < src.size(); ++q) {
        idx_out[q] = src[q];
        w_out[q] = weight;
    }
    return PGC_OK;
}

int pgc_malloc_device(pgc_ctx *ctx, size_t bytes, void **out)
{
    PGC_REQUIRE(ctx && out, "pgc_malloc_device: null argument");
    PGC_CUDA(cudaSetDevice(ctx->device));
    PGC_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return PGC_OK;
}

int pgc_free_device(pgc_ctx *ctx, void *ptr)
{
    PGC_REQUIRE(ctx, "pgc_free_device: null context");
    PGC_CUDA(cudaSetDevice(ctx->device));
    PGC_CUDA(cudaFree(ptr));
    return PGC_OK;
}

int pgc_malloc_pinned(pgc_ctx *ctx, size_t bytes, void **out)
{
    PGC_REQUIRE(ctx && out, "pgc_malloc_pinned: null argument");
    PGC_CUDA(cudaSetDevice(ctx->device));
    PGC_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return PGC_OK;
}

int pgc_free_pinned(pgc_ctx *ctx, void *ptr)
{
    PGC_REQUIRE(ctx, "pgc_free_pinned: null context");
    PGC_CUDA(cudaFreeHost(ptr));
    return PGC_OK;
}

int pgc_memcpy_h2d(pgc_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    PGC_REQUIRE(ctx, "pgc_memcpy_h2d: null context");
    PGC_CUDA(cudaSetDevice(ctx->device));
    PGC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    PGC_CUDA(cudaStreamSynchronize(ctx->stream));
    return PGC_OK;
}

int pgc_memcpy_d2h(pgc_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    PGC_REQUIRE(ctx, "pgc_memcpy_d2h: null context");
    PGC_CUDA(cudaSetDevice(ctx->device));
    PGC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PGC_CUDA(cudaStreamSynchronize(ctx->stream));
    return PGC_OK;
}

int pgc_measure_fp64_peak(pgc_ctx *ctx, int iters, double *tflops)
{
    PGC_REQUIRE(ctx && tflops && iters > 0, "pgc_measure_fp64_peak: bad argument");
    return fp64_peak(ctx, iters, tflops);
}

int pgc_debug_fp64_mix_probe(pgc_ctx *ctx, int iters, int total_warps, int dmma_warps, double *tflops2)
{
    PGC_REQUIRE(ctx && tflops2 && iters > 0 && total_warps >= 1 && total_warps <= 16 && dmma_warps >= 0
                    && dmma_warps <= total_warps,
                "pgc_debug_fp64_mix_probe: bad argument");
    return fp64_mix_probe(ctx, iters, total_warps, dmma_warps, tflops2);
}

int pgc_measure_fp64_mma_peak(pgc_ctx *ctx, int iters, double *tflops)
{
    PGC_REQUIRE(ctx && tflops && iters > 0, "pgc_measure_fp64_mma_peak: bad argument");
    return fp64_mma_peak(ctx, iters, tflops);
}

} // extern "C"
