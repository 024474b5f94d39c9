// pgc_internal.cuh - shared internals of libpgc.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pagmo_cuda/pgc.h"

namespace pgc
{

// ---- error plumbing -------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define PGC_CUDA(call)                                                                                                 \
    do {                                                                                                               \
        cudaError_t e__ = (call);                                                                                      \
        if (e__ != cudaSuccess) return ::pgc::cuda_fail(e__, #call, __FILE__, __LINE__);                               \
    } while (0)

#define PGC_REQUIRE(cond, ...)                                                                                         \
    do {                                                                                                               \
        if (!(cond)) {                                                                                                 \
            ::pgc::set_error(__VA_ARGS__);                                                                             \
            return PGC_ERR_INVALID_ARGUMENT;                                                                           \
        }                                                                                                              \
    } while (0)

// ---- CEC2014 recipe (host built, see cec2014_recipe.cpp) ------------------------------------------------
// Primitive ids: the 14 basic functions used by f1..f30 (reference cec2014.cpp:375-783).
enum Prim : int {
    P_ELLIPS = 0,
    P_BENT_CIGAR,
    P_DISCUS,
    P_ROSENBROCK,
    P_ACKLEY,
    P_WEIERSTRASS,
    P_GRIEWANK,
    P_RASTRIGIN,
    P_SCHWEFEL,
    P_KATSUURA,
    P_HAPPYCAT,
    P_HGBAT,
    P_GRIE_ROSEN,
    P_ESCAFFER6,
    P_COUNT
};

constexpr int kMaxGroups = 5;
constexpr int kMaxStages = 5;

// One group of a stage: a primitive applied (with its own sh_rate, s_flag=r_flag=0 semantics) to the
// coordinate range [off, off+len) of the stage vector.
struct GroupDesc {
    int prim;
    int off;
    int len;
    int tab_off;   // offset (doubles) into the problem's constant table: per-coordinate coefficients
    double rate;   // post-rotation scale (1.0 for the basic functions: exact no-op)
    double c0, c1; // primitive-specific host-precomputed constants
};

// One stage = one shift(+scale)(+rotate)(+permute) followed by 1..5 groups (cec2014.cpp sr_func + hfXX/cfXX).
struct StageDesc {
    int comp;       // component index i: Os offset i*D, Mr offset i*D*D, S offset i*D
    int rotate;     // r_flag
    int permute;    // hybrid: y[j] = z[S[j]-1]
    int ngroups;
    double pre_rate; // sh_rate applied before the rotation
    int scaled;      // composition: fit = mul*fit/div
    double mul, div;
    GroupDesc g[kMaxGroups];
};

struct Cec2014Recipe {
    int func;
    int dim;
    int nstages;
    int composition; // cf_cal combine
    double delta[kMaxStages];
    double cbias[kMaxStages];
    double fbias; // 100*func
    StageDesc st[kMaxStages];
    std::vector<double> table; // coefficient tables referenced by GroupDesc::tab_off
    double flops_per_eval;
    double transc_per_eval;
};

int build_cec2014_recipe(unsigned func, unsigned dim, Cec2014Recipe &out);

struct Cec2013Plan; // eval_cec2013.cu

} // namespace pgc

// ---- host-side helper: a few threads that copy one buffer together ------------------------------------------
// pgc_eval_host stages PAGEABLE caller memory (a std::vector<double> coming through pagmo::bfe) in pinned chunks; one thread's
// memcpy runs at ~10 GB/s, a fifth of what the PCIe link behind it takes, so the chunk is copied by several threads at once.
namespace pgc
{
class CopyPool
{
public:
    explicit CopyPool(unsigned nthreads)
    {
        for (unsigned t = 0; t < nthreads; ++t) m_threads.emplace_back([this, t, nthreads] { worker(t, nthreads); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_mtx);
            m_stop = true;
            ++m_epoch;
        }
        m_cv.notify_all();
        for (auto &t : m_threads) t.join();
    }
    // dst[0, bytes) = src[0, bytes), split into one contiguous piece per helper plus one for the caller
    void copy(void *dst, const void *src, size_t bytes)
    {
        const size_t parts = m_threads.size() + 1;
        if (bytes < (4u << 20) || m_threads.empty()) {
            std::memcpy(dst, src, bytes);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m_mtx);
            m_dst = static_cast<char *>(dst);
            m_src = static_cast<const char *>(src);
            m_bytes = bytes;
            m_pending = static_cast<unsigned>(m_threads.size());
            ++m_epoch;
        }
        m_cv.notify_all();
        piece(parts - 1, parts);
        std::unique_lock<std::mutex> lk(m_mtx);
        m_done.wait(lk, [this] { return m_pending == 0; });
    }

private:
    void piece(size_t i, size_t parts) const
    {
        const size_t unit = ((m_bytes + parts - 1) / parts + 4095) & ~static_cast<size_t>(4095);
        const size_t lo = std::min(m_bytes, i * unit), hi = std::min(m_bytes, lo + unit);
        if (hi > lo) std::memcpy(m_dst + lo, m_src + lo, hi - lo);
    }
    void worker(unsigned t, unsigned nthreads)
    {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_mtx);
                m_cv.wait(lk, [&] { return m_epoch != seen; });
                seen = m_epoch;
                if (m_stop) return;
            }
            piece(t, nthreads + 1u);
            {
                std::lock_guard<std::mutex> lk(m_mtx);
                --m_pending;
            }
            m_done.notify_one();
        }
    }
    std::vector<std::thread> m_threads;
    std::mutex m_mtx;
    std::condition_variable m_cv, m_done;
    char *m_dst = nullptr;
    const char *m_src = nullptr;
    size_t m_bytes = 0;
    unsigned m_pending = 0;
    unsigned long long m_epoch = 0;
    bool m_stop = false;
};
} // namespace pgc

// ---- opaque handle layouts -----------------------------------------------------------------------------
struct pgc_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;     // compute
    cudaStream_t copy_stream[2] = {nullptr, nullptr};
    std::atomic<uint64_t> launches{0};
    // pinned + device staging ring for pgc_eval_host
    static constexpr int kRing = 3;
    void *h_in[kRing] = {};
    void *h_out[kRing] = {};
    void *d_in[kRing] = {};
    void *d_out[kRing] = {};
    size_t ring_in_bytes = 0, ring_out_bytes = 0;
    cudaEvent_t ev_in[kRing] = {}, ev_k[kRing] = {}, ev_out[kRing] = {};
    // scratch (composition stage outputs etc.)
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    pgc::CopyPool *copy_pool = nullptr; // created on the first pageable pgc_eval_host call
    int sharers = 1; // contexts expected to run concurrently on this device (pgc_ctx_set_sharers): launch-shape heuristics divide the SMs by it
};

namespace pgc
{
// A cached workspace of a generation loop bound to one population (de.cu: scratch buffers + the instantiated CUDA graph of a
// batch of generations).  Owned by the problem handle; freed with it.
struct LoopWorkspace {
    virtual ~LoopWorkspace() = default;
    unsigned long long last_use = 0;
    int users = 0; // calls currently running on this workspace (guarded by pgc_problem::work_mu): never evicted while > 0
};
} // namespace pgc

struct pgc_problem {
    pgc_ctx *ctx = nullptr;
    pgc_problem_desc desc{}; // table pointers nulled after the copy
    size_t nx = 0, nobj = 1;
    size_t nix = 0; // integer dimension (the last nix genes, problem::get_nix): zdt5 only; a meta-problem inherits its inner one's
    // constraints (problem::get_nec / get_nic / get_c_tol): hock_schittkowski_71 and luksan_vlcek1 only; fitness rows are
    // [nobj | nec | nic] wide.  Generation operators refuse nec + nic > 0, as the reference's algorithms do; unconstrain removes them.
    size_t nec = 0, nic = 0;
    std::vector<double> c_tol;
    size_t nf() const { return nobj + nec + nic; }
    std::vector<double> lb, ub;
    std::string name;
    // device tables
    double *d_rotation = nullptr; // cec: padded/retiled per component (see eval_cec2014.cu)
    double *d_shift = nullptr;
    int *d_shuffle = nullptr;
    double *d_table = nullptr;
    pgc::Cec2014Recipe cec14;
    pgc::Cec2013Plan *cec13 = nullptr;
    double flops_per_eval = 0, transc_per_eval = 0;
    int strict = 0; // pgc_problem_set_strict: reference summation order in the rotations (cec2013)
    unsigned config_epoch = 0; // bumped by every setter that changes what the evaluation kernels are launched with
    std::mutex work_mu;
    std::vector<std::unique_ptr<pgc::LoopWorkspace>> work; // generation-loop workspaces, least recently used evicted (de.cu)
    unsigned long long work_clock = 0;
    // meta-problems (meta.cu): the wrapped problem (borrowed), translation | weight+z on the device, decomposition method
    pgc_problem *inner = nullptr;
    double *d_meta = nullptr;
    int meta_method = 0;
};

namespace pgc
{
// Stream-ordered scratch from the device memory pool, released (in stream order) when it goes out of scope - whatever path leaves
// the function, so an early return after a failed allocation or copy does not leak the buffers taken before it.
struct StreamScratch {
    cudaStream_t st;
    std::vector<void *> owned;
    explicit StreamScratch(cudaStream_t s) : st(s) {}
    StreamScratch(const StreamScratch &) = delete;
    StreamScratch &operator=(const StreamScratch &) = delete;
    ~StreamScratch()
    {
        for (void *p : owned) cudaFreeAsync(p, st);
    }
    template <class T> cudaError_t get(T **out, size_t bytes)
    {
        void *p = nullptr;
        const cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 1, st);
        if (e == cudaSuccess) owned.push_back(p);
        *out = static_cast<T *>(p);
        return e;
    }
};
} // namespace pgc

namespace pgc
{
int ensure_scratch(pgc_ctx *ctx, size_t bytes);
int ctx_sync(pgc_ctx *ctx); // the context's own streams (never the whole device: see capi.cu)
// family back-ends: validate + upload tables (create) and launch (eval, asynchronous on `stream`)
int simple_create(pgc_problem *p);
int simple_eval(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream);
int lj_create(pgc_problem *p);
int lj_eval(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream);
int mo_create(pgc_problem *p);
int mo_eval(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream);
int cec2014_create(pgc_problem *p, const pgc_problem_desc *d);
int cec2014_eval(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream);
void cec2014_destroy(pgc_problem *p);
int cec2013_create(pgc_problem *p, const pgc_problem_desc *d);
int cec2013_eval(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream);
void cec2013_destroy(pgc_problem *p);
int constrained_create(pgc_problem *p);
int constrained_eval(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream);
int feasibility_rows(pgc_problem *p, const double *d_f, size_t n, unsigned char *d_feasible, cudaStream_t stream);
int meta_create(pgc_problem *inner, int family, const double *a, const double *b, size_t len, int method, pgc_problem **out);
int meta_eval(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t s);
void meta_destroy(pgc_problem *p);
int cec2014_phase_cycles(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, unsigned long long *out);
int fnds_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t m, unsigned *d_rank, unsigned *d_dom_count, unsigned *d_order,
                unsigned *d_front_off, unsigned *nfronts_out, cudaStream_t st, unsigned stop_after = 0, unsigned *d_key_out = nullptr);
struct SelectedRanking { // fast_non_dominated_sorting of the individuals select_best_device picked, in their new numbering
    unsigned *rank, *order, *front_off; // device, N / N / N + 1 entries (caller-owned)
    unsigned nfronts;
    bool valid;
};
int crowding_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t m, const unsigned *d_order, const unsigned *d_front_off,
                    unsigned nfronts, int small_rule, double *d_cd, cudaStream_t st);
int select_best_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t m, size_t N, unsigned *d_out, unsigned *nout, cudaStream_t st,
                       SelectedRanking *ranking = nullptr);
int sort_population_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t m, unsigned *d_out, cudaStream_t st, size_t limit = 0);
int problem_eval_device(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t s);
int philox_permutation_device(pgc_ctx *ctx, unsigned n, unsigned long long seed, unsigned tag, unsigned generation, unsigned *d_perm,
                              cudaStream_t st);
int de_init_adaptation_device(unsigned NP, unsigned algo, unsigned variant_adptv, const unsigned *allowed, unsigned n_allowed,
                              unsigned long long seed, unsigned generation, double *d_F, double *d_CR, unsigned *d_variant, cudaStream_t st);
int pso_init_velocity_device(pgc_problem *prob, unsigned n, double max_vel, unsigned long long seed, unsigned generation, double *d_v,
                             cudaStream_t st);
// ---- UDA logs (algo_log.cu): where the generation loops append the reference's log lines while a logged evolve() runs ------------
struct LogSink {
    double *d_rows = nullptr;   // [max_rows x row_len]
    unsigned *d_count = nullptr; // rows written so far (advanced on the device, in stream order)
    unsigned verbosity = 0, max_rows = 0, row_len = 0;
    std::vector<double> host_rows; // loops that run on the host (cmaes) append their lines here instead
};
extern thread_local LogSink *tls_log;
// generation `gen` (1-based within this evolve() call) is one the reference logs (de.cpp:327)
inline bool log_due(unsigned gen)
{
    const LogSink *L = tls_log;
    return L && L->verbosity && (gen % L->verbosity == 1u || L->verbosity == 1u);
}
int log_ideal_device(pgc_ctx *ctx, const double *d_f, unsigned n, unsigned m, unsigned gen, unsigned long long fevals, cudaStream_t st);
int log_sga_device(pgc_ctx *ctx, const double *d_f_parents, const double *d_f_children, unsigned n, unsigned gen, unsigned long long fevals,
                   cudaStream_t st);
int log_pso_device(pgc_ctx *ctx, const double *d_X, const double *d_V, const double *d_lbfit, const double *d_lb, const double *d_ub, unsigned n,
                   unsigned dim, unsigned gen, unsigned long long fevals, cudaStream_t st);
int nspso_init_memory_device(pgc_problem *prob, const double *d_x, const double *d_f, unsigned NP, double v_coeff, unsigned long long seed,
                             unsigned generation, double *d_vel, double *d_best_x, double *d_best_f, cudaStream_t st);
int moead_gen_evolve_device(pgc_problem *prob, double *d_x, double *d_f, unsigned NP, unsigned gens, const double *h_weights,
                            const unsigned *h_neigh, unsigned T, int decomposition, double CR, double F, double eta_m, double realb, unsigned limit,
                            int preserve_diversity, unsigned long long seed, unsigned first_generation,
                            int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st);
int gaco_evolve_device(pgc_problem *prob, double *d_x, double *d_f, unsigned n, unsigned gens, unsigned ker, double acc, unsigned threshold,
                       unsigned n_gen_mark, unsigned impstop, unsigned evalstop, double focus, unsigned long long seed, unsigned first_generation,
                       pgc_gaco_state *state, unsigned *gens_done,
                       int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st);
int maco_evolve_device(pgc_problem *prob, double *d_x, double *d_f, unsigned n, unsigned gens, unsigned ker, unsigned threshold,
                       unsigned n_gen_mark, unsigned evalstop, double focus, unsigned long long seed, unsigned first_generation,
                       pgc_maco_state *state, unsigned *gens_done,
                       int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st);
int xnes_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t lam, unsigned gens, double eta_mu, double eta_sigma, double eta_b,
                       double sigma0, double ftol, double xtol, int force_bounds, unsigned long long seed, unsigned first_generation,
                       unsigned *gens_done, double *sigma_out, int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t),
                       cudaStream_t st, double *es_state = nullptr, size_t es_state_len = 0);
int nspso_evolve_device(pgc_problem *prob, double *d_x, double *d_f, unsigned NP, unsigned gens, double omega, double c1, double c2, double chi,
                        double v_coeff, unsigned leader_selection_range, unsigned diversity, unsigned long long seed, unsigned first_generation,
                        double *d_vel, double *d_best_x, double *d_best_f,
                        int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st);
int nsga2_variation_device(pgc_ctx *ctx, const double *d_x, const unsigned *d_rank, const double *d_cd, unsigned NP, unsigned nx,
                           const double *d_lb, const double *d_ub, const unsigned *d_sh1, const unsigned *d_sh2, double cr,
                           double eta_c, double m, double eta_m, unsigned long long seed, unsigned generation, double *d_children,
                           cudaStream_t st, unsigned nix = 0);
int nsga2_evolve_device(pgc_problem *prob, double *d_x, double *d_f, unsigned NP, unsigned gens, double cr, double eta_c, double m,
                        double eta_m, unsigned long long seed, unsigned first_generation,
                        int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st);
int pso_evolve_device(pgc_problem *prob, double *d_x, double *d_f, double *d_v, double *d_xcur, unsigned n, unsigned gens, double omega,
                      double eta1, double eta2, double max_vel, unsigned variant, unsigned neighb_type, unsigned neighb_param,
                      unsigned long long seed, unsigned first_generation,
                      int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st);
int pso_shard_step_device(pgc_problem *prob, double *d_X, double *d_V, double *d_lbX_ext, double *d_lbfit_ext, unsigned n_loc, unsigned radius,
                          unsigned index_offset, double omega, double eta1, double eta2, double max_vel, unsigned variant,
                          unsigned long long seed, unsigned generation, int init_velocity,
                          int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st, double *d_cand = nullptr);
int de_evolve_device(pgc_problem *prob, double *d_x, double *d_f, unsigned NP, unsigned gens, unsigned algo, unsigned variant,
                     unsigned variant_adptv, double F, double CR, const unsigned *allowed, unsigned n_allowed, double ftol, double xtol,
                     double *d_F, double *d_CR, unsigned *d_variant, unsigned long long seed, unsigned first_generation,
                     unsigned *gens_done, int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st);
int population_init_device(pgc_problem *prob, size_t n, unsigned long long seed, double *d_x, double *d_f, unsigned long long *d_ids,
                           int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st);
int select_best_policy_device(pgc_ctx *ctx, const unsigned long long *d_ids, const double *d_x, const double *d_f, size_t n, size_t nx,
                              size_t nobj, int rate_is_frac, double rate, unsigned long long *d_ids_out, double *d_x_out, double *d_f_out,
                              size_t *n_out, cudaStream_t st);
int fair_replace_policy_device(pgc_ctx *ctx, unsigned long long *d_ids, double *d_x, double *d_f, size_t n, size_t nx, size_t nobj,
                               int rate_is_frac, double rate, const unsigned long long *d_mids, const double *d_mx, const double *d_mf,
                               size_t nm, cudaStream_t st);
int select_best_con_policy_device(pgc_ctx *ctx, const unsigned long long *d_ids, const double *d_x, const double *d_f, size_t n, size_t nx,
                                  size_t nec, size_t nic, const double *tol, int rate_is_frac, double rate, unsigned long long *d_ids_out,
                                  double *d_x_out, double *d_f_out, size_t *n_out, cudaStream_t st);
int fair_replace_con_policy_device(pgc_ctx *ctx, unsigned long long *d_ids, double *d_x, double *d_f, size_t n, size_t nx, size_t nec, size_t nic,
                                   const double *tol, int rate_is_frac, double rate, const unsigned long long *d_mids, const double *d_mx,
                                   const double *d_mf, size_t nm, cudaStream_t st);
int sort_population_con_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t nec, size_t nic, const double *tol, unsigned *d_order,
                               cudaStream_t st);
int hv_fpras_host(pgc_ctx *ctx, const double *points, size_t n, size_t m, const double *r, double eps, double delta, unsigned long long seed,
                  double *hv_out);
int hv_approx_extreme_host(pgc_ctx *ctx, const double *points, size_t n, size_t m, const double *r, int greatest, int use_exact,
                           unsigned trivial_subcase_size, double eps, double delta, double delta_multiplier, double alpha,
                           double initial_delta_coeff, double gamma, unsigned long long seed, size_t *idx_out);
int policy_rate_count(const char *who, int rate_is_frac, double rate, size_t n, size_t *out);
int so_best_indices_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t k, unsigned *d_sel, cudaStream_t st);
int sga_evolve_device(pgc_problem *prob, double *d_x, double *d_f, unsigned NP, unsigned gens, double cr, double eta_c, double m,
                      double param_m, unsigned param_s, unsigned crossover, unsigned mutation, unsigned selection, unsigned long long seed,
                      unsigned first_generation, int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st);
int ring_connections(size_t n, size_t i, std::vector<size_t> &out);
int hv_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t m, const double *r, int compute, double *d_out, cudaStream_t st);
int weighted_gram_device(pgc_ctx *ctx, const double *d_rows, const unsigned *d_idx, const double *d_center, const double *d_w, size_t k,
                         size_t D, double scale_div, double *d_out, cudaStream_t st);
int weighted_mean_device(pgc_ctx *ctx, const double *d_rows, const unsigned *d_idx, const double *d_w, size_t k, size_t D, double *d_out,
                         cudaStream_t st);
int cmaes_sample_device(pgc_ctx *ctx, const double *d_mean, const double *d_bd, double sigma, size_t lambda, size_t D, unsigned long long seed,
                        unsigned generation, double *d_z, double *d_x, cudaStream_t st);
int cmaes_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t lam, unsigned gens, double cc, double cs, double c1, double cmu,
                        double sigma0, double ftol, double xtol, int force_bounds, unsigned long long seed, unsigned first_generation,
                        unsigned *gens_done, double *sigma_out, int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t),
                        cudaStream_t st, double *es_state = nullptr, size_t es_state_len = 0);
size_t es_state_doubles(int algo, size_t D);
int fp64_peak(pgc_ctx *ctx, int iters, double *tflops);
int fp64_mma_peak(pgc_ctx *ctx, int iters, double *tflops);
int fp64_mix_probe(pgc_ctx *ctx, int iters, int total_warps, int dmma_warps, double *tflops_out);
} // namespace pgc
