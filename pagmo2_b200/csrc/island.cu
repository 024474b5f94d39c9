// island.cu - device-resident islands and their migration over NCCL (sm_100a, NVLink 5 / NVSwitch).
//
// Replaces, for GPU islands, the data path of reference island::evolve (src/island.cpp:428-652) and of the archipelago's
// migrants database (src/archipelago.cpp:658-714): an island's population (ids | x | f) stays in HBM between evolve() calls;
// select_best (src/s_policies/select_best.cpp:63-171) packs the emigrants into a device "outbox", pgc_migrate moves outboxes
// along topology edges device to device - ncclSend / ncclRecv inside one group, or a device-to-device copy when both islands
// share a GPU - into "inbox" slots of the destination islands, and fair_replace (src/r_policies/fair_replace.cpp:63-221)
// merges an inbox into the population.  WHICH edges carry migrants in a round (topology, Bernoulli(weight), p2p / broadcast,
// preserve / evict: island.cpp:461-620) is decided by the host caller (include/pagmo_cuda/cuda_island.hpp), as in the reference.
//
// Packed group layout (doubles): [0] = number of rows k, then k ids (the u64 bit patterns), k x nx decision vectors, k x nf
// fitness vectors at fixed offsets for the island's capacity `cap` (so a whole group is one contiguous message of
// 1 + cap * (1 + nx + nf) doubles: 53 doubles for one CEC2013 D=50 migrant).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the process's already-loaded copy when the host application brought one),
// so libpgc.so loads on machines without NCCL and everything else keeps working; pgc_comm_* then fail with PGC_ERR_UNSUPPORTED.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <vector>

#include "pgc_internal.cuh"

struct pgc_island {
    pgc_problem *prob = nullptr;
    pgc_ctx *ctx = nullptr;
    size_t n = 0, nx = 0, nf = 0, cap = 0, slots = 0;
    unsigned generation = 1; // Philox generation counter: successive evolve() calls continue the random stream
    unsigned long long *d_ids = nullptr;
    double *d_x = nullptr, *d_f = nullptr;
    double *d_outbox = nullptr; // one packed group
    double *d_inbox = nullptr;  // `slots` packed groups
    // merged immigrants (rows of all used inbox slots, in slot order) for fair_replace + the acceptance flags of the migration log
    unsigned long long *d_mids = nullptr;
    double *d_mx = nullptr, *d_mf = nullptr;
    unsigned char *d_flags = nullptr;
    // what an algorithm built with memory = true keeps between evolve() calls (pgc_algo_memory), resident like the population
    pgc_algo_memory mem{};
    int mem_algo = 0;
    std::vector<double> es_state; // cmaes / xnes with memory: their host-side state (pgc_es_state_len)
    double *h_heads = nullptr; // pinned: slot headers
    unsigned char *h_flags = nullptr;
    unsigned long long *h_mids = nullptr; // pinned: the immigrants' ids of the last replace (migration log)
    cudaEvent_t ev = nullptr;
    cudaEvent_t ev_log = nullptr;         // recorded after the log copies of the last replace
    std::vector<size_t> log_cnt;          // rows per slot of the last replace whose log has not been collected yet
    bool log_pending = false;
    size_t group_doubles() const { return 1 + cap * (1 + nx + nf); }
    unsigned long long *ids_of(double *g) const { return reinterpret_cast<unsigned long long *>(g + 1); }
    double *x_of(double *g) const { return g + 1 + cap; }
    double *f_of(double *g) const { return g + 1 + cap + cap * nx; }
};

struct pgc_comm {
    int nranks = 0;
    struct Local {
        int device, rank;
        ncclComm_t comm;
    };
    std::vector<Local> local;
    const Local *find(int rank) const
    {
        for (const auto &l : local)
            if (l.rank == rank) return &l;
        return nullptr;
    }
};

namespace pgc
{
namespace
{
// ---- NCCL, bound at run time ---------------------------------------------------------------------------------------------
struct Nccl {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string why;
};

Nccl &nccl()
{
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = std::getenv("PGC_NCCL_LIBRARY");
        const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm || !*nm) continue;
            n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.handle) break;
            n.why = dlerror();
        }
        if (!n.handle) return;
#define PGC_NCCL_SYM(name)                                                                                             \
    n.name = reinterpret_cast<decltype(n.name)>(dlsym(n.handle, "nccl" #name));                                        \
    if (!n.name) {                                                                                                     \
        n.why = "symbol nccl" #name " not found";                                                                      \
        n.handle = nullptr;                                                                                            \
        return;                                                                                                        \
    }
        PGC_NCCL_SYM(GetUniqueId)
        PGC_NCCL_SYM(CommInitRank)
        PGC_NCCL_SYM(CommInitAll)
        PGC_NCCL_SYM(CommDestroy)
        PGC_NCCL_SYM(GroupStart)
        PGC_NCCL_SYM(GroupEnd)
        PGC_NCCL_SYM(Send)
        PGC_NCCL_SYM(Recv)
        PGC_NCCL_SYM(GetErrorString)
        PGC_NCCL_SYM(GetVersion)
#undef PGC_NCCL_SYM
    });
    return n;
}

int need_nccl(const char *who)
{
    if (nccl().handle) return PGC_OK;
    set_error("%s: NCCL is not available (%s); set PGC_NCCL_LIBRARY to the path of libnccl.so.2", who, nccl().why.c_str());
    return PGC_ERR_UNSUPPORTED;
}

#define PGC_NCCL(call)                                                                                                 \
    do {                                                                                                               \
        ncclResult_t r__ = (call);                                                                                     \
        if (r__ != ncclSuccess) {                                                                                      \
            set_error("NCCL error %d (%s) in `%s` at %s:%d", static_cast<int>(r__), nccl().GetErrorString(r__), #call, __FILE__, __LINE__); \
            return PGC_ERR_CUDA;                                                                                       \
        }                                                                                                              \
    } while (0)

__global__ void set_header_kernel(double *group, double k) { group[0] = k; }

// flag[j] = 1 when immigrant j's id is in the population now (the migration log's "made it in", island.cpp:525-536)
__global__ void accepted_kernel(const unsigned long long *__restrict__ ids, unsigned n, const unsigned long long *__restrict__ mids, unsigned nm,
                                unsigned char *flag)
{
    const unsigned j = blockIdx.x;
    if (j >= nm) return;
    __shared__ int hit;
    if (threadIdx.x == 0) hit = 0;
    __syncthreads();
    const unsigned long long want = mids[j];
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x)
        if (ids[i] == want) hit = 1;
    __syncthreads();
    if (threadIdx.x == 0) flag[j] = static_cast<unsigned char>(hit);
}
} // namespace
} // namespace pgc

using namespace pgc;

extern "C" {

// ---- islands ------------------------------------------------------------------------------------------------------------
int pgc_island_create(pgc_problem *prob, size_t n, size_t max_migrants, size_t max_in_edges, pgc_island **out)
{
    PGC_REQUIRE(prob && out, "pgc_island_create: null argument");
    PGC_REQUIRE(n >= 1 && n < 0xffffffffull, "pgc_island_create: population size %zu out of range", n);
    // the resident island evolves with the device UDAs, none of which takes constraints (as in the reference): unconstrain first
    PGC_REQUIRE(prob->nec + prob->nic == 0, "pgc_island_create: '%s' has %zu constraints; wrap it with pgc_problem_unconstrain",
                prob->name.c_str(), prob->nec + prob->nic);
    *out = nullptr;
    pgc_island *isl = new (std::nothrow) pgc_island;
    if (!isl) return PGC_ERR_OUT_OF_MEMORY;
    isl->prob = prob;
    isl->ctx = prob->ctx;
    isl->n = n;
    isl->nx = prob->nx;
    isl->nf = prob->nobj;
    isl->cap = max_migrants ? max_migrants : 1;
    isl->slots = max_in_edges ? max_in_edges : 1;
    PGC_CUDA(cudaSetDevice(isl->ctx->device));
    const size_t gd = isl->group_doubles(), mcap = isl->cap * isl->slots;
    int rc = PGC_OK;
    auto dev = [&](void **p, size_t bytes) {
        if (rc == PGC_OK && cudaMalloc(p, bytes ? bytes : 8) != cudaSuccess) {
            cudaGetLastError();
            set_error("pgc_island_create: out of device memory (%zu bytes)", bytes);
            rc = PGC_ERR_OUT_OF_MEMORY;
        }
    };
    dev(reinterpret_cast<void **>(&isl->d_ids), 8 * n);
    dev(reinterpret_cast<void **>(&isl->d_x), 8 * n * isl->nx);
    dev(reinterpret_cast<void **>(&isl->d_f), 8 * n * isl->nf);
    dev(reinterpret_cast<void **>(&isl->d_outbox), 8 * gd);
    dev(reinterpret_cast<void **>(&isl->d_inbox), 8 * gd * isl->slots);
    dev(reinterpret_cast<void **>(&isl->d_mids), 8 * mcap);
    dev(reinterpret_cast<void **>(&isl->d_mx), 8 * mcap * isl->nx);
    dev(reinterpret_cast<void **>(&isl->d_mf), 8 * mcap * isl->nf);
    dev(reinterpret_cast<void **>(&isl->d_flags), mcap);
    if (rc == PGC_OK && (cudaMallocHost(&isl->h_heads, 8 * isl->slots) != cudaSuccess || cudaMallocHost(&isl->h_flags, mcap) != cudaSuccess
                         || cudaMallocHost(&isl->h_mids, 8 * mcap) != cudaSuccess
                         || cudaEventCreateWithFlags(&isl->ev, cudaEventDisableTiming) != cudaSuccess
                         || cudaEventCreateWithFlags(&isl->ev_log, cudaEventDisableTiming) != cudaSuccess)) {
        set_error("pgc_island_create: pinned allocation failed");
        rc = PGC_ERR_OUT_OF_MEMORY;
    }
    if (rc == PGC_OK && (cudaMemsetAsync(isl->d_outbox, 0, 8 * gd, isl->ctx->stream) != cudaSuccess
                         || cudaMemsetAsync(isl->d_inbox, 0, 8 * gd * isl->slots, isl->ctx->stream) != cudaSuccess))
        rc = PGC_ERR_CUDA;
    if (rc != PGC_OK) {
        pgc_island_destroy(isl);
        return rc;
    }
    *out = isl;
    return PGC_OK;
}

int pgc_island_destroy(pgc_island *isl)
{
    if (!isl) return PGC_OK;
    cudaSetDevice(isl->ctx->device);
    cudaStreamSynchronize(isl->ctx->stream);
    for (void *p : {static_cast<void *>(isl->d_ids), static_cast<void *>(isl->d_x), static_cast<void *>(isl->d_f), static_cast<void *>(isl->d_outbox),
                    static_cast<void *>(isl->d_inbox), static_cast<void *>(isl->d_mids), static_cast<void *>(isl->d_mx),
                    static_cast<void *>(isl->d_mf), static_cast<void *>(isl->d_flags), static_cast<void *>(isl->mem.a),
                    static_cast<void *>(isl->mem.b), static_cast<void *>(isl->mem.c), static_cast<void *>(isl->mem.u)})
        if (p) cudaFree(p);
    if (isl->h_heads) cudaFreeHost(isl->h_heads);
    if (isl->h_flags) cudaFreeHost(isl->h_flags);
    if (isl->h_mids) cudaFreeHost(isl->h_mids);
    if (isl->ev) cudaEventDestroy(isl->ev);
    if (isl->ev_log) cudaEventDestroy(isl->ev_log);
    delete isl;
    return PGC_OK;
}

int pgc_island_size(const pgc_island *isl, size_t *n, size_t *nx, size_t *nf)
{
    PGC_REQUIRE(isl, "pgc_island_size: null island");
    if (n) *n = isl->n;
    if (nx) *nx = isl->nx;
    if (nf) *nf = isl->nf;
    return PGC_OK;
}

int pgc_island_pointers(pgc_island *isl, uint64_t **d_ids, double **d_x, double **d_f)
{
    PGC_REQUIRE(isl, "pgc_island_pointers: null island");
    if (d_ids) *d_ids = reinterpret_cast<uint64_t *>(isl->d_ids);
    if (d_x) *d_x = isl->d_x;
    if (d_f) *d_f = isl->d_f;
    return PGC_OK;
}

int pgc_island_generation(const pgc_island *isl, uint32_t *generation)
{
    PGC_REQUIRE(isl && generation, "pgc_island_generation: null argument");
    *generation = isl->generation;
    return PGC_OK;
}

int pgc_island_set_generation(pgc_island *isl, uint32_t generation)
{
    PGC_REQUIRE(isl, "pgc_island_set_generation: null island");
    isl->generation = generation;
    return PGC_OK;
}

int pgc_island_upload(pgc_island *isl, const uint64_t *ids, const double *x, const double *f)
{
    PGC_REQUIRE(isl, "pgc_island_upload: null island");
    PGC_CUDA(cudaSetDevice(isl->ctx->device));
    cudaStream_t st = isl->ctx->stream;
    if (ids) PGC_CUDA(cudaMemcpyAsync(isl->d_ids, ids, 8 * isl->n, cudaMemcpyHostToDevice, st));
    if (x) PGC_CUDA(cudaMemcpyAsync(isl->d_x, x, 8 * isl->n * isl->nx, cudaMemcpyHostToDevice, st));
    if (f) PGC_CUDA(cudaMemcpyAsync(isl->d_f, f, 8 * isl->n * isl->nf, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaStreamSynchronize(st)); // the host arrays may be pageable and are the caller's
    return PGC_OK;
}

int pgc_island_download(pgc_island *isl, uint64_t *ids, double *x, double *f)
{
    PGC_REQUIRE(isl, "pgc_island_download: null island");
    PGC_CUDA(cudaSetDevice(isl->ctx->device));
    cudaStream_t st = isl->ctx->stream;
    if (ids) PGC_CUDA(cudaMemcpyAsync(ids, isl->d_ids, 8 * isl->n, cudaMemcpyDeviceToHost, st));
    if (x) PGC_CUDA(cudaMemcpyAsync(x, isl->d_x, 8 * isl->n * isl->nx, cudaMemcpyDeviceToHost, st));
    if (f) PGC_CUDA(cudaMemcpyAsync(f, isl->d_f, 8 * isl->n * isl->nf, cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    return PGC_OK;
}

int pgc_island_init(pgc_island *isl, uint64_t seed)
{
    PGC_REQUIRE(isl, "pgc_island_init: null island");
    return pgc_population_init_device(isl->prob, isl->n, seed, isl->d_x, isl->d_f, reinterpret_cast<uint64_t *>(isl->d_ids), nullptr);
}

int pgc_island_evolve(pgc_island *isl, const pgc_algo_desc *algo, unsigned *gens_done)
{
    PGC_REQUIRE(isl && algo, "pgc_island_evolve: null argument");
    unsigned done = 0;
    int rc;
    const bool keeps_state = algo->memory
                             && (algo->algo == PGC_ALGO_SADE || algo->algo == PGC_ALGO_DE1220 || algo->algo == PGC_ALGO_PSO_GEN
                                 || algo->algo == PGC_ALGO_NSPSO || algo->algo == PGC_ALGO_CMAES || algo->algo == PGC_ALGO_XNES);
    if (keeps_state) {
        if (isl->mem_algo != algo->algo) { // another algorithm took the island over: its first evolve() starts from nothing
            isl->mem.initialized = 0;
            isl->mem_algo = algo->algo;
        }
        if (!isl->mem.a) {
            PGC_CUDA(cudaSetDevice(isl->ctx->device));
            const size_t n = isl->n ? isl->n : 1u;
            PGC_CUDA(cudaMalloc(&isl->mem.a, sizeof(double) * n * isl->nx));
            PGC_CUDA(cudaMalloc(&isl->mem.b, sizeof(double) * n * isl->nx));
            PGC_CUDA(cudaMalloc(&isl->mem.c, sizeof(double) * n * isl->nf));
            PGC_CUDA(cudaMalloc(&isl->mem.u, sizeof(uint32_t) * n));
        }
        if (algo->algo == PGC_ALGO_CMAES || algo->algo == PGC_ALGO_XNES) {
            size_t len = 0;
            if ((rc = pgc_es_state_len(algo->algo, isl->nx, &len))) return rc;
            if (isl->es_state.size() != len) {
                isl->es_state.assign(len, 0.);
                isl->mem.initialized = 0;
            }
            isl->mem.h_state = isl->es_state.data();
            isl->mem.h_state_len = len;
        }
        rc = pgc_algo_evolve_memory_device(isl->prob, algo, isl->d_x, isl->d_f, isl->n, isl->generation, &done, &isl->mem, nullptr);
    } else {
        rc = pgc_algo_evolve_device(isl->prob, algo, isl->d_x, isl->d_f, isl->n, isl->generation, &done, nullptr);
    }
    if (rc != PGC_OK) return rc;
    isl->generation += algo->gens ? algo->gens : 1u;
    if (gens_done) *gens_done = done;
    return PGC_OK;
}

int pgc_island_select(pgc_island *isl, int rate_is_frac, double rate, size_t *k_out)
{
    PGC_REQUIRE(isl, "pgc_island_select: null island");
    PGC_CUDA(cudaSetDevice(isl->ctx->device));
    cudaStream_t st = isl->ctx->stream;
    size_t k = 0;
    int rc = policy_rate_count("Select best", rate_is_frac, rate, isl->n, &k);
    if (rc != PGC_OK) return rc;
    PGC_REQUIRE(k <= isl->cap, "pgc_island_select: the policy selects %zu individuals but the island was created for at most %zu migrants", k,
                isl->cap);
    double *g = isl->d_outbox;
    if ((rc = select_best_policy_device(isl->ctx, isl->d_ids, isl->d_x, isl->d_f, isl->n, isl->nx, isl->nf, rate_is_frac, rate, isl->ids_of(g),
                                        isl->x_of(g), isl->f_of(g), &k, st)))
        return rc;
    set_header_kernel<<<1, 1, 0, st>>>(g, static_cast<double>(k));
    PGC_CUDA(cudaGetLastError());
    isl->ctx->launches.fetch_add(1, std::memory_order_relaxed);
    if (k_out) *k_out = k;
    return PGC_OK;
}

int pgc_island_clear_outbox(pgc_island *isl) // archipelago::extract_migrants leaves an empty entry (archipelago.cpp:690-714)
{
    PGC_REQUIRE(isl, "pgc_island_clear_outbox: null island");
    PGC_CUDA(cudaSetDevice(isl->ctx->device));
    set_header_kernel<<<1, 1, 0, isl->ctx->stream>>>(isl->d_outbox, 0.);
    PGC_CUDA(cudaGetLastError());
    return PGC_OK;
}

static int download_group(pgc_island *isl, double *g, uint64_t *ids, double *x, double *f, size_t *k_out)
{
    cudaStream_t st = isl->ctx->stream;
    PGC_CUDA(cudaMemcpyAsync(isl->h_heads, g, 8, cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    const size_t k = static_cast<size_t>(isl->h_heads[0]);
    PGC_REQUIRE(k <= isl->cap, "island: corrupt group header (%zu rows, capacity %zu)", k, isl->cap);
    if (k) {
        if (ids) PGC_CUDA(cudaMemcpyAsync(ids, isl->ids_of(g), 8 * k, cudaMemcpyDeviceToHost, st));
        if (x) PGC_CUDA(cudaMemcpyAsync(x, isl->x_of(g), 8 * k * isl->nx, cudaMemcpyDeviceToHost, st));
        if (f) PGC_CUDA(cudaMemcpyAsync(f, isl->f_of(g), 8 * k * isl->nf, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
    }
    if (k_out) *k_out = k;
    return PGC_OK;
}

int pgc_island_outbox_download(pgc_island *isl, uint64_t *ids, double *x, double *f, size_t *k)
{
    PGC_REQUIRE(isl, "pgc_island_outbox_download: null island");
    PGC_CUDA(cudaSetDevice(isl->ctx->device));
    return download_group(isl, isl->d_outbox, ids, x, f, k);
}

int pgc_island_inbox_upload(pgc_island *isl, size_t slot, const uint64_t *ids, const double *x, const double *f, size_t k)
{
    PGC_REQUIRE(isl && slot < isl->slots && k <= isl->cap, "pgc_island_inbox_upload: slot %zu / %zu rows out of range (slots %zu, capacity %zu)",
                slot, k, isl ? isl->slots : 0, isl ? isl->cap : 0);
    PGC_CUDA(cudaSetDevice(isl->ctx->device));
    cudaStream_t st = isl->ctx->stream;
    double *g = isl->d_inbox + slot * isl->group_doubles();
    if (k) {
        PGC_REQUIRE(ids && x && f, "pgc_island_inbox_upload: null rows");
        PGC_CUDA(cudaMemcpyAsync(isl->ids_of(g), ids, 8 * k, cudaMemcpyHostToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(isl->x_of(g), x, 8 * k * isl->nx, cudaMemcpyHostToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(isl->f_of(g), f, 8 * k * isl->nf, cudaMemcpyHostToDevice, st));
    }
    set_header_kernel<<<1, 1, 0, st>>>(g, static_cast<double>(k));
    PGC_CUDA(cudaGetLastError());
    PGC_CUDA(cudaStreamSynchronize(st));
    return PGC_OK;
}

// r_policy step: the rows of inbox slots [0, n_slots) are the immigrants of this round (island.cpp:505-517 / :577-585).
// accepted_* (optional, sized cap * slots): ids and source slots of the immigrants that are in the population afterwards.
// The two halves of a replace.  enqueue: everything that touches the device, asynchronous on the island's stream - the rows of the
// inbox slots merged in slot order, fair_replace, and (want_log) the acceptance flags and immigrant ids copied to pinned memory behind
// an event.  collect: wait for that event only and decode the migration-log rows.  A driver that knows the slot counts (the senders'
// policies fix them) never blocks between the end of one evolve() and the start of the next: pgc_island_replace_enqueue, evolve,
// select, and the log of this round read at the start of the next one (pgc_island_replace_collect).
static int replace_enqueue(pgc_island *isl, int rate_is_frac, double rate, size_t n_slots, const size_t *cnt, bool want_log)
{
    cudaStream_t st = isl->ctx->stream;
    const size_t gd = isl->group_doubles();
    PGC_REQUIRE(!isl->log_pending, "pgc_island_replace: the migration log of the previous replace was not collected (pgc_island_replace_collect)");
    size_t nm = 0;
    for (size_t s = 0; s < n_slots; ++s) {
        PGC_REQUIRE(cnt[s] <= isl->cap, "pgc_island_replace: corrupt inbox header in slot %zu (%zu rows, capacity %zu)", s, cnt[s], isl->cap);
        double *g = isl->d_inbox + s * gd;
        if (cnt[s]) {
            PGC_CUDA(cudaMemcpyAsync(isl->d_mids + nm, isl->ids_of(g), 8 * cnt[s], cudaMemcpyDeviceToDevice, st));
            PGC_CUDA(cudaMemcpyAsync(isl->d_mx + nm * isl->nx, isl->x_of(g), 8 * cnt[s] * isl->nx, cudaMemcpyDeviceToDevice, st));
            PGC_CUDA(cudaMemcpyAsync(isl->d_mf + nm * isl->nf, isl->f_of(g), 8 * cnt[s] * isl->nf, cudaMemcpyDeviceToDevice, st));
        }
        nm += cnt[s];
    }
    int rc = fair_replace_policy_device(isl->ctx, isl->d_ids, isl->d_x, isl->d_f, isl->n, isl->nx, isl->nf, rate_is_frac, rate, isl->d_mids,
                                        isl->d_mx, isl->d_mf, nm, st);
    if (rc != PGC_OK) return rc;
    if (nm && want_log) {
        accepted_kernel<<<static_cast<unsigned>(nm), 128, 0, st>>>(isl->d_ids, static_cast<unsigned>(isl->n), isl->d_mids,
                                                                    static_cast<unsigned>(nm), isl->d_flags);
        PGC_CUDA(cudaGetLastError());
        isl->ctx->launches.fetch_add(1, std::memory_order_relaxed);
        PGC_CUDA(cudaMemcpyAsync(isl->h_flags, isl->d_flags, nm, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaMemcpyAsync(isl->h_mids, isl->d_mids, 8 * nm, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaEventRecord(isl->ev_log, st));
        isl->log_cnt.assign(cnt, cnt + n_slots);
        isl->log_pending = true;
    }
    return PGC_OK;
}

static int replace_collect(pgc_island *isl, uint64_t *accepted_ids, uint32_t *accepted_slot, size_t *n_accepted)
{
    if (n_accepted) *n_accepted = 0;
    if (!isl->log_pending) return PGC_OK;
    PGC_CUDA(cudaEventSynchronize(isl->ev_log));
    isl->log_pending = false;
    size_t a = 0, j = 0;
    for (size_t s = 0; s < isl->log_cnt.size(); ++s)
        for (size_t r = 0; r < isl->log_cnt[s]; ++r, ++j)
            if (isl->h_flags[j]) {
                if (accepted_ids) accepted_ids[a] = isl->h_mids[j];
                if (accepted_slot) accepted_slot[a] = static_cast<uint32_t>(s);
                ++a;
            }
    if (n_accepted) *n_accepted = a;
    return PGC_OK;
}

int pgc_island_replace(pgc_island *isl, int rate_is_frac, double rate, size_t n_slots, uint64_t *accepted_ids, uint32_t *accepted_slot,
                       size_t *n_accepted)
{
    PGC_REQUIRE(isl && n_slots <= isl->slots, "pgc_island_replace: %zu slots requested, the island has %zu", n_slots, isl ? isl->slots : 0);
    if (n_accepted) *n_accepted = 0;
    PGC_CUDA(cudaSetDevice(isl->ctx->device));
    cudaStream_t st = isl->ctx->stream;
    const size_t gd = isl->group_doubles();
    // the counts arrive with the rows (a sender's policy fixes them, but the receiver may sit in another process)
    PGC_CUDA(cudaMemcpy2DAsync(isl->h_heads, 8, isl->d_inbox, 8 * gd, 8, n_slots ? n_slots : 1, cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    std::vector<size_t> cnt(n_slots);
    for (size_t s = 0; s < n_slots; ++s) cnt[s] = static_cast<size_t>(isl->h_heads[s]);
    const bool want_log = accepted_ids || n_accepted;
    int rc = replace_enqueue(isl, rate_is_frac, rate, n_slots, cnt.data(), want_log);
    if (rc != PGC_OK) return rc;
    return want_log ? replace_collect(isl, accepted_ids, accepted_slot, n_accepted) : PGC_OK;
}

int pgc_island_replace_enqueue(pgc_island *isl, int rate_is_frac, double rate, size_t n_slots, const size_t *counts, int want_log)
{
    PGC_REQUIRE(isl && n_slots <= isl->slots, "pgc_island_replace_enqueue: %zu slots requested, the island has %zu", n_slots, isl ? isl->slots : 0);
    PGC_REQUIRE(counts || n_slots == 0, "pgc_island_replace_enqueue: null counts");
    PGC_CUDA(cudaSetDevice(isl->ctx->device));
    return replace_enqueue(isl, rate_is_frac, rate, n_slots, counts, want_log != 0);
}

int pgc_island_replace_collect(pgc_island *isl, uint64_t *accepted_ids, uint32_t *accepted_slot, size_t *n_accepted)
{
    PGC_REQUIRE(isl, "pgc_island_replace_collect: null island");
    PGC_CUDA(cudaSetDevice(isl->ctx->device));
    return replace_collect(isl, accepted_ids, accepted_slot, n_accepted);
}

// best individual of a single-objective island (population::champion_x / champion_f of what the island holds now)
int pgc_island_champion(pgc_island *isl, double *x, double *f)
{
    PGC_REQUIRE(isl && isl->nf == 1, "pgc_island_champion: the champion is defined for single-objective islands only");
    PGC_CUDA(cudaSetDevice(isl->ctx->device));
    cudaStream_t st = isl->ctx->stream;
    unsigned *d_sel = nullptr;
    PGC_CUDA(cudaMallocAsync(&d_sel, sizeof(unsigned), st));
    int rc = so_best_indices_device(isl->ctx, isl->d_f, isl->n, 1, d_sel, st);
    unsigned sel = 0;
    if (rc == PGC_OK && cudaMemcpyAsync(&sel, d_sel, sizeof(unsigned), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = PGC_ERR_CUDA;
    if (rc == PGC_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = PGC_ERR_CUDA;
    cudaFreeAsync(d_sel, st);
    if (rc != PGC_OK) return rc;
    if (x) PGC_CUDA(cudaMemcpyAsync(x, isl->d_x + static_cast<size_t>(sel) * isl->nx, 8 * isl->nx, cudaMemcpyDeviceToHost, st));
    if (f) PGC_CUDA(cudaMemcpyAsync(f, isl->d_f + sel, 8, cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    return PGC_OK;
}

// ---- communicators ------------------------------------------------------------------------------------------------------
int pgc_comm_nccl_version(int *version)
{
    PGC_REQUIRE(version, "pgc_comm_nccl_version: null output");
    int rc = need_nccl("pgc_comm_nccl_version");
    if (rc != PGC_OK) return rc;
    PGC_NCCL(nccl().GetVersion(version));
    return PGC_OK;
}

int pgc_comm_init(int ndev, const int *devices, pgc_comm **out)
{
    PGC_REQUIRE(out && ndev >= 1 && devices, "pgc_comm_init: bad arguments");
    *out = nullptr;
    int rc = need_nccl("pgc_comm_init");
    if (rc != PGC_OK) return rc;
    int count = 0;
    PGC_CUDA(cudaGetDeviceCount(&count));
    for (int i = 0; i < ndev; ++i) {
        PGC_REQUIRE(devices[i] >= 0 && devices[i] < count, "pgc_comm_init: device %d out of range (%d CUDA devices visible)", devices[i], count);
        for (int j = 0; j < i; ++j) PGC_REQUIRE(devices[i] != devices[j], "pgc_comm_init: device %d listed twice", devices[i]);
    }
    std::vector<ncclComm_t> comms(static_cast<size_t>(ndev));
    PGC_NCCL(nccl().CommInitAll(comms.data(), ndev, devices));
    pgc_comm *c = new pgc_comm;
    c->nranks = ndev;
    for (int i = 0; i < ndev; ++i) c->local.push_back({devices[i], i, comms[static_cast<size_t>(i)]});
    *out = c;
    return PGC_OK;
}

int pgc_comm_unique_id(void *id, size_t len)
{
    PGC_REQUIRE(id && len >= sizeof(ncclUniqueId), "pgc_comm_unique_id: the buffer must hold %zu bytes", sizeof(ncclUniqueId));
    int rc = need_nccl("pgc_comm_unique_id");
    if (rc != PGC_OK) return rc;
    ncclUniqueId u;
    PGC_NCCL(nccl().GetUniqueId(&u));
    std::memcpy(id, &u, sizeof(u));
    return PGC_OK;
}

int pgc_comm_init_rank(int device, int nranks, int rank, const void *id, size_t len, pgc_comm **out)
{
    PGC_REQUIRE(out && id && len >= sizeof(ncclUniqueId) && nranks >= 1 && rank >= 0 && rank < nranks, "pgc_comm_init_rank: bad arguments");
    *out = nullptr;
    int rc = need_nccl("pgc_comm_init_rank");
    if (rc != PGC_OK) return rc;
    PGC_CUDA(cudaSetDevice(device));
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    ncclComm_t comm;
    PGC_NCCL(nccl().CommInitRank(&comm, nranks, u, rank));
    pgc_comm *c = new pgc_comm;
    c->nranks = nranks;
    c->local.push_back({device, rank, comm});
    *out = c;
    return PGC_OK;
}

int pgc_comm_destroy(pgc_comm *c)
{
    if (!c) return PGC_OK;
    for (auto &l : c->local) {
        cudaSetDevice(l.device);
        cudaDeviceSynchronize();
        nccl().CommDestroy(l.comm);
    }
    delete c;
    return PGC_OK;
}

int pgc_comm_size(const pgc_comm *c, int *nranks, int *nlocal)
{
    PGC_REQUIRE(c, "pgc_comm_size: null communicator");
    if (nranks) *nranks = c->nranks;
    if (nlocal) *nlocal = static_cast<int>(c->local.size());
    return PGC_OK;
}

// One migration step: for every edge e, the outbox of island edge_src[e] is delivered into inbox slot edge_slot[e] of island
// edge_dst[e].  islands[i] is NULL for islands owned by another process; owner_rank[i] = the communicator rank whose GPU holds
// island i.  Every process passes the SAME edge list (order matters: NCCL pairs the sends and receives of two ranks in issue
// order).  comm may be NULL when every edge stays inside one GPU.  Asynchronous: transfers are ordered on the islands' streams
// (the outbox of the source after its select, the inbox of the destination before its replace).
int pgc_migrate(pgc_comm *comm, pgc_island *const *islands, const int *owner_rank, size_t n_islands, const uint32_t *edge_src,
                const uint32_t *edge_dst, const uint32_t *edge_slot, size_t n_edges)
{
    PGC_REQUIRE(islands && owner_rank && (n_edges == 0 || (edge_src && edge_dst && edge_slot)), "pgc_migrate: null argument");
    bool grouped = false;
    auto fail = [&](int rc) {
        if (grouped) nccl().GroupEnd();
        return rc;
    };
    for (size_t e = 0; e < n_edges; ++e) {
        const uint32_t s = edge_src[e], d = edge_dst[e];
        PGC_REQUIRE(s < n_islands && d < n_islands, "pgc_migrate: edge %zu (%u -> %u) names an island outside [0, %zu)", e, s, d, n_islands);
        pgc_island *src = islands[s], *dst = islands[d];
        if (!src && !dst) continue; // both ends live elsewhere
        pgc_island *any = src ? src : dst;
        if (dst) PGC_REQUIRE(edge_slot[e] < dst->slots, "pgc_migrate: edge %zu targets inbox slot %u, island %u has %zu", e, edge_slot[e], d, dst->slots);
        if (src && dst) {
            PGC_REQUIRE(src->group_doubles() == dst->group_doubles(), "pgc_migrate: islands %u and %u have different group layouts", s, d);
        }
        const size_t gd = any->group_doubles();
        if (owner_rank[s] == owner_rank[d]) { // same GPU: a device-to-device copy ordered after the source's stream
            if (!src || !dst) return fail((set_error("pgc_migrate: islands %u and %u share rank %d but only one is local", s, d, owner_rank[s]),
                                           PGC_ERR_INVALID_ARGUMENT));
            if (cudaSetDevice(dst->ctx->device) != cudaSuccess || cudaEventRecord(src->ev, src->ctx->stream) != cudaSuccess
                || cudaStreamWaitEvent(dst->ctx->stream, src->ev, 0) != cudaSuccess
                || cudaMemcpyAsync(dst->d_inbox + edge_slot[e] * gd, src->d_outbox, 8 * gd, cudaMemcpyDeviceToDevice, dst->ctx->stream)
                       != cudaSuccess)
                return fail(cuda_fail(cudaGetLastError(), "pgc_migrate: device-to-device copy", __FILE__, __LINE__));
            continue;
        }
        if (!comm) return fail((set_error("pgc_migrate: edge %u -> %u crosses GPUs but no communicator was given", s, d), PGC_ERR_INVALID_ARGUMENT));
        if (!grouped) {
            int rc = need_nccl("pgc_migrate");
            if (rc != PGC_OK) return rc;
            PGC_NCCL(nccl().GroupStart());
            grouped = true;
        }
        if (src) {
            const pgc_comm::Local *l = comm->find(owner_rank[s]);
            if (!l || l->device != src->ctx->device)
                return fail((set_error("pgc_migrate: island %u is local but rank %d is not one of this process's communicator ranks on device %d", s,
                                       owner_rank[s], src->ctx->device),
                             PGC_ERR_INVALID_ARGUMENT));
            cudaSetDevice(l->device);
            ncclResult_t r = nccl().Send(src->d_outbox, gd, ncclDouble, owner_rank[d], l->comm, src->ctx->stream);
            if (r != ncclSuccess) return fail((set_error("pgc_migrate: ncclSend failed: %s", nccl().GetErrorString(r)), PGC_ERR_CUDA));
        }
        if (dst) {
            const pgc_comm::Local *l = comm->find(owner_rank[d]);
            if (!l || l->device != dst->ctx->device)
                return fail((set_error("pgc_migrate: island %u is local but rank %d is not one of this process's communicator ranks on device %d", d,
                                       owner_rank[d], dst->ctx->device),
                             PGC_ERR_INVALID_ARGUMENT));
            cudaSetDevice(l->device);
            ncclResult_t r = nccl().Recv(dst->d_inbox + edge_slot[e] * gd, gd, ncclDouble, owner_rank[s], l->comm, dst->ctx->stream);
            if (r != ncclSuccess) return fail((set_error("pgc_migrate: ncclRecv failed: %s", nccl().GetErrorString(r)), PGC_ERR_CUDA));
        }
    }
    if (grouped) PGC_NCCL(nccl().GroupEnd());
    return PGC_OK;
}

} // extern "C"
