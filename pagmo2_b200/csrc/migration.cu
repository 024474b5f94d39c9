// migration.cu - the migration step of an archipelago on device-resident populations (sm_100a).
//
// Replaces reference select_best::select (src/s_policies/select_best.cpp:63-171) and fair_replace::replace
// (src/r_policies/fair_replace.cpp:63-221) for unconstrained problems (every UDP of the device path), and the in-edge lists of
// ring / fully_connected (src/topologies/ring.cpp:74-116, fully_connected.cpp:86-115) that island::evolve consults
// (src/island.cpp:461-470).  Populations are flat row-major groups: ids[n] (u64), x[n x nx], f[n x nobj].
//   single objective : order by fitness, NaN last (detail::less_than_f) - a stable radix sort on order-preserving u64 keys
//   multi objective  : select_best_N_mo (mo_utils.cu)
// The reference sorts with std::sort (order of ties unspecified); here ties keep index order, residents before migrants.
// HBM-bound index work: bytes = 8*(n + k) keys + 8*(nx + nobj + 1) per moved row.
#include <cub/device/device_radix_sort.cuh>

#include <vector>

#include "pgc_internal.cuh"

namespace pgc
{

namespace
{

__device__ __forceinline__ double max0(double a) { return a < 0. ? 0. : a; } // std::max(a, 0.): a NaN stays a NaN (never satisfied)

__global__ void so_keys_kernel(const double *__restrict__ f, size_t stride, unsigned n, unsigned long long *keys, unsigned *idx)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = f[i * stride];
    unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
    // order-preserving map of IEEE doubles; every NaN goes to the top (less_than_f: NaN is greater than anything)
    b = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
    keys[i] = (v != v) ? 0xffffffffffffffffull : b;
    idx[i] = i;
}

// out row r = (r < n0 ? A : B)[sel[r]]: gathers ids / x / f rows of a merged group A (n0 rows) ++ B
__global__ void gather_rows_kernel(const unsigned *__restrict__ sel, unsigned rows, unsigned n0, unsigned nx, unsigned nf,
                                   const unsigned long long *idA, const double *xA, const double *fA, const unsigned long long *idB,
                                   const double *xB, const double *fB, unsigned long long *id_out, double *x_out, double *f_out)
{
    const unsigned r = blockIdx.x;
    if (r >= rows) return;
    const unsigned s = sel ? sel[r] : r;
    const bool a = s < n0;
    const unsigned src = a ? s : s - n0;
    const double *x = (a ? xA : xB) + static_cast<size_t>(src) * nx, *f = (a ? fA : fB) + static_cast<size_t>(src) * nf;
    for (unsigned j = threadIdx.x; j < nx; j += blockDim.x) x_out[static_cast<size_t>(r) * nx + j] = x[j];
    for (unsigned j = threadIdx.x; j < nf; j += blockDim.x) f_out[static_cast<size_t>(r) * nf + j] = f[j];
    if (threadIdx.x == 0) id_out[r] = (a ? idA : idB)[src];
}

struct Tmp { // stream-ordered scratch
    cudaStream_t st;
    std::vector<void *> ptrs;
    explicit Tmp(cudaStream_t s) : st(s) {}
    template <class T> int get(T **out, size_t count)
    {
        void *p = nullptr;
        PGC_CUDA(cudaMallocAsync(&p, (count ? count : 1) * sizeof(T), st));
        ptrs.push_back(p);
        *out = static_cast<T *>(p);
        return PGC_OK;
    }
    ~Tmp()
    {
        for (void *p : ptrs) cudaFreeAsync(p, st);
    }
};

// Constrained single-objective groups (select_best.cpp:137-152, fair_replace.cpp:158-188): rows [f | nec equality | nic inequality
// constraints], ordered as sort_population_con orders them (compare_fc, constrained.cpp:76-118): more satisfied constraints first;
// among the feasible, the smaller objective; among equally infeasible ones, the smaller violation norm.  The reference measures the
// left argument's violation as the SUM of its equality and inequality norms and the right argument's as their Euclidean combination;
// the two coincide whenever an individual violates constraints of one kind only, and the device uses the Euclidean norm throughout
// (a strict weak order, sortable by keys).  Keys: primary = number of violated constraints, secondary = objective or norm.
struct ConSpec {
    size_t nec = 0, nic = 0;
    const double *d_tol = nullptr; // device, [nec + nic]
    bool on() const { return nec + nic > 0; }
};
thread_local ConSpec tls_con; // set by the *_con entry points around a policy call

__global__ void con_keys_kernel(const double *__restrict__ f, unsigned n, unsigned nec, unsigned nic, const double *__restrict__ tol,
                                unsigned long long *primary, unsigned long long *secondary, unsigned *idx)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned nf = 1u + nec + nic;
    const double *row = f + static_cast<size_t>(i) * nf;
    double l2e = 0., l2i = 0.;
    unsigned nsat = 0;
    for (unsigned j = 0; j < nec; ++j) { // detail::test_eq_constraints, constrained.hpp:49-62
        const double err = max0(fabs(row[1u + j]) - tol[j]);
        l2e += err * err;
        nsat += err <= 0. ? 1u : 0u;
    }
    for (unsigned j = 0; j < nic; ++j) { // detail::test_ineq_constraints, :67-80
        const double err = max0(row[1u + nec + j] - tol[nec + j]);
        l2i += err * err;
        nsat += err <= 0. ? 1u : 0u;
    }
    const double e = sqrt(l2e), q = sqrt(l2i);
    const double v = nsat == nec + nic ? row[0] : sqrt(e * e + q * q);
    unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
    b = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
    secondary[i] = (v != v) ? 0xffffffffffffffffull : b;
    primary[i] = nec + nic - nsat;
    idx[i] = i;
}

__global__ void gather_keys_kernel(const unsigned long long *keys, const unsigned *idx, unsigned n, unsigned long long *out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = keys[idx[i]];
}

// d_sel[0..k) = the first k of sort_population_con(f): two stable radix passes, secondary key first
int con_best_indices(pgc_ctx *ctx, const double *d_f, size_t n, const ConSpec &con, size_t k, unsigned *d_sel, Tmp &tmp, cudaStream_t st)
{
    unsigned long long *prim = nullptr, *sec = nullptr, *k1 = nullptr, *k2 = nullptr;
    unsigned *i0 = nullptr, *i1 = nullptr, *i2 = nullptr;
    int rc;
    if ((rc = tmp.get(&prim, n)) || (rc = tmp.get(&sec, n)) || (rc = tmp.get(&k1, n)) || (rc = tmp.get(&k2, n)) || (rc = tmp.get(&i0, n))
        || (rc = tmp.get(&i1, n)) || (rc = tmp.get(&i2, n)))
        return rc;
    const unsigned un = static_cast<unsigned>(n);
    con_keys_kernel<<<(un + 255) / 256, 256, 0, st>>>(d_f, un, static_cast<unsigned>(con.nec), static_cast<unsigned>(con.nic), con.d_tol, prim, sec, i0);
    PGC_CUDA(cudaGetLastError());
    size_t bytes = 0;
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, sec, k1, i0, i1, static_cast<int>(n), 0, 64, st));
    unsigned char *ws = nullptr;
    if ((rc = tmp.get(&ws, bytes))) return rc;
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(ws, bytes, sec, k1, i0, i1, static_cast<int>(n), 0, 64, st));
    gather_keys_kernel<<<(un + 255) / 256, 256, 0, st>>>(prim, i1, un, k1);
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(ws, bytes, k1, k2, i1, i2, static_cast<int>(n), 0, 32, st));
    PGC_CUDA(cudaMemcpyAsync(d_sel, i2, k * sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
    ctx->launches.fetch_add(5, std::memory_order_relaxed);
    return PGC_OK;
}

// d_sel[0..k) = indices of the best k rows of f [n x nobj] (nobj = the row width: 1 + nec + nic for a constrained group)
int best_indices(pgc_ctx *ctx, const double *d_f, size_t n, size_t nobj, size_t k, unsigned *d_sel, Tmp &tmp, cudaStream_t st)
{
    if (k == 0 || n == 0) return PGC_OK;
    if (tls_con.on()) return con_best_indices(ctx, d_f, n, tls_con, k, d_sel, tmp, st);
    if (nobj > 1) {
        unsigned nout = 0;
        unsigned *full = nullptr;
        int rc = tmp.get(&full, n);
        if (rc != PGC_OK) return rc;
        rc = select_best_device(ctx, d_f, n, nobj, k, full, &nout, st);
        if (rc != PGC_OK) return rc;
        PGC_REQUIRE(nout == k, "migration: select_best_N_mo returned %u of %zu individuals", nout, k);
        PGC_CUDA(cudaMemcpyAsync(d_sel, full, k * sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
        return PGC_OK;
    }
    unsigned long long *k0 = nullptr, *k1 = nullptr;
    unsigned *i0 = nullptr, *i1 = nullptr;
    int rc;
    if ((rc = tmp.get(&k0, n)) || (rc = tmp.get(&k1, n)) || (rc = tmp.get(&i0, n)) || (rc = tmp.get(&i1, n))) return rc;
    const unsigned un = static_cast<unsigned>(n);
    so_keys_kernel<<<(un + 255) / 256, 256, 0, st>>>(d_f, 1, un, k0, i0);
    PGC_CUDA(cudaGetLastError());
    size_t bytes = 0;
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0, k1, i0, i1, static_cast<int>(n), 0, 64, st));
    unsigned char *ws = nullptr;
    if ((rc = tmp.get(&ws, bytes))) return rc;
    PGC_CUDA(cub::DeviceRadixSort::SortPairs(ws, bytes, k0, k1, i0, i1, static_cast<int>(n), 0, 64, st));
    PGC_CUDA(cudaMemcpyAsync(d_sel, i1, k * sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
    ctx->launches.fetch_add(3, std::memory_order_relaxed);
    return PGC_OK;
}

int rate_count(const char *who, int rate_is_frac, double rate, size_t n, size_t *out)
{
    if (rate_is_frac) { // base_sr_policy.cpp:46-55
        PGC_REQUIRE(rate >= 0.0 && rate <= 1.0 && rate == rate,
                    "Invalid fractional migration rate specified in the constructor of a replacement/selection policy: the rate must be in "
                    "the [0., 1.] range, but it is %g instead",
                    rate);
        const size_t c = static_cast<size_t>(rate * static_cast<double>(n));
        *out = c < n ? c : n;
        return PGC_OK;
    }
    PGC_REQUIRE(rate >= 0.0, "%s: negative absolute migration rate", who);
    const size_t c = static_cast<size_t>(rate);
    PGC_REQUIRE(c <= n, "The absolute migration rate (%zu) in a '%s' policy is larger than the number of input individuals (%zu)", c, who, n);
    *out = c;
    return PGC_OK;
}

} // namespace

// number of individuals a policy with this rate moves for a group of n (base_sr_policy.cpp:46-55, select_best.cpp:80-100)
int policy_rate_count(const char *who, int rate_is_frac, double rate, size_t n, size_t *out) { return rate_count(who, rate_is_frac, rate, n, out); }

// d_sel[0..k) = indices of the k smallest of the n single-objective fitness values d_f, in ascending order (NaN last, ties by index)
int so_best_indices_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t k, unsigned *d_sel, cudaStream_t st)
{
    Tmp tmp(st);
    return best_indices(ctx, d_f, n, 1, k, d_sel, tmp, st);
}

int select_best_policy_device(pgc_ctx *ctx, const unsigned long long *d_ids, const double *d_x, const double *d_f, size_t n, size_t nx,
                              size_t nobj, int rate_is_frac, double rate, unsigned long long *d_ids_out, double *d_x_out, double *d_f_out,
                              size_t *n_out, cudaStream_t st)
{
    size_t k = 0;
    int rc = rate_count("Select best", rate_is_frac, rate, n, &k); // select_best.cpp:80-100
    if (rc != PGC_OK) return rc;
    *n_out = k;
    if (k == 0) return PGC_OK;
    Tmp tmp(st);
    unsigned *sel = nullptr;
    if ((rc = tmp.get(&sel, k)) || (rc = best_indices(ctx, d_f, n, nobj, k, sel, tmp, st))) return rc;
    gather_rows_kernel<<<static_cast<unsigned>(k), 64, 0, st>>>(sel, static_cast<unsigned>(k), static_cast<unsigned>(n), static_cast<unsigned>(nx),
                                                                static_cast<unsigned>(nobj), d_ids, d_x, d_f, d_ids, d_x, d_f, d_ids_out, d_x_out,
                                                                d_f_out);
    PGC_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

int fair_replace_policy_device(pgc_ctx *ctx, unsigned long long *d_ids, double *d_x, double *d_f, size_t n, size_t nx, size_t nobj,
                               int rate_is_frac, double rate, const unsigned long long *d_mids, const double *d_mx, const double *d_mf,
                               size_t nm, cudaStream_t st)
{
    size_t k = 0;
    int rc = rate_count("fair_replace", rate_is_frac, rate, n, &k); // fair_replace.cpp:80-103
    if (rc != PGC_OK) return rc;
    if (k > nm) k = nm; // :105-107
    if (n == 0) return PGC_OK;
    Tmp tmp(st);
    // the top k migrants (:118-123 / :191), appended to the residents (:127-132)
    unsigned *top = nullptr, *keep = nullptr;
    unsigned long long *gid = nullptr, *oid = nullptr;
    double *gx = nullptr, *gf = nullptr, *ox = nullptr, *of = nullptr;
    const size_t tot = n + k;
    if ((rc = tmp.get(&top, k)) || (rc = tmp.get(&keep, n)) || (rc = tmp.get(&gid, k)) || (rc = tmp.get(&gx, k * nx)) || (rc = tmp.get(&gf, tot * nobj))
        || (rc = tmp.get(&oid, n)) || (rc = tmp.get(&ox, n * nx)) || (rc = tmp.get(&of, n * nobj)))
        return rc;
    if (k) {
        if ((rc = best_indices(ctx, d_mf, nm, nobj, k, top, tmp, st))) return rc;
        // merged fitness = residents ++ chosen migrants (rows n..tot-1 of gf); chosen migrants' ids / x in gid / gx
        gather_rows_kernel<<<static_cast<unsigned>(k), 64, 0, st>>>(top, static_cast<unsigned>(k), static_cast<unsigned>(nm), static_cast<unsigned>(nx),
                                                                    static_cast<unsigned>(nobj), d_mids, d_mx, d_mf, d_mids, d_mx, d_mf, gid, gx,
                                                                    gf + n * nobj);
        PGC_CUDA(cudaGetLastError());
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
    }
    PGC_CUDA(cudaMemcpyAsync(gf, d_f, n * nobj * sizeof(double), cudaMemcpyDeviceToDevice, st));
    // the best n of the merged group, in sorted order (:134-155 / :203-217): this REORDERS the island's population
    if ((rc = best_indices(ctx, gf, tot, nobj, n, keep, tmp, st))) return rc;
    gather_rows_kernel<<<static_cast<unsigned>(n), 64, 0, st>>>(keep, static_cast<unsigned>(n), static_cast<unsigned>(n), static_cast<unsigned>(nx),
                                                                static_cast<unsigned>(nobj), d_ids, d_x, d_f, gid, gx, gf + n * nobj, oid, ox, of);
    PGC_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    PGC_CUDA(cudaMemcpyAsync(d_ids, oid, n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(d_x, ox, n * nx * sizeof(double), cudaMemcpyDeviceToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(d_f, of, n * nobj * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return PGC_OK;
}

// the constrained branches: the same policies with the group's rows ordered by sort_population_con (tol: HOST array [nec + nic])
namespace
{
struct ConScope {
    double *d_tol = nullptr;
    int begin(size_t nec, size_t nic, const double *tol, cudaStream_t st)
    {
        PGC_REQUIRE(nec + nic > 0 && tol, "constrained policy: needs at least one constraint and its tolerances");
        PGC_CUDA(cudaMalloc(&d_tol, sizeof(double) * (nec + nic)));
        PGC_CUDA(cudaMemcpyAsync(d_tol, tol, sizeof(double) * (nec + nic), cudaMemcpyHostToDevice, st));
        tls_con.nec = nec, tls_con.nic = nic, tls_con.d_tol = d_tol;
        return PGC_OK;
    }
    ~ConScope()
    {
        tls_con = ConSpec{};
        if (d_tol) cudaFree(d_tol);
    }
};
} // namespace

int select_best_con_policy_device(pgc_ctx *ctx, const unsigned long long *d_ids, const double *d_x, const double *d_f, size_t n, size_t nx,
                                  size_t nec, size_t nic, const double *tol, int rate_is_frac, double rate, unsigned long long *d_ids_out,
                                  double *d_x_out, double *d_f_out, size_t *n_out, cudaStream_t st)
{
    ConScope scope;
    if (int rc = scope.begin(nec, nic, tol, st)) return rc;
    const int rc = select_best_policy_device(ctx, d_ids, d_x, d_f, n, nx, 1 + nec + nic, rate_is_frac, rate, d_ids_out, d_x_out, d_f_out, n_out, st);
    cudaStreamSynchronize(st); // the tolerances are freed with the scope
    return rc;
}

int fair_replace_con_policy_device(pgc_ctx *ctx, unsigned long long *d_ids, double *d_x, double *d_f, size_t n, size_t nx, size_t nec, size_t nic,
                                   const double *tol, int rate_is_frac, double rate, const unsigned long long *d_mids, const double *d_mx,
                                   const double *d_mf, size_t nm, cudaStream_t st)
{
    ConScope scope;
    if (int rc = scope.begin(nec, nic, tol, st)) return rc;
    const int rc = fair_replace_policy_device(ctx, d_ids, d_x, d_f, n, nx, 1 + nec + nic, rate_is_frac, rate, d_mids, d_mx, d_mf, nm, st);
    cudaStreamSynchronize(st);
    return rc;
}

int sort_population_con_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t nec, size_t nic, const double *tol, unsigned *d_order,
                               cudaStream_t st)
{
    if (n == 0) return PGC_OK;
    ConScope scope;
    if (int rc = scope.begin(nec, nic, tol, st)) return rc;
    Tmp tmp(st);
    const int rc = con_best_indices(ctx, d_f, n, tls_con, n, d_order, tmp, st);
    cudaStreamSynchronize(st);
    return rc;
}

// in-edge sources of vertex i of a ring built by n push_back() calls, in base_bgl_topology::get_connections order (the in-edge
// list of the reference's vecS/bidirectionalS graph keeps insertion order; ring.cpp:83-110 is replayed on plain vectors)
int ring_connections(size_t n, size_t i, std::vector<size_t> &out)
{
    PGC_REQUIRE(i < n, "invalid vertex index in a BGL topology: the index is %zu, but the number of vertices is only %zu", i, n);
    std::vector<std::vector<size_t>> in(n);
    auto add = [&](size_t u, size_t v) { in[v].push_back(u); };
    auto del = [&](size_t u, size_t v) {
        for (size_t q = 0; q < in[v].size(); ++q)
            if (in[v][q] == u) {
                in[v].erase(in[v].begin() + static_cast<long>(q));
                break;
            }
    };
    for (size_t size = 1; size <= n; ++size) {
        if (size == 2) {
            add(0, 1), add(1, 0);
        } else if (size == 3) {
            add(1, 2), add(2, 1), add(2, 0), add(0, 2);
        } else if (size > 3) {
            del(size - 2, 0), del(0, size - 2);
            add(size - 2, size - 1), add(size - 1, size - 2), add(0, size - 1), add(size - 1, 0);
        }
    }
    out = in[i];
    return PGC_OK;
}

} // namespace pgc
