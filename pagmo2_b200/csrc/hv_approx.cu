// hv_approx.cu - the Bringmann-Friedrich approximations of the hypervolume on the device:
//   bf_fpras  (reference src/utils/hv_algos/hv_bf_fpras.cpp:91-146): (eps, delta) approximation of the hypervolume of a union of boxes
//             by the Karp-Luby estimator - pick a box with probability proportional to its volume, a point in it, then draw points
//             of the front until one dominates the sample; T * V / (n * M) with T the trials made and M the completed rounds.
//   bf_approx (hv_bf_approx.cpp:131-470): least / greatest contributor by rounds of Monte-Carlo sampling inside every point's
//             bounding box, with confidence radii that eliminate candidates round by round (and the reference's switch to the exact
//             exclusive volume for small / expensive boxes).
// Both are sequential loops over ONE random engine in the reference.  Here the samples are independent Philox substreams: fpras
// splits the trial budget over threads that each run whole rounds; approx keeps the reference's round structure on the host (a few
// numbers per point) and draws all outstanding samples of a round on the device, one CTA per point.  The outputs are the same
// estimators with the same guarantees, not the same numbers: parity is statistical (tests/test_gpu_hv_approx.py).
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

#include "pgc_internal.cuh"
#include "philox.cuh"

namespace pgc
{
namespace
{

constexpr unsigned kMaxDim = 12;

struct FprasParams {
    const double *pts;  // [n x m]
    const double *sums; // [n] partial sums of the box volumes
    unsigned n, m;
    double r[kMaxDim];
    double V;
    unsigned long long budget; // trials per thread
    unsigned long long seed;
    unsigned nthreads;
    unsigned long long *rounds; // [2] completed rounds and the trials they took, over all threads
};

__global__ void fpras_kernel(const FprasParams P)
{
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.nthreads) return;
    PhiloxStream rs(P.seed, kTagHvApprox, 0u, t);
    unsigned long long trials = 0, M = 0;
    double pt[kMaxDim];
    for (;;) {
        // the box, with probability sums[i] / V (std::lower_bound, :117-121)
        const double u = rs.next() * P.V;
        unsigned lo = 0, hi = P.n;
        while (lo < hi) {
            const unsigned mid = (lo + hi) >> 1;
            if (P.sums[mid] < u) lo = mid + 1;
            else hi = mid;
        }
        const unsigned i = lo < P.n ? lo : P.n - 1u;
        for (unsigned c = 0; c < P.m; ++c) { // a point inside it, :124-126
            const double a = P.pts[static_cast<size_t>(i) * P.m + c];
            pt[c] = a + rs.next() * (P.r[c] - a);
        }
        for (;;) { // draw points of the front until one dominates the sample, :129-137
            unsigned j = static_cast<unsigned>(static_cast<double>(P.n) * rs.next());
            if (j >= P.n) j = P.n - 1u;
            ++trials;
            const double *q = P.pts + static_cast<size_t>(j) * P.m;
            bool le = true, lt = false; // hv_algorithm::dom_cmp(sample, q) == B_DOMINATES_A: q <= sample everywhere, < somewhere
            for (unsigned c = 0; c < P.m; ++c) {
                le = le && q[c] <= pt[c];
                lt = lt || q[c] < pt[c];
            }
            if (le && lt) break;
        }
        ++M;
        // The reference stops in the middle of the round in which the budget runs out and divides the WHOLE budget by the completed
        // rounds; with thousands of threads each doing that, the dropped rounds (long ones, preferentially) would bias the ratio.
        // A thread finishes the round it is in and reports the trials it really made: trials / rounds is then a ratio of sums of
        // whole rounds (Wald), the same estimator without the truncation.
        if (trials >= P.budget) break;
    }
    atomicAdd(P.rounds, M);
    atomicAdd(P.rounds + 1, trials);
}

struct ApproxParams {
    const double *pts;      // [n x m]
    const double *boxes;    // [n x m] opposite corners of the bounding boxes
    const unsigned *bp_off; // [n + 1] CSR of the points inside every box
    const unsigned *bp_idx;
    unsigned m;
    const unsigned *who;             // [count] the points sampled in this launch
    const unsigned long long *first; // [count] samples each of them has already had (the substream position)
    const unsigned long long *todo;  // [count] samples to draw now
    unsigned long long *succ, *ops;  // [count] out: successful samples, dominance operations (hv_bf_approx.cpp:311)
    unsigned long long seed;
};

// sample_successful (:292-320), `todo` times for the point of this CTA
__global__ void approx_sample_kernel(const ApproxParams P)
{
    __shared__ unsigned long long s_succ[256], s_ops[256];
    const unsigned b = blockIdx.x, idx = P.who[b], m = P.m;
    const double *lb = P.pts + static_cast<size_t>(idx) * m, *ub = P.boxes + static_cast<size_t>(idx) * m;
    const unsigned k0 = P.bp_off[idx], k1 = P.bp_off[idx + 1u];
    unsigned long long succ = 0, ops = 0;
    double pt[kMaxDim];
    for (unsigned long long s = threadIdx.x; s < P.todo[b]; s += blockDim.x) {
        const unsigned long long g = P.first[b] + s; // global sample number of this point: a fixed window of its substream
        PhiloxStream rk(P.seed, kTagHvApprox, static_cast<unsigned>(g >> 20) + 1u, idx);
        rk.slot = static_cast<uint32_t>(g & 0xfffffull) * 16u; // <= kMaxDim draws per sample
        for (unsigned c = 0; c < m; ++c) pt[c] = lb[c] + rk.next() * (ub[c] - lb[c]);
        bool ok = true;
        for (unsigned k = k0; k < k1 && ok; ++k) {
            const double *q = P.pts + static_cast<size_t>(P.bp_idx[k]) * m;
            ops += m + 1u;
            bool dominates = true;
            for (unsigned c = 0; c < m; ++c)
                if (pt[c] < q[c]) {
                    dominates = false;
                    break;
                }
            ok = !dominates;
        }
        succ += ok ? 1u : 0u;
    }
    s_succ[threadIdx.x] = succ, s_ops[threadIdx.x] = ops;
    __syncthreads();
    for (unsigned w = blockDim.x / 2; w; w >>= 1) {
        if (threadIdx.x < w) {
            s_succ[threadIdx.x] += s_succ[threadIdx.x + w];
            s_ops[threadIdx.x] += s_ops[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        P.succ[b] = s_succ[0];
        P.ops[b] = s_ops[0];
    }
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf()
    {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes)
    {
        PGC_CUDA(cudaMalloc(&p, bytes ? bytes : 8));
        return PGC_OK;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

int check_points(const double *points, size_t n, size_t m, const double *r) // hv_algorithm::assert_minimisation, hv_algorithm.cpp:226-258
{
    for (size_t i = 0; i < n; ++i) {
        bool outside = false, all_equal = true;
        for (size_t c = 0; c < m; ++c) {
            outside = outside || r[c] < points[i * m + c];
            all_equal = all_equal && r[c] == points[i * m + c];
        }
        PGC_REQUIRE(!outside && !all_equal,
                    "Reference point is invalid: another point seems to be outside the reference point boundary, or be equal to it");
    }
    return PGC_OK;
}

double expected_hv_operations(size_t n, size_t d) // detail::expected_hv_operations, hypervolume.cpp:465-475
{
    if (d <= 3u) return static_cast<double>(d) * static_cast<double>(n) * std::log(static_cast<double>(n));
    if (d == 4u) return 4.0 * static_cast<double>(n) * static_cast<double>(n);
    return 0.0005 * static_cast<double>(d) * std::pow(static_cast<double>(n), static_cast<double>(d) * 0.5);
}

} // namespace

int hv_fpras_host(pgc_ctx *ctx, const double *points, size_t n, size_t m, const double *r, double eps, double delta, unsigned long long seed,
                  double *hv_out)
{
    PGC_REQUIRE(eps > 0. && eps <= 1., "Epsilon needs to be a probability greater then zero"); // hv_bf_fpras.cpp:54-59
    PGC_REQUIRE(delta > 0. && delta <= 1., "Delta needs to be a probability greater than zero");
    PGC_REQUIRE(m >= 2 && m <= kMaxDim, "bf_fpras on the device: between 2 and %u objectives, %zu requested", kMaxDim, m);
    PGC_REQUIRE(n < 0x7fffffffull, "bf_fpras: too many points");
    if (int rc = check_points(points, n, m, r)) return rc;
    if (n == 0) {
        *hv_out = 0.;
        return PGC_OK;
    }
    cudaStream_t st = ctx->stream;
    const double T = std::floor(12. * std::log(1. / delta) / std::log(2.) * static_cast<double>(n) / eps / eps); // :96
    std::vector<double> sums(n);
    double V = 0.;
    for (size_t i = 0; i < n; ++i) { // :112-114
        double v = 1.;
        for (size_t c = 0; c < m; ++c) v *= (points[i * m + c] - r[c]);
        V = (sums[i] = V + std::fabs(v));
    }
    if (!(T >= 1.)) { // delta = 1: no trial is made and the reference divides by M = 0
        *hv_out = std::numeric_limits<double>::quiet_NaN();
        return PGC_OK;
    }
    FprasParams P{};
    P.n = static_cast<unsigned>(n), P.m = static_cast<unsigned>(m), P.V = V, P.seed = seed;
    for (size_t c = 0; c < m; ++c) P.r[c] = r[c];
    // whole rounds per thread: enough threads to fill the device, few enough that a thread's unfinished last round is noise
    const double want = std::min(T / 4096., static_cast<double>(ctx->sm_count ? ctx->sm_count : 148) * 2048.);
    P.nthreads = static_cast<unsigned>(std::max(1., std::floor(want)));
    P.budget = static_cast<unsigned long long>(std::ceil(T / P.nthreads));
    PGC_REQUIRE(static_cast<double>(P.budget) * static_cast<double>(m + 3) < 4.0e9, "bf_fpras: eps / delta ask for more draws per thread than a substream holds");
    DevBuf d_pts, d_sums, d_rounds;
    int rc;
    if ((rc = d_pts.alloc(sizeof(double) * n * m)) || (rc = d_sums.alloc(sizeof(double) * n)) || (rc = d_rounds.alloc(2 * sizeof(unsigned long long)))) return rc;
    PGC_CUDA(cudaMemcpyAsync(d_pts.p, points, sizeof(double) * n * m, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(d_sums.p, sums.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemsetAsync(d_rounds.p, 0, 2 * sizeof(unsigned long long), st));
    P.pts = d_pts.as<double>(), P.sums = d_sums.as<double>(), P.rounds = d_rounds.as<unsigned long long>();
    fpras_kernel<<<(P.nthreads + 127) / 128, 128, 0, st>>>(P);
    PGC_CUDA(cudaGetLastError());
    unsigned long long MT[2] = {0, 0};
    PGC_CUDA(cudaMemcpyAsync(MT, d_rounds.p, sizeof(MT), cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    *hv_out = (static_cast<double>(MT[1]) * V) / (static_cast<double>(n) * static_cast<double>(MT[0])); // T * V / (n * M), :131
    return PGC_OK;
}

int hv_approx_extreme_host(pgc_ctx *ctx, const double *points, size_t n, size_t m, const double *r, int greatest, int use_exact,
                           unsigned trivial_subcase_size, double eps, double delta, double delta_multiplier, double alpha,
                           double initial_delta_coeff, double gamma, unsigned long long seed, size_t *idx_out)
{
    PGC_REQUIRE(eps >= 0. && eps <= 1., "Epsilon needs to be a probability."); // hv_bf_approx.cpp:60-65
    PGC_REQUIRE(delta >= 0. && delta <= 1., "Delta needs to be a probability.");
    PGC_REQUIRE(m >= 2 && m <= kMaxDim, "bf_approx on the device: between 2 and %u objectives, %zu requested", kMaxDim, m);
    PGC_REQUIRE(n >= 1 && n < 0x7fffffffull, "bf_approx: the number of points must be in [1, 2^31)");
    if (int rc = check_points(points, n, m, r)) return rc;
    cudaStream_t st = ctx->stream;
    const auto X = [&](size_t i, size_t c) { return points[i * m + c]; };
    // strict Pareto dominance of row a over row b (hv_algorithm::dom_cmp == A_DOMINATES_B)
    const auto dominates = [&](const double *a, const double *b) {
        bool le = true, lt = false;
        for (size_t c = 0; c < m; ++c) {
            le = le && a[c] <= b[c];
            lt = lt || a[c] < b[c];
        }
        return le && lt;
    };
    // bounding boxes and the points that reach into them, :365-392 (compute_bounding_box :148-177, point_in_box :188-200)
    std::vector<double> boxes(n * m), box_volume(n), approx_volume(n, 0.), point_delta(n, 0.);
    std::vector<unsigned long long> no_samples(n, 0), no_succ(n, 0), no_ops(n, 1);
    std::vector<std::vector<unsigned>> box_points(n);
    double r_delta = 0.;
    for (size_t idx = 0; idx < n; ++idx) {
        double *z = boxes.data() + idx * m;
        std::copy(r, r + m, z);
        for (size_t j = 0; j < n; ++j) { // a point worse than p in exactly one objective bounds the box there
            bool flag = false;
            size_t worse = 0;
            for (size_t c = 0; c < m; ++c)
                if (X(j, c) >= X(idx, c)) {
                    if (flag) {
                        flag = false;
                        break;
                    }
                    worse = c;
                    flag = true;
                }
            if (flag) z[worse] = std::min(z[worse], X(j, worse));
        }
        double v = 1.;
        for (size_t c = 0; c < m; ++c) v *= (X(idx, c) - z[c]);
        box_volume[idx] = std::fabs(v);
        r_delta = std::max(r_delta, box_volume[idx]);
        const double *a = points + idx * m;
        for (size_t j = 0; j < n; ++j) {
            if (j == idx) continue;
            const double *p = points + j * m;
            const bool equal = std::equal(a, a + m, p);
            if (equal || dominates(p, a)) { // point_in_box 3 / 2: idx contributes nothing, :383-390
                if (!greatest) {
                    *idx_out = idx;
                    return PGC_OK;
                }
            } else if (dominates(p, z)) { // point_in_box 1
                box_points[idx].push_back(static_cast<unsigned>(j));
            }
        }
    }
    // device copies: points, boxes, the CSR of the box points
    std::vector<unsigned> bp_off(n + 1, 0), bp_idx;
    for (size_t i = 0; i < n; ++i) {
        bp_off[i + 1] = bp_off[i] + static_cast<unsigned>(box_points[i].size());
        bp_idx.insert(bp_idx.end(), box_points[i].begin(), box_points[i].end());
    }
    DevBuf d_pts, d_boxes, d_off, d_idx, d_who, d_first, d_todo, d_succ, d_ops, d_sub, d_hv;
    int rc;
    if ((rc = d_pts.alloc(sizeof(double) * n * m)) || (rc = d_boxes.alloc(sizeof(double) * n * m)) || (rc = d_off.alloc(4 * (n + 1)))
        || (rc = d_idx.alloc(4 * bp_idx.size())) || (rc = d_who.alloc(4 * n)) || (rc = d_first.alloc(8 * n)) || (rc = d_todo.alloc(8 * n))
        || (rc = d_succ.alloc(8 * n)) || (rc = d_ops.alloc(8 * n)) || (rc = d_sub.alloc(sizeof(double) * n * m)) || (rc = d_hv.alloc(sizeof(double) * n)))
        return rc;
    PGC_CUDA(cudaMemcpyAsync(d_pts.p, points, sizeof(double) * n * m, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(d_boxes.p, boxes.data(), sizeof(double) * n * m, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(d_off.p, bp_off.data(), 4 * (n + 1), cudaMemcpyHostToDevice, st));
    if (!bp_idx.empty()) PGC_CUDA(cudaMemcpyAsync(d_idx.p, bp_idx.data(), 4 * bp_idx.size(), cudaMemcpyHostToDevice, st));
    ApproxParams AP{};
    AP.pts = d_pts.as<double>(), AP.boxes = d_boxes.as<double>(), AP.bp_off = d_off.as<unsigned>(), AP.bp_idx = d_idx.as<unsigned>();
    AP.m = static_cast<unsigned>(m), AP.who = d_who.as<unsigned>(), AP.first = d_first.as<unsigned long long>();
    AP.todo = d_todo.as<unsigned long long>(), AP.succ = d_succ.as<unsigned long long>(), AP.ops = d_ops.as<unsigned long long>(), AP.seed = seed;
    const double log_factor = std::log(2. * static_cast<double>(n) * (1. + gamma) / (delta * gamma)); // :357
    std::vector<unsigned> who;
    std::vector<unsigned long long> first, todo, succ, ops;
    std::vector<double> sub;
    // sampling_round (:211-262) for a batch of points: the exact sub-cases on the spot, all outstanding samples in one launch
    const auto sampling_round = [&](const std::vector<size_t> &batch, double dlt, unsigned round_no) -> int {
        who.clear(), first.clear(), todo.clear();
        const double lf = (1. + gamma) * std::log(static_cast<double>(round_no)) + log_factor;
        for (size_t idx : batch) {
            if (use_exact) {
                if (no_ops[idx] == 0) continue;
                const auto &bp = box_points[idx];
                if (bp.size() <= trivial_subcase_size || static_cast<double>(no_ops[idx]) >= expected_hv_operations(bp.size(), m)) {
                    if (bp.empty()) {
                        approx_volume[idx] = box_volume[idx];
                    } else if (bp.size() == 1u) { // one box point: what it covers of the box is a box itself
                        double v = 1.;
                        for (size_t c = 0; c < m; ++c) v *= (std::max(X(idx, c), X(bp[0], c)) - boxes[idx * m + c]);
                        approx_volume[idx] = box_volume[idx] - std::fabs(v);
                    } else { // the exclusive volume itself: the box minus what the clipped box points cover, :226-243
                        sub.resize(bp.size() * m);
                        for (size_t q = 0; q < bp.size(); ++q)
                            for (size_t c = 0; c < m; ++c) sub[q * m + c] = std::max(X(idx, c), X(bp[q], c));
                        PGC_CUDA(cudaMemcpyAsync(d_sub.p, sub.data(), sizeof(double) * sub.size(), cudaMemcpyHostToDevice, st));
                        if (int r2 = hv_device(ctx, d_sub.as<double>(), bp.size(), m, boxes.data() + idx * m, 1, d_hv.as<double>(), st)) return r2;
                        double hv = 0.;
                        PGC_CUDA(cudaMemcpyAsync(&hv, d_hv.p, sizeof(double), cudaMemcpyDeviceToHost, st));
                        PGC_CUDA(cudaStreamSynchronize(st));
                        approx_volume[idx] = box_volume[idx] - hv;
                    }
                    point_delta[idx] = 0.0;
                    no_ops[idx] = 0;
                    continue;
                }
            }
            const double tmp = box_volume[idx] / dlt, required = 0.5 * lf * tmp * tmp;
            unsigned long long need = 0;
            if (static_cast<double>(no_samples[idx]) < required) {
                PGC_REQUIRE(required < 4.0e15, "bf_approx: a round asks for %g samples of one point", required);
                need = static_cast<unsigned long long>(std::ceil(required)) - no_samples[idx];
            }
            who.push_back(static_cast<unsigned>(idx)), first.push_back(no_samples[idx]), todo.push_back(need);
        }
        if (who.empty()) return PGC_OK;
        const size_t cnt = who.size();
        PGC_CUDA(cudaMemcpyAsync(d_who.p, who.data(), 4 * cnt, cudaMemcpyHostToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(d_first.p, first.data(), 8 * cnt, cudaMemcpyHostToDevice, st));
        PGC_CUDA(cudaMemcpyAsync(d_todo.p, todo.data(), 8 * cnt, cudaMemcpyHostToDevice, st));
        approx_sample_kernel<<<static_cast<unsigned>(cnt), 256, 0, st>>>(AP);
        PGC_CUDA(cudaGetLastError());
        succ.resize(cnt), ops.resize(cnt);
        PGC_CUDA(cudaMemcpyAsync(succ.data(), d_succ.p, 8 * cnt, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaMemcpyAsync(ops.data(), d_ops.p, 8 * cnt, cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        for (size_t k = 0; k < cnt; ++k) {
            const size_t idx = who[k];
            no_samples[idx] += todo[k];
            no_succ[idx] += succ[k];
            no_ops[idx] += ops[k];
            approx_volume[idx] = static_cast<double>(no_succ[idx]) / static_cast<double>(no_samples[idx]) * box_volume[idx];
            point_delta[idx] = std::sqrt(0.5 * lf / static_cast<double>(no_samples[idx])) * box_volume[idx]; // compute_point_delta, :136-140
        }
        return PGC_OK;
    };
    std::vector<size_t> point_set(n);
    for (size_t i = 0; i < n; ++i) point_set[i] = i;
    const auto better = [&](double a, double b) { return greatest ? a > b : a < b; };
    unsigned round_no = 0;
    size_t EC = 0;
    bool stop = false;
    r_delta *= initial_delta_coeff;
    do { // :397-458
        r_delta *= delta_multiplier;
        ++round_no;
        if ((rc = sampling_round(point_set, r_delta, round_no))) return rc;
        if ((rc = sampling_round(std::vector<size_t>{EC}, alpha * r_delta, round_no))) return rc;
        for (size_t idx : point_set)
            if (better(approx_volume[idx], approx_volume[EC])) EC = idx;
        std::vector<size_t> kept;
        for (size_t idx : point_set) {
            const bool erase = greatest ? (approx_volume[idx] + point_delta[idx]) < (approx_volume[EC] - point_delta[EC])  // gc_erase_condition
                                        : (approx_volume[idx] - point_delta[idx]) > (approx_volume[EC] + point_delta[EC]); // lc_erase_condition
            if (idx == EC || !erase) kept.push_back(idx);
        }
        point_set.swap(kept);
        stop = true;
        if (point_set.size() > 1)
            for (size_t idx : point_set) {
                if (idx == EC) continue;
                const double d = greatest ? (approx_volume[idx] + point_delta[idx]) / (approx_volume[EC] - point_delta[EC])
                                          : (approx_volume[EC] + point_delta[EC]) / (approx_volume[idx] - point_delta[idx]);
                if (d <= 0 || d > 1 + eps) {
                    stop = false;
                    break;
                }
            }
        PGC_REQUIRE(round_no < 100000u, "bf_approx: no decision after %u rounds", round_no);
    } while (!stop);
    *idx_out = EC;
    return PGC_OK;
}

} // namespace pgc
