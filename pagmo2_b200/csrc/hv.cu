// hv.cu - exact hypervolume indicator and exclusive contributions for 2 and 3 objectives on sm_100a.
//
// Replaces reference hypervolume::compute / contributions (src/utils/hypervolume.cpp:196-330) with the algorithms the reference
// selects for these dimensions: hv2d (hv_hv2d.cpp:59-84,133-148) and hv3d / HyCon3D (hv_hv3d.cpp:107-166,170-343).  Both reference
// algorithms are sequential sweeps over a balanced tree; the device formulation gives every point p its own sweep instead:
//   E_p(z) = area of p's quadrant [p.x, r.x) x [p.y, r.y) NOT covered by the points q != p with q.z <= z
//   contribution(p) = integral of E_p(z) dz over [p.z, r.z)         (exclusive volume, what HyCon3D accumulates box by box)
//   HV(S)           = sum_p E_p^{before}(p.z) * (r.z - p.z)          (E^{before}: only the points sorted before p count)
// For one p the state of the sweep is tiny: every other point either covers the quadrant (E = 0 from there on), clips it from
// the left / from below (two running minima), or lies strictly inside it - only those go into a 2D staircase.  Points are
// visited in ascending z (one radix sort for everybody), one thread per p, all threads reading the same q (broadcast loads).
// E is always assembled as a sum of POSITIVE column areas (never box minus covered), so small contributions keep full relative
// accuracy like the reference's box sums.  2 objectives = the same code with z = 0 and r.z = 1 (hv2d::contributions does the
// same with hv3d, hv_hv2d.cpp:135-147).
// Work is O(n^2) point visits + staircase updates - compare/min throughput bound, no atomics; bytes = 8*n*m in, 8*n out.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>

#include <vector>

#include "pgc_internal.cuh"

namespace pgc
{

namespace
{

struct HvParams {
    const double *f;       // [n x m]
    const unsigned *order; // indices by ascending z (m == 3) or nullptr (identity)
    unsigned n, m;
    double rx, ry, rz;
    double *out;           // per point: contribution, or its term of HV(S)
    double *sx, *sy;       // staircase scratch: [threads in this launch][cap]
    unsigned cap;
    unsigned p0, pcount;   // points handled by this launch: p0 + t, or plist[t]
    const unsigned *plist;
    unsigned *overflow;    // [0] = count, [1..] = points whose staircase outgrew cap (nullptr: cap is the worst case)
    int compute;           // 1: terms of HV(S)
};

__global__ void hv_check_kernel(const double *f, unsigned n, unsigned m, double rx, double ry, double rz, int *bad)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double r[3] = {rx, ry, rz};
    bool outside = false, all_equal = true; // hv_algorithm::assert_minimisation, hv_algorithm.cpp:226-258
    for (unsigned d = 0; d < m; ++d) {
        const double v = f[static_cast<size_t>(i) * m + d];
        outside |= (r[d] < v) || (v != v);
        all_equal &= (r[d] == v);
    }
    if (outside || all_equal) atomicExch(bad, 1);
}

__global__ void hv_zkeys_kernel(const double *f, unsigned n, unsigned long long *keys, unsigned *idx)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(f[static_cast<size_t>(i) * 3 + 2]));
    keys[i] = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
    idx[i] = i;
}

struct Stair { // staircase of the points strictly inside p's quadrant: x ascending, y strictly descending
    double *x, *y;
    unsigned s;
    // returns false when the staircase would outgrow `cap` (the caller reruns this point with a full-size scratch)
    __device__ bool insert(double qx, double qy, unsigned cap)
    {
        unsigned lo = 0, hi = s; // first i with x[i] >= qx
        while (lo < hi) {
            const unsigned mid = (lo + hi) >> 1;
            if (x[mid] < qx) lo = mid + 1;
            else hi = mid;
        }
        const unsigned i = lo;
        if (i > 0 && y[i - 1] <= qy) return true;          // an earlier point with smaller x is at least as low
        if (i < s && x[i] == qx && y[i] <= qy) return true; // same x, not lower
        unsigned j = i;
        while (j < s && y[j] >= qy) ++j;                    // points q now covers: x >= qx and y >= qy
        if (j == i) {                                       // make room
            if (s == cap) return false;
            for (unsigned k = s; k > i; --k) {
                x[k] = x[k - 1];
                y[k] = y[k - 1];
            }
            ++s;
        } else if (j > i + 1) {
            const unsigned gone = j - i - 1;
            for (unsigned k = j; k < s; ++k) {
                x[k - gone] = x[k];
                y[k - gone] = y[k];
            }
            s -= gone;
        }
        x[i] = qx;
        y[i] = qy;
        return true;
    }
    // uncovered area of [px, xr) x [py, yt) under the staircase, as a sum of positive columns
    __device__ double uncovered(double px, double py, double xr, double yt) const
    {
        double area = 0.0, h = yt, prevx = px;
        for (unsigned i = 0; i < s; ++i) {
            if (y[i] >= yt) continue; // above the clip: these come first
            if (x[i] >= xr) break;
            area += (x[i] - prevx) * (h - py);
            h = y[i];
            prevx = x[i];
        }
        return area + (xr - prevx) * (h - py);
    }
};

__global__ void hv_sweep_kernel(const HvParams P)
{
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.pcount) return;
    const unsigned p = P.plist ? P.plist[t] : P.p0 + t;
    const bool three = P.m == 3;
    const double *fp = P.f + static_cast<size_t>(p) * P.m;
    const double px = fp[0], py = fp[1], pz = three ? fp[2] : 0.0;
    Stair st{P.sx + static_cast<size_t>(t) * P.cap, P.sy + static_cast<size_t>(t) * P.cap, 0};
    double xr = P.rx, yt = P.ry; // the quadrant after clipping from below / from the left
    double V = 0.0, cur_z = pz, E = 0.0;
    bool dirty = true, covered = false;
    for (unsigned k = 0; k < P.n; ++k) {
        const unsigned q = P.order ? P.order[k] : k;
        if (q == p) {
            if (P.compute) break; // only the points sorted before p
            continue;
        }
        const double *fq = P.f + static_cast<size_t>(q) * P.m;
        const double qx = fq[0], qy = fq[1], qz = three ? fq[2] : 0.0;
        if (!P.compute && qz > cur_z) { // a new slab starts: integrate the finished one
            if (dirty) {
                E = st.uncovered(px, py, xr, yt);
                dirty = false;
            }
            V += E * (qz - cur_z);
            cur_z = qz;
        }
        if (qx <= px) {
            if (qy <= py) { // q covers the whole quadrant from max(q.z, p.z) upwards
                covered = true;
                break;
            }
            if (qy < yt) {
                yt = qy;
                dirty = true;
            }
        } else if (qy <= py) {
            if (qx < xr) {
                xr = qx;
                dirty = true;
            }
        } else if (qx < xr && qy < yt) {
            if (!st.insert(qx, qy, P.cap)) {
                P.overflow[1 + atomicAdd(P.overflow, 1u)] = p;
                return;
            }
            dirty = true;
        }
    }
    if (P.compute) {
        P.out[p] = covered ? 0.0 : st.uncovered(px, py, xr, yt) * (P.rz - pz);
        return;
    }
    if (!covered) {
        if (dirty) E = st.uncovered(px, py, xr, yt);
        V += E * (P.rz - cur_z);
    }
    P.out[p] = V;
}

struct Scratch {
    cudaStream_t st;
    std::vector<void *> ptrs;
    explicit Scratch(cudaStream_t s) : st(s) {}
    template <class T> int get(T **out, size_t count)
    {
        void *p = nullptr;
        PGC_CUDA(cudaMallocAsync(&p, (count ? count : 1) * sizeof(T), st));
        ptrs.push_back(p);
        *out = static_cast<T *>(p);
        return PGC_OK;
    }
    ~Scratch()
    {
        for (void *p : ptrs) cudaFreeAsync(p, st);
    }
};

} // namespace

// mode 0: d_out[n] = exclusive contributions; mode 1: d_out[0] = hypervolume (d_out needs n doubles of space)
int hv_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t m, const double *r, int compute, double *d_out, cudaStream_t st)
{
    PGC_REQUIRE(m == 2 || m == 3, "hypervolume: the device path implements hv2d and hv3d (2 or 3 objectives), %zu requested", m);
    PGC_REQUIRE(n < 0x7fffffffull, "hypervolume: too many points");
    if (n == 0) {
        if (compute) PGC_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), st));
        return PGC_OK;
    }
    Scratch tmp(st);
    const unsigned un = static_cast<unsigned>(n);
    const double rz = m == 3 ? r[2] : 1.0;
    int rc, *d_bad = nullptr;
    if ((rc = tmp.get(&d_bad, 1))) return rc;
    PGC_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    hv_check_kernel<<<(un + 255) / 256, 256, 0, st>>>(d_f, un, static_cast<unsigned>(m), r[0], r[1], m == 3 ? r[2] : 0.0, d_bad);
    PGC_CUDA(cudaGetLastError());
    int bad = 0;
    PGC_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    PGC_REQUIRE(!bad, "Reference point is invalid: another point seems to be outside the reference point boundary, or be equal to it");
    unsigned *order = nullptr;
    if (m == 3) {
        unsigned long long *k0 = nullptr, *k1 = nullptr;
        unsigned *i0 = nullptr;
        if ((rc = tmp.get(&k0, n)) || (rc = tmp.get(&k1, n)) || (rc = tmp.get(&i0, n)) || (rc = tmp.get(&order, n))) return rc;
        hv_zkeys_kernel<<<(un + 255) / 256, 256, 0, st>>>(d_f, un, k0, i0);
        PGC_CUDA(cudaGetLastError());
        size_t bytes = 0;
        PGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0, k1, i0, order, static_cast<int>(n), 0, 64, st));
        unsigned char *ws = nullptr;
        if ((rc = tmp.get(&ws, bytes))) return rc;
        PGC_CUDA(cub::DeviceRadixSort::SortPairs(ws, bytes, k0, k1, i0, order, static_cast<int>(n), 0, 64, st));
        ctx->launches.fetch_add(2, std::memory_order_relaxed);
    }
    // Staircase scratch.  First pass: every point at once with a small per-point capacity; the (rare) points whose staircase
    // outgrows it are collected and rerun with the worst-case capacity n, in batches bounded to ~1 GiB of scratch.
    const unsigned cap1 = un < 512u ? un : 512u;
    double *sx = nullptr, *sy = nullptr, *terms = nullptr;
    unsigned *ovf = nullptr;
    if ((rc = tmp.get(&sx, n * cap1)) || (rc = tmp.get(&sy, n * cap1)) || (rc = tmp.get(&ovf, n + 1))) return rc;
    if (compute && (rc = tmp.get(&terms, n))) return rc;
    PGC_CUDA(cudaMemsetAsync(ovf, 0, sizeof(unsigned), st));
    HvParams P;
    P.f = d_f;
    P.order = order;
    P.n = un;
    P.m = static_cast<unsigned>(m);
    P.rx = r[0];
    P.ry = r[1];
    P.rz = rz;
    P.out = compute ? terms : d_out;
    P.sx = sx;
    P.sy = sy;
    P.cap = cap1;
    P.p0 = 0;
    P.pcount = un;
    P.plist = nullptr;
    P.overflow = ovf;
    P.compute = compute;
    hv_sweep_kernel<<<(un + 63) / 64, 64, 0, st>>>(P);
    PGC_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    if (cap1 < un) {
        unsigned nover = 0;
        PGC_CUDA(cudaMemcpyAsync(&nover, ovf, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        if (nover) {
            size_t batch = (size_t(1) << 30) / (2 * sizeof(double) * n);
            if (batch < 64) batch = 64;
            if (batch > nover) batch = nover;
            double *bx = nullptr, *by = nullptr;
            if ((rc = tmp.get(&bx, batch * n)) || (rc = tmp.get(&by, batch * n))) return rc;
            P.sx = bx;
            P.sy = by;
            P.cap = un;
            P.overflow = nullptr; // cannot overflow any more
            for (size_t o0 = 0; o0 < nover; o0 += batch) {
                P.plist = ovf + 1 + o0;
                P.pcount = static_cast<unsigned>(nover - o0 < batch ? nover - o0 : batch);
                hv_sweep_kernel<<<(P.pcount + 63) / 64, 64, 0, st>>>(P);
                PGC_CUDA(cudaGetLastError());
                ctx->launches.fetch_add(1, std::memory_order_relaxed);
            }
        }
    }
    if (compute) {
        size_t bytes = 0;
        PGC_CUDA(cub::DeviceReduce::Sum(nullptr, bytes, terms, d_out, static_cast<int>(n), st));
        unsigned char *ws = nullptr;
        if ((rc = tmp.get(&ws, bytes))) return rc;
        PGC_CUDA(cub::DeviceReduce::Sum(ws, bytes, terms, d_out, static_cast<int>(n), st));
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
    }
    return PGC_OK;
}

} // namespace pgc
