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
#include <cub/device/device_scan.cuh>

#include <algorithm>
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


// ---- first pass: one WARP per point ------------------------------------------------------------------------------------------
// The 32 lanes look at 32 consecutive points of the z-order at once.  Most of them change nothing (a left / below point only
// matters when it beats the running minimum, an interior point only when the staircase does not already cover it - each lane
// checks that on its own, with a binary search in the warp's shared-memory staircase), so the lanes vote and the warp steps
// only through the few EVENTS of the chunk, in z order: integrate the slab since the last event, apply the event, and
// recompute the uncovered area co-operatively (lanes strided over the staircase).  The staircase keeps only its live range
// [lo, hi): a smaller yt retires entries from the front, a smaller xr from the back, without moving data.
constexpr int kHvWarps = 8, kHvCap = 512;

struct HvWarpParams {
    const double *pts;     // [n][3] sorted by z (2 objectives: z = 0, input order)
    const double *ptsx;    // [n][4] the same points by ascending (x, y): x, y, z, position in the z-order
    const unsigned *order; // sorted position -> original index (nullptr: identity)
    unsigned n;
    double rx, ry, rz;
    double *out;
    unsigned *overflow;    // [0] = count, [1..] = ORIGINAL indices to rerun with the full-size scratch
    int compute;
};

struct WarpStair {
    double *x, *y; // shared memory, capacity kHvCap
    int lo, hi;    // live range (uniform across the warp)

    // is (qx, qy) covered by a live staircase point?  (called per lane, independently)
    __device__ bool covers(double qx, double qy) const
    {
        int a = lo, b = hi; // first i in [lo, hi) with x[i] >= qx
        while (a < b) {
            const int mid = (a + b) >> 1;
            if (x[mid] < qx) a = mid + 1;
            else b = mid;
        }
        if (a > lo && y[a - 1] <= qy) return true;
        return a < hi && x[a] == qx && y[a] <= qy;
    }
    // warp-uniform insert; returns 0 = already covered, 1 = inserted, -1 = out of capacity
    __device__ int insert(double qx, double qy, int lane)
    {
        int a = lo, b = hi;
        while (a < b) {
            const int mid = (a + b) >> 1;
            if (x[mid] < qx) a = mid + 1;
            else b = mid;
        }
        const int i = a;
        if (i > lo && y[i - 1] <= qy) return 0;
        if (i < hi && x[i] == qx && y[i] <= qy) return 0;
        a = i, b = hi; // first j >= i with y[j] < qy (y descends): [i, j) are the entries q covers
        while (a < b) {
            const int mid = (a + b) >> 1;
            if (y[mid] >= qy) a = mid + 1;
            else b = mid;
        }
        const int j = a;
        if (j == i) { // grow by one: shift [i, hi) right
            if (hi == kHvCap) {
                if (lo == 0) return -1;
                for (int base = lo; base < hi; base += 32) { // compact to the front first (reads run ahead of the writes)
                    const int k = base + lane;
                    double vx = 0, vy = 0;
                    if (k < hi) {
                        vx = x[k];
                        vy = y[k];
                    }
                    __syncwarp();
                    if (k < hi) {
                        x[k - lo] = vx;
                        y[k - lo] = vy;
                    }
                    __syncwarp();
                }
                const int shift = lo;
                hi -= shift;
                lo = 0;
                return insert_at(i - shift, i - shift, qx, qy, lane);
            }
            return insert_at(i, j, qx, qy, lane);
        }
        return insert_at(i, j, qx, qy, lane);
    }
    __device__ int insert_at(int i, int j, double qx, double qy, int lane)
    {
        if (j == i) {
            for (int top = hi; top > i; top -= 32) { // chunks from the back so nothing is overwritten before it is read
                const int k = top - 1 - lane;
                double vx = 0, vy = 0;
                if (k >= i) {
                    vx = x[k];
                    vy = y[k];
                }
                __syncwarp();
                if (k >= i) {
                    x[k + 1] = vx;
                    y[k + 1] = vy;
                }
                __syncwarp();
            }
            ++hi;
        } else if (j > i + 1) {
            const int gone = j - i - 1;
            for (int base = j; base < hi; base += 32) {
                const int k = base + lane;
                double vx = 0, vy = 0;
                if (k < hi) {
                    vx = x[k];
                    vy = y[k];
                }
                __syncwarp();
                if (k < hi) {
                    x[k - gone] = vx;
                    y[k - gone] = vy;
                }
                __syncwarp();
            }
            hi -= gone;
        }
        if (lane == 0) {
            x[i] = qx;
            y[i] = qy;
        }
        __syncwarp();
        return 1;
    }
    // retire entries outside the clip: y >= yt at the front, x >= xr at the back
    __device__ void trim(double xr, double yt)
    {
        int a = lo, b = hi;
        while (a < b) { // first i with y[i] < yt
            const int mid = (a + b) >> 1;
            if (y[mid] >= yt) a = mid + 1;
            else b = mid;
        }
        lo = a;
        a = lo, b = hi;
        while (a < b) { // first i with x[i] >= xr
            const int mid = (a + b) >> 1;
            if (x[mid] < xr) a = mid + 1;
            else b = mid;
        }
        hi = a;
    }
    // uncovered area of [px, xr) x [py, yt) under the live staircase, all lanes get the sum (positive columns only)
    __device__ double uncovered(double px, double py, double xr, double yt, int lane) const
    {
        double part = 0.0;
        for (int i = lo + lane; i < hi; i += 32) {
            const double prevx = i == lo ? px : x[i - 1], h = i == lo ? yt : y[i - 1];
            part += (x[i] - prevx) * (h - py);
        }
        if (lane == 0) {
            const double lastx = hi > lo ? x[hi - 1] : px, h = hi > lo ? y[hi - 1] : yt;
            part += (xr - lastx) * (h - py);
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
        return part;
    }
};

__global__ void __launch_bounds__(kHvWarps * 32) hv_warp_kernel(const HvWarpParams P)
{
    extern __shared__ double s_stair_raw[]; // [kHvWarps][2][kHvCap]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned pp = blockIdx.x * kHvWarps + warp; // position of this warp's point in the z-order
    if (pp >= P.n) return;
    const double px = P.pts[3 * pp], py = P.pts[3 * pp + 1], pz = P.pts[3 * pp + 2];
    WarpStair st{s_stair_raw + static_cast<size_t>(warp) * 2 * kHvCap, s_stair_raw + static_cast<size_t>(warp) * 2 * kHvCap + kHvCap, 0, 0};
    const double kInf = __longlong_as_double(0x7ff0000000000000ll);
    const unsigned kFullMask = 0xffffffffu;
    double xr = P.rx, yt = P.ry;
    bool covered = false, overflow = false;

    // ---- phase 1: everything at or below p's height (compute: everything sorted before p).  No volume is swept yet, so the
    // order is free: the points are visited by ascending (x, y), which makes the staircase APPEND-ONLY - an interior point
    // joins iff its y beats every earlier one - and the whole chunk is handled with votes and prefix minima, no serial events.
    {
        double ymin = kInf; // y of the last staircase entry
        int hi = 0;
        for (unsigned base = 0; base < P.n && !covered; base += 32) {
            const unsigned k = base + lane;
            double qx = 0, qy = 0, qz = 0;
            unsigned qpos = pp;
            if (k < P.n) {
                const double2 v0 = reinterpret_cast<const double2 *>(P.ptsx)[2 * static_cast<size_t>(k)];
                const double2 v1 = reinterpret_cast<const double2 *>(P.ptsx)[2 * static_cast<size_t>(k) + 1];
                qx = v0.x;
                qy = v0.y;
                qz = v1.x;
                qpos = static_cast<unsigned>(v1.y);
            }
            const bool in1 = qpos != pp && (P.compute ? qpos < pp : qz <= pz);
            const bool lefty = qx <= px;
            if (__any_sync(kFullMask, in1 && lefty && qy <= py)) { // p's quadrant is covered from its own height on
                covered = true;
                break;
            }
            // the reductions below are skipped for the (vast majority of) chunks that cannot change the state
            if (__any_sync(kFullMask, in1 && lefty && qy < yt)) {
                double cy = (in1 && lefty) ? qy : kInf;
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) cy = fmin(cy, __shfl_xor_sync(kFullMask, cy, m));
                yt = fmin(yt, cy);
            }
            if (__any_sync(kFullMask, in1 && !lefty && qy <= py && qx < xr)) {
                double cx = (in1 && !lefty && qy <= py) ? qx : kInf;
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) cx = fmin(cx, __shfl_xor_sync(kFullMask, cx, m));
                xr = fmin(xr, cx);
            }
            const bool inter = in1 && !lefty && qy > py && qx < xr && qy < yt && qy < ymin;
            if (!__any_sync(kFullMask, inter)) continue;
            // exclusive prefix minimum of the interior y's in lane (= x) order, seeded with the staircase's last y
            const double v = inter ? qy : kInf;
            double pm = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double o = __shfl_up_sync(kFullMask, pm, d);
                if (lane >= d) pm = fmin(pm, o);
            }
            double ex = __shfl_up_sync(kFullMask, pm, 1);
            ex = lane == 0 ? ymin : fmin(ex, ymin);
            const bool rec = inter && qy < ex;
            const unsigned m = __ballot_sync(kFullMask, rec);
            if (m) {
                const int at = hi + __popc(m & ((1u << lane) - 1u));
                if (rec && at < kHvCap) {
                    st.x[at] = qx;
                    st.y[at] = qy;
                }
                hi += __popc(m);
                if (hi > kHvCap) {
                    overflow = true;
                    break;
                }
                ymin = fmin(ymin, __shfl_sync(kFullMask, pm, 31));
            }
        }
        __syncwarp();
        st.lo = 0;
        st.hi = hi;
        if (!covered && !overflow) st.trim(xr, yt);
    }

    // ---- phase 2 (contributions only): the points above p in ascending z.  Each chunk of 32 is pre-filtered in parallel; the few
    // EVENTS that change the state are stepped through in z order: integrate the slab since the last event, apply the event.
    double V = 0.0, cur_z = pz, E = 0.0;
    bool dirty = true;
    if (!P.compute && !covered && !overflow) {
        unsigned a = pp + 1, b = P.n; // first position with z > pz (z ascending; ties with p belong to phase 1)
        while (a < b) {
            const unsigned mid = (a + b) >> 1;
            if (P.pts[3 * mid + 2] <= pz) a = mid + 1;
            else b = mid;
        }
        for (unsigned base = a; base < P.n && !covered && !overflow; base += 32) {
            const unsigned k = base + lane;
            const bool act = k < P.n;
            double qx = 0, qy = 0, qz = 0;
            if (act) {
                qx = P.pts[3 * k];
                qy = P.pts[3 * k + 1];
                qz = P.pts[3 * k + 2];
            }
            int type = 0; // 1 cover, 2 left, 3 below, 4 interior; pre-filter against the state at the start of the chunk
            if (act) {
                if (qx <= px) type = (qy <= py) ? 1 : (qy < yt ? 2 : 0);
                else if (qy <= py) type = (qx < xr) ? 3 : 0;
                else if (qx < xr && qy < yt && !st.covers(qx, qy)) type = 4;
            }
            unsigned mask = __ballot_sync(kFullMask, type != 0);
            while (mask) {
                const int l = __ffs(mask) - 1;
                mask &= mask - 1;
                const int et = __shfl_sync(kFullMask, type, l);
                const double ex = __shfl_sync(kFullMask, qx, l), ey = __shfl_sync(kFullMask, qy, l), ez = __shfl_sync(kFullMask, qz, l);
                // an earlier event of this chunk may have made this one irrelevant
                if ((et == 2 && !(ey < yt)) || (et == 3 && !(ex < xr)) || (et == 4 && !(ex < xr && ey < yt))) continue;
                if (ez > cur_z) { // a slab ends here
                    if (dirty) {
                        E = st.uncovered(px, py, xr, yt, lane);
                        dirty = false;
                    }
                    V += E * (ez - cur_z);
                    cur_z = ez;
                }
                if (et == 1) {
                    covered = true;
                    break;
                }
                if (et == 2) {
                    yt = ey;
                    st.trim(xr, yt);
                    dirty = true;
                } else if (et == 3) {
                    xr = ex;
                    st.trim(xr, yt);
                    dirty = true;
                } else {
                    const int r = st.insert(ex, ey, lane);
                    if (r < 0) {
                        overflow = true;
                        break;
                    }
                    dirty |= r > 0;
                }
            }
        }
    }
    const unsigned orig = P.order ? P.order[pp] : pp;
    if (overflow) {
        if (lane == 0) P.overflow[1 + atomicAdd(P.overflow, 1u)] = orig;
        return;
    }
    double res;
    if (P.compute) {
        res = covered ? 0.0 : st.uncovered(px, py, xr, yt, lane) * (P.rz - pz);
    } else {
        // `covered` from phase 1 means a point at or below p's height dominates it: nothing is exclusive.  In phase 2 it only ends
        // the sweep (the slab up to the covering point has been added).
        if (!covered || cur_z > pz) {
            if (!covered) {
                if (dirty) E = st.uncovered(px, py, xr, yt, lane);
                V += E * (P.rz - cur_z);
            }
        }
        res = V;
    }
    if (lane == 0) P.out[orig] = res;
}

__global__ void hv_gather_sorted_kernel(const double *f, const unsigned *order, unsigned n, unsigned m, double *pts)
{
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t q = order ? order[k] : k;
    pts[3 * k] = f[q * m];
    pts[3 * k + 1] = f[q * m + 1];
    pts[3 * k + 2] = m == 3 ? f[q * m + 2] : 0.0;
}

// keys of one coordinate of the z-sorted points (for the (x, y) ordering of phase 1), and the final gather
__global__ void hv_coord_keys_kernel(const double *pts, unsigned n, int coord, const unsigned *idx_in, unsigned long long *keys, unsigned *idx)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned src = idx_in ? idx_in[i] : i;
    unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(pts[3 * static_cast<size_t>(src) + coord]));
    keys[i] = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
    idx[i] = src;
}

__global__ void hv_gather_xorder_kernel(const double *pts, const unsigned *xorder, unsigned n, double *ptsx)
{
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const unsigned q = xorder[k]; // position in the z-order
    ptsx[4 * k] = pts[3 * static_cast<size_t>(q)];
    ptsx[4 * k + 1] = pts[3 * static_cast<size_t>(q) + 1];
    ptsx[4 * k + 2] = pts[3 * static_cast<size_t>(q) + 2];
    ptsx[4 * k + 3] = static_cast<double>(q);
}

// ---- 2 objectives, indicator only: the reference's own sweep (hv2d::compute, hv_hv2d.cpp:59-84) is a sort by y, a running
// maximum of the width r.x - x and a sum - i.e. a radix sort, an inclusive max-scan and a reduction.
__global__ void hv2d_ykeys_kernel(const double *f, unsigned n, unsigned long long *keys, unsigned *idx)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(f[2 * static_cast<size_t>(i) + 1]));
    keys[i] = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
    idx[i] = i;
}
__global__ void hv2d_width_kernel(const double *f, const unsigned *order, unsigned n, double rx, double *w)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = rx - f[2 * static_cast<size_t>(order[i])];
}
__global__ void hv2d_terms_kernel(const double *f, const unsigned *order, const double *wmax, unsigned n, double ry, double *terms)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double y = f[2 * static_cast<size_t>(order[i]) + 1];
    const double ynext = i + 1 < n ? f[2 * static_cast<size_t>(order[i + 1]) + 1] : ry;
    terms[i] = (ynext - y) * wmax[i]; // :75-80
}
struct MaxOp {
    __device__ __forceinline__ double operator()(double a, double b) const { return fmax(a, b); }
};

// ---- four and more objectives: WFG (hv_hvwfg.cpp:153-296) ---------------------------------------------------------------------
// hv(S) with S sorted by the last objective, larger first, is sum_i |p_i[last] - r[last]| * (vol_{d-1}(p_i) - hv_{d-1}(L_i)) where
// L_i = the non-dominated subset of { max(p_i, p_j) : j > i } with the last objective dropped (it is p_i[last] for all of them).
// The TERMS of that sum are independent: one thread per term, each running the recursion below it sequentially on its own frames
// (m - 2 levels of at most `cap` rows, in global memory).  The recursion goes down to one objective, where the hypervolume of a
// limit set is r_0 - min_j max(p_0, q_0) - no separate two-objective sweep.  hvwfg::contributions (:92-117) is the same kernel
// over n frames: frame p = the limit set of p against every other point, contribution(p) = vol(p) - hv(frame p).
constexpr unsigned kWfgMaxM = 12;

struct WfgParams {
    const double *frames;  // [nframes x cap x m]: each frame sorted by objective m-1, larger first
    const unsigned *sizes; // [nframes]; nullptr: every frame holds `cap` rows
    unsigned nframes, cap, m; // m = row stride (objectives of the problem)
    unsigned d;               // objectives the frames live in (m for contributions, m - 1 for the second level of compute)
    double r[kWfgMaxM];
    double *arena;         // per thread of a launch: (d - 2) x cap x m doubles
    double *terms;         // [nframes x cap]
    unsigned long long first, count; // this launch covers the (frame, i) pairs first .. first + count - 1, numbered frame * cap + i
};

// rows of `stride` doubles; 1: a dominates-or-equals b on the first d coordinates (minimisation)
__device__ __forceinline__ bool wfg_weakly_dominates(const double *a, const double *b, unsigned d)
{
    for (unsigned c = 0; c < d; ++c)
        if (a[c] > b[c]) return false;
    return true;
}

// limitset (:153-209): out <- non-dominated subset of { max(p, F[j]) on the first d coordinates : j in [begin, k), j != skip }
__device__ unsigned wfg_limitset(const double *F, unsigned k, unsigned m, unsigned d, unsigned begin, unsigned skip, const double *p, double *out)
{
    unsigned no = 0;
    for (unsigned j = begin; j < k; ++j) {
        if (j == skip) continue;
        double *s = out + static_cast<size_t>(no) * m;
        for (unsigned c = 0; c < d; ++c) s[c] = fmax(F[static_cast<size_t>(j) * m + c], p[c]);
        bool keep = true;
        for (unsigned q = 0; q < no && keep; ++q) {
            const double *g = out + static_cast<size_t>(q) * m;
            bool g_dominates = true, equal = true;
            for (unsigned c = 0; c < d; ++c) {
                g_dominates = g_dominates && g[c] <= s[c];
                equal = equal && g[c] == s[c];
            }
            keep = !(g_dominates && !equal); // an equal row is replaced by s below, as the reference does
        }
        if (!keep) continue;
        unsigned prev = 0;
        for (unsigned q = 0; q < no; ++q) {
            const double *g = out + static_cast<size_t>(q) * m;
            if (!wfg_weakly_dominates(s, g, d)) {
                if (prev < q)
                    for (unsigned c = 0; c < d; ++c) out[static_cast<size_t>(prev) * m + c] = g[c];
                ++prev;
            }
        }
        if (prev < no)
            for (unsigned c = 0; c < d; ++c) out[static_cast<size_t>(prev) * m + c] = s[c];
        no = prev + 1u;
    }
    return no;
}

// cmp_points (:303-313): larger objective d-1 first, ties by the earlier objectives; insertion sort (frames are short and the
// dominance filter above is quadratic anyway)
__device__ void wfg_sort_desc(double *F, unsigned k, unsigned m, unsigned d)
{
    double row[kWfgMaxM];
    for (unsigned i = 1; i < k; ++i) {
        for (unsigned c = 0; c < d; ++c) row[c] = F[static_cast<size_t>(i) * m + c];
        unsigned j = i;
        while (j > 0) {
            const double *q = F + static_cast<size_t>(j - 1) * m;
            bool before = false; // row sorts before q
            for (unsigned c = d; c-- > 0;) {
                if (row[c] > q[c]) { before = true; break; }
                if (row[c] < q[c]) break;
            }
            if (!before) break;
            for (unsigned c = 0; c < d; ++c) F[static_cast<size_t>(j) * m + c] = q[c];
            --j;
        }
        if (j != i)
            for (unsigned c = 0; c < d; ++c) F[static_cast<size_t>(j) * m + c] = row[c];
    }
}

__device__ __forceinline__ double wfg_volume(const double *p, const double *r, unsigned d)
{
    double v = 1.0;
    for (unsigned c = 0; c < d; ++c) v *= (p[c] - r[c]);
    return fabs(v);
}

// hypervolume of the k rows at arena level 0 on their first d0 coordinates (compute_hv, :227-296), recursion unrolled on a stack
__device__ double wfg_hv(double *arena, unsigned cap, unsigned m, const double *r, unsigned k0, unsigned d0)
{
    struct Level {
        unsigned k, i;
        double H, a, incl;
    } st[kWfgMaxM];
    const size_t level_stride = static_cast<size_t>(cap) * m;
    unsigned lev = 0;
    wfg_sort_desc(arena, k0, m, d0);
    st[0] = {k0, 0u, 0.0, 0.0, 0.0};
    for (;;) {
        Level &L = st[lev];
        const unsigned d = d0 - lev;
        double *F = arena + lev * level_stride;
        if (L.i == L.k) {
            const double ret = L.H;
            if (lev == 0u) return ret;
            --lev;
            Level &P = st[lev];
            P.H += fabs(P.a * (P.incl - ret));
            ++P.i;
            continue;
        }
        const double *p = F + static_cast<size_t>(L.i) * m;
        const double a = p[d - 1u] - r[d - 1u], incl = wfg_volume(p, r, d - 1u);
        if (d == 1u) { // a set of numbers: r_0 - min
            double mn = p[0];
            for (unsigned j = 1; j < L.k; ++j) mn = fmin(mn, F[static_cast<size_t>(j) * m]);
            L.H = r[0] - mn;
            L.i = L.k;
            continue;
        }
        if (d == 2u) { // the limit sets below are numbers: hv_1 = r_0 - min_j max(p_0, q_0)
            double sub = 0.0;
            if (L.i + 1u < L.k) {
                double mn = INFINITY;
                for (unsigned j = L.i + 1u; j < L.k; ++j) mn = fmin(mn, fmax(p[0], F[static_cast<size_t>(j) * m]));
                sub = r[0] - mn;
            }
            L.H += fabs(a * (incl - sub));
            ++L.i;
            continue;
        }
        double *C = F + level_stride;
        const unsigned no = wfg_limitset(F, L.k, m, d - 1u, L.i + 1u, 0xffffffffu, p, C);
        if (no <= 1u) { // exclusive_hv, :212-224
            L.H += fabs(a * (incl - (no ? wfg_volume(C, r, d - 1u) : 0.0)));
            ++L.i;
            continue;
        }
        wfg_sort_desc(C, no, m, d - 1u);
        L.a = a;
        L.incl = incl;
        ++lev;
        st[lev] = {no, 0u, 0.0, 0.0, 0.0};
    }
}

__global__ void wfg_terms_kernel(const WfgParams P)
{
    const unsigned long long t = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= P.count) return;
    const unsigned long long e = P.first + t;
    const unsigned fr = static_cast<unsigned>(e / P.cap), i = static_cast<unsigned>(e % P.cap);
    const unsigned k = P.sizes ? P.sizes[fr] : P.cap;
    if (i >= k) return;
    const unsigned m = P.m, d = P.d;
    const double *F = P.frames + static_cast<size_t>(fr) * P.cap * m;
    const double *p = F + static_cast<size_t>(i) * m;
    const double a = p[d - 1u] - P.r[d - 1u], incl = wfg_volume(p, P.r, d - 1u);
    double sub = 0.0;
    if (d == 2u) { // the limit set is a set of numbers
        if (i + 1u < k) {
            double mn = INFINITY;
            for (unsigned j = i + 1u; j < k; ++j) mn = fmin(mn, fmax(p[0], F[static_cast<size_t>(j) * m]));
            sub = P.r[0] - mn;
        }
    } else {
        double *arena = P.arena + t * (static_cast<size_t>(d - 2u) * P.cap * m);
        const unsigned no = wfg_limitset(F, k, m, d - 1u, i + 1u, 0xffffffffu, p, arena);
        if (no == 1u) sub = wfg_volume(arena, P.r, d - 1u);
        else if (no > 1u) sub = wfg_hv(arena, P.cap, m, P.r, no, d - 1u);
    }
    P.terms[e] = fabs(a * (incl - sub));
}

// frame p <- the limit set of point p, sorted for wfg_terms_kernel.  all_others != 0: against every other point, in all m objectives
// (hvwfg::contributions, :102-105); else against the points after p in the sorted top frame, last objective dropped (compute_hv's
// loop, :287-291)
__global__ void wfg_frames_kernel(const double *f, unsigned n, unsigned m, int all_others, double *frames, unsigned *sizes)
{
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    double *out = frames + static_cast<size_t>(p) * n * m;
    const unsigned d = all_others ? m : m - 1u;
    const unsigned no = wfg_limitset(f, n, m, d, all_others ? 0u : p + 1u, all_others ? p : 0xffffffffu, f + static_cast<size_t>(p) * m, out);
    wfg_sort_desc(out, no, m, d);
    sizes[p] = no;
}

// per frame, in index order as compute_hv accumulates: H = sum of the frame's terms;
// contributions: out[p] = vol_m(p) - H; compute: out[p] = |p[m-1] - r[m-1]| * (vol_{m-1}(p) - H), the top-level term of point p
__global__ void wfg_sum_kernel(const double *terms, const unsigned *sizes, unsigned cap, unsigned nframes, const double *f, unsigned m,
                               const WfgParams P, int compute, double *out)
{
    const unsigned fr = blockIdx.x * blockDim.x + threadIdx.x;
    if (fr >= nframes) return;
    const unsigned k = sizes[fr];
    double H = 0.0;
    for (unsigned i = 0; i < k; ++i) H += terms[static_cast<size_t>(fr) * cap + i];
    const double *p = f + static_cast<size_t>(fr) * m;
    out[fr] = compute ? fabs((p[m - 1u] - P.r[m - 1u]) * (wfg_volume(p, P.r, m - 1u) - H)) : wfg_volume(p, P.r, m) - H;
}

__global__ void wfg_total_kernel(const double *top_terms, unsigned n, double *out)
{
    if (blockIdx.x || threadIdx.x) return;
    double H = 0.0;
    for (unsigned i = 0; i < n; ++i) H += top_terms[i];
    out[0] = H;
}

// top frame: rows in decreasing order of the last objective (cmp_points; the order among equal keys does not change the sum's terms
// beyond rounding).  keys: order-preserving bits of the last objective, flipped
__global__ void wfg_last_keys_kernel(const double *f, unsigned n, unsigned m, unsigned long long *keys, unsigned *idx)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(f[static_cast<size_t>(i) * m + m - 1u]));
    b = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
    keys[i] = ~b;
    idx[i] = i;
}

__global__ void wfg_gather_rows_kernel(const double *f, const unsigned *order, unsigned n, unsigned m, double *frame)
{
    const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<size_t>(n) * m) return;
    frame[e] = f[static_cast<size_t>(order[e / m]) * m + e % m];
}

__global__ void wfg_check_kernel(const double *f, unsigned n, const WfgParams P, int *bad)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool outside = false, all_equal = true; // assert_minimisation, hv_algorithm.cpp:226-258
    for (unsigned c = 0; c < P.m; ++c) {
        const double v = f[static_cast<size_t>(i) * P.m + c];
        outside = outside || P.r[c] < v;
        all_equal = all_equal && P.r[c] == v;
    }
    if (outside || all_equal) *bad = 1;
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

int hv_wfg_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t m, const double *r, int compute, double *d_out, cudaStream_t st)
{
    PGC_REQUIRE(m >= 3 && m <= kWfgMaxM, "hypervolume (WFG): between 3 and %u objectives, %zu requested", kWfgMaxM, m);
    PGC_REQUIRE(n < 0x7fffffffull, "hypervolume: too many points");
    if (n == 0) {
        if (compute) PGC_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), st));
        return PGC_OK;
    }
    Scratch tmp(st);
    const unsigned un = static_cast<unsigned>(n), um = static_cast<unsigned>(m);
    WfgParams P{};
    P.m = um;
    P.cap = un;
    for (unsigned c = 0; c < um; ++c) P.r[c] = r[c];
    int rc, *d_bad = nullptr;
    if ((rc = tmp.get(&d_bad, 1))) return rc;
    PGC_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    wfg_check_kernel<<<(un + 255) / 256, 256, 0, st>>>(d_f, un, P, d_bad);
    int bad = 0;
    PGC_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    PGC_REQUIRE(!bad, "Reference point is invalid: another point seems to be outside the reference point boundary, or be equal to it");
    // both modes: n frames (one per point) and one thread per (frame, term); compute works one level down (m - 1 objectives)
    double *frames = nullptr, *terms = nullptr, *top = nullptr, *top_terms = nullptr;
    unsigned *sizes = nullptr;
    const size_t frame_doubles = n * m;
    PGC_REQUIRE(static_cast<double>(n) * static_cast<double>(frame_doubles) * 8. < 16e9,
                "hypervolume (WFG): %zu points in %zu objectives need more than 16 GB of frames", n, m);
    if ((rc = tmp.get(&frames, n * frame_doubles)) || (rc = tmp.get(&terms, n * n)) || (rc = tmp.get(&sizes, n))) return rc;
    const double *src = d_f;
    unsigned launches = 4;
    if (compute) {
        unsigned long long *k0 = nullptr, *k1 = nullptr;
        unsigned *i0 = nullptr, *order = nullptr;
        if ((rc = tmp.get(&top, frame_doubles)) || (rc = tmp.get(&top_terms, n)) || (rc = tmp.get(&k0, n)) || (rc = tmp.get(&k1, n))
            || (rc = tmp.get(&i0, n)) || (rc = tmp.get(&order, n)))
            return rc;
        wfg_last_keys_kernel<<<(un + 255) / 256, 256, 0, st>>>(d_f, un, um, k0, i0);
        size_t bytes = 0;
        PGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0, k1, i0, order, static_cast<int>(n), 0, 64, st));
        unsigned char *ws = nullptr;
        if ((rc = tmp.get(&ws, bytes))) return rc;
        PGC_CUDA(cub::DeviceRadixSort::SortPairs(ws, bytes, k0, k1, i0, order, static_cast<int>(n), 0, 64, st));
        wfg_gather_rows_kernel<<<static_cast<unsigned>((frame_doubles + 255) / 256), 256, 0, st>>>(d_f, order, un, um, top);
        src = top;
        launches += 3;
    }
    wfg_frames_kernel<<<(un + 31) / 32, 32, 0, st>>>(src, un, um, compute ? 0 : 1, frames, sizes);
    PGC_CUDA(cudaGetLastError());
    P.d = compute ? um - 1u : um;
    // in batches bounded by the per-thread recursion frames (~2 GiB per launch)
    const size_t per_thread = std::max<size_t>(static_cast<size_t>(P.d - 2u) * n * m, 1); // doubles
    const unsigned long long total = static_cast<unsigned long long>(n) * n;
    unsigned long long batch = (size_t(2) << 30) / (per_thread * sizeof(double));
    if (batch < 256) batch = 256;
    if (batch > total) batch = total;
    double *arena = nullptr;
    if ((rc = tmp.get(&arena, static_cast<size_t>(batch) * per_thread))) return rc;
    P.frames = frames;
    P.sizes = sizes;
    P.nframes = un;
    P.arena = arena;
    P.terms = terms;
    for (unsigned long long first = 0; first < total; first += batch) {
        P.first = first;
        P.count = std::min(batch, total - first);
        wfg_terms_kernel<<<static_cast<unsigned>((P.count + 63) / 64), 64, 0, st>>>(P);
        PGC_CUDA(cudaGetLastError());
        ++launches;
    }
    wfg_sum_kernel<<<(un + 63) / 64, 64, 0, st>>>(terms, sizes, un, un, src, um, P, compute, compute ? top_terms : d_out);
    if (compute) wfg_total_kernel<<<1, 1, 0, st>>>(top_terms, un, d_out);
    PGC_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(launches, std::memory_order_relaxed);
    return PGC_OK;
}

// mode 0: d_out[n] = exclusive contributions; mode 1: d_out[0] = hypervolume (d_out needs n doubles of space)
int hv_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t m, const double *r, int compute, double *d_out, cudaStream_t st)
{
    PGC_REQUIRE(m >= 2, "hypervolume: at least 2 objectives are needed, %zu requested", m);
    if (m >= 4) return hv_wfg_device(ctx, d_f, n, m, r, compute, d_out, st); // hypervolume::get_best_compute, hypervolume.cpp:208-220
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
    if (m == 2 && compute) {
        unsigned long long *k0 = nullptr, *k1 = nullptr;
        unsigned *i0 = nullptr, *ord = nullptr;
        double *w = nullptr, *wmax = nullptr, *terms2 = nullptr;
        if ((rc = tmp.get(&k0, n)) || (rc = tmp.get(&k1, n)) || (rc = tmp.get(&i0, n)) || (rc = tmp.get(&ord, n)) || (rc = tmp.get(&w, n))
            || (rc = tmp.get(&wmax, n)) || (rc = tmp.get(&terms2, n)))
            return rc;
        hv2d_ykeys_kernel<<<(un + 255) / 256, 256, 0, st>>>(d_f, un, k0, i0);
        size_t b1 = 0, b2 = 0, b3 = 0;
        PGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b1, k0, k1, i0, ord, static_cast<int>(n), 0, 64, st));
        PGC_CUDA(cub::DeviceScan::InclusiveScan(nullptr, b2, w, wmax, MaxOp(), static_cast<int>(n), st));
        PGC_CUDA(cub::DeviceReduce::Sum(nullptr, b3, terms2, d_out, static_cast<int>(n), st));
        unsigned char *ws = nullptr;
        if ((rc = tmp.get(&ws, std::max(b1, std::max(b2, b3))))) return rc;
        PGC_CUDA(cub::DeviceRadixSort::SortPairs(ws, b1, k0, k1, i0, ord, static_cast<int>(n), 0, 64, st));
        hv2d_width_kernel<<<(un + 255) / 256, 256, 0, st>>>(d_f, ord, un, r[0], w);
        PGC_CUDA(cub::DeviceScan::InclusiveScan(ws, b2, w, wmax, MaxOp(), static_cast<int>(n), st));
        hv2d_terms_kernel<<<(un + 255) / 256, 256, 0, st>>>(d_f, ord, wmax, un, r[1], terms2);
        PGC_CUDA(cub::DeviceReduce::Sum(ws, b3, terms2, d_out, static_cast<int>(n), st));
        PGC_CUDA(cudaGetLastError());
        ctx->launches.fetch_add(6, std::memory_order_relaxed);
        return PGC_OK;
    }
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
    // First pass: one warp per point with the staircase in shared memory (hv_warp_kernel).  The (rare) points whose staircase
    // outgrows it are collected and rerun by the one-thread-per-point kernel with the worst-case capacity n in global memory,
    // in batches bounded to ~1 GiB of scratch.
    const unsigned cap1 = 0; // the second pass, when needed, always uses cap = n
    double *pts = nullptr, *terms = nullptr;
    unsigned *ovf = nullptr;
    if ((rc = tmp.get(&pts, 3 * n)) || (rc = tmp.get(&ovf, n + 1))) return rc;
    if (compute && (rc = tmp.get(&terms, n))) return rc;
    PGC_CUDA(cudaMemsetAsync(ovf, 0, sizeof(unsigned), st));
    hv_gather_sorted_kernel<<<(un + 255) / 256, 256, 0, st>>>(d_f, order, un, static_cast<unsigned>(m), pts);
    double *ptsx = nullptr;
    {   // (x, y) lexicographic order of the z-sorted points: stable radix sort by y, then by x
        unsigned long long *ka = nullptr, *kb = nullptr;
        unsigned *ia = nullptr, *ib = nullptr, *ic = nullptr;
        if ((rc = tmp.get(&ptsx, 4 * n)) || (rc = tmp.get(&ka, n)) || (rc = tmp.get(&kb, n)) || (rc = tmp.get(&ia, n)) || (rc = tmp.get(&ib, n))
            || (rc = tmp.get(&ic, n)))
            return rc;
        size_t bytes = 0;
        PGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, ka, kb, ia, ib, static_cast<int>(n), 0, 64, st));
        unsigned char *ws = nullptr;
        if ((rc = tmp.get(&ws, bytes))) return rc;
        hv_coord_keys_kernel<<<(un + 255) / 256, 256, 0, st>>>(pts, un, 1, nullptr, ka, ia);
        PGC_CUDA(cub::DeviceRadixSort::SortPairs(ws, bytes, ka, kb, ia, ib, static_cast<int>(n), 0, 64, st));
        hv_coord_keys_kernel<<<(un + 255) / 256, 256, 0, st>>>(pts, un, 0, ib, ka, ia);
        PGC_CUDA(cub::DeviceRadixSort::SortPairs(ws, bytes, ka, kb, ia, ic, static_cast<int>(n), 0, 64, st));
        hv_gather_xorder_kernel<<<(un + 255) / 256, 256, 0, st>>>(pts, ic, un, ptsx);
        ctx->launches.fetch_add(5, std::memory_order_relaxed);
    }
    HvWarpParams W{pts, ptsx, order, un, r[0], r[1], rz, compute ? terms : d_out, ovf, compute};
    constexpr size_t kStairBytes = sizeof(double) * kHvWarps * 2 * kHvCap;
    PGC_CUDA(cudaFuncSetAttribute(hv_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kStairBytes)));
    hv_warp_kernel<<<(un + kHvWarps - 1) / kHvWarps, kHvWarps * 32, kStairBytes, st>>>(W);
    PGC_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(2, std::memory_order_relaxed);
    HvParams P;
    P.f = d_f;
    P.order = order;
    P.n = un;
    P.m = static_cast<unsigned>(m);
    P.rx = r[0];
    P.ry = r[1];
    P.rz = rz;
    P.out = compute ? terms : d_out;
    P.compute = compute;
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
