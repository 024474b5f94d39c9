// eval_cec2014.cu - CEC2014 f1..f30 batch fitness on sm_100a (the headline path).
//
// Replaces, for a whole batch, reference pagmo::cec2014::fitness (src/problems/cec2014.cpp:119-247) as driven by
// thread_bfe (src/batch_evaluators/thread_bfe.cpp:94-137).
//
// Design:
//   * one "stage kernel" launch per recipe stage (cec2014_recipe.cpp).  A stage = sr_func (shift, scale,
//     rotate; cec2014.cpp:1238-1274) + optional hybrid permutation (:807-809) + 1..5 primitive groups.
//     Compositions (f23-f30) launch one stage kernel per component plus a tiny cf_cal kernel (:1319-1353).
//   * rotation z = Mr*y is the FP64-pipe-bound part (2*D^2 flop/eval).  Persistent CTAs (one per SM, 16 warps at D = 100)
//     keep the whole matrix in shared memory (row-major, padded stride).  Every warp is an independent worker
//     on tiles of 8 individuals with a private shared-memory buffer:
//     load+shift+scale -> FP64 GEMM on mma.sync.m8n8k4.f64 (DMMA; NT accumulator tiles per warp) -> z back
//     to the buffer -> primitive epilogue (4 lanes per individual).  Warps de-synchronise, so one warp's load /
//     epilogue overlaps the other warps' DMMA streams; no block-level barrier inside the tile loop.
//     Why DMMA and not SIMT DFMA: a register-tiled DFMA version (profiles/r1a_simt_*) was shared-memory-
//     bandwidth bound - an LDS costs (bytes per lane x 32)/128 cycles whatever the broadcast pattern, so a
//     4x13 register tile needs 1.3x more LSU cycles than FP64-pipe cycles and stalled at 42% pipe utilisation.
//     An m8n8k4 tile does 8 FMAs per lane per operand pair loaded (vs 3), and the measured DMMA ceiling on B200
//     (37.0 TFLOP/s) is not below the DFMA one (34.2 TFLOP/s).
//   * the inner-index accumulation order inside one DMMA is the hardware's; differences to the reference's
//     sequential non-fused sum (:1231-1233) are ~1e-16 relative per term.  The file is compiled with -fmad=false
//     so every expression outside the rotation keeps the reference's operation order and roundings.
#include <cfloat>
#include <cmath>
#include <cstring>

#include "cec_device.cuh"
#include "pgc_internal.cuh"

namespace pgc
{

namespace
{

using namespace cecdev;

constexpr int kMT = 1;              // m8n8k4 row tiles per warp tile
static_assert(kMT == 1 && cecdev::kTileInd == 8 * kMT, "warp tile = one m8n8k4 row tile");
#ifndef PGC_WARPS
#define PGC_WARPS 16
#endif
#ifndef PGC_GEMM_UNROLL
#define PGC_GEMM_UNROLL 2
#endif
constexpr int kGemmUnroll = PGC_GEMM_UNROLL; // 4-deep steps of the rotation unrolled together
constexpr int kWarps = PGC_WARPS;   // independent workers per CTA (kWarps / 4 per SM sub-partition)

// After the rotation the warp keeps z TRANSPOSED: zT[coordinate][individual], row stride kZS, so that the epilogue
// (lane = individual + kTileInd * q) reads consecutive doubles per coordinate whatever the permutation; the +2 pad
// makes the accumulator-tile stores conflict-free.
// (An un-padded, XOR-swizzled zT that makes room for 20 warps was measured: 0.80 ms per rotated launch with 16 warps and
// 0.87 ms with 20 warps at 96 registers, against 0.755 ms for this layout - profiles/r1o_variants.txt.)
constexpr int kZS = kTileInd + 2;
__device__ __forceinline__ int zt_index(int c, int ind) { return c * kZS + ind; }
__host__ __device__ constexpr int warp_buf_elems(int d)
{
    return (kTileInd * ystride(d) > pad8(d) * kZS) ? kTileInd * ystride(d) : pad8(d) * kZS;
}


struct StageParams {
    const double *x;    // [n x D]
    const double *mr;   // shared-memory image of this component's rotation (see make_rotation_image)
    const double *os;   // shift of this component: D
    const int *perm;    // 0-based permutation of this component (D) or nullptr
    const double *table; // problem constant table
    double *out;        // final f (non-composition) or fit[stage] (composition)
    double *wout;       // composition: w[stage]; nullptr otherwise
    long long n;
    double fbias;
    int aligned16;
    unsigned long long *prof; // optional (pgc_debug_cec2014_phase_cycles): per-phase cycle totals, see kPh*
    StageDesc st;
};

enum { kPhLoad = 0, kPhWeight, kPhTokenWait, kPhGemm, kPhStoreZ, kPhEpilogue, kPhTiles, kPhCount };


struct Elem {
    const double *zt;  // the warp's zT tile
    int ind;           // this lane's individual
    const int *idx;    // may be nullptr
    int off;
    __device__ __forceinline__ double operator()(int j) const // zT already carries the group's own sh_rate
    {
        const int jj = idx ? idx[off + j] : off + j;
        return zt[zt_index(jj, ind)];
    }
};


// ---- per-coordinate terms shared by the tile epilogue (eval_group) and the separable fast path ----------------------
template <bool CM = false> __device__ __forceinline__ double rastrigin_term(double z) // cec2014.cpp:541-543
{
    const double two_pi = 2.0 * 3.141592653589793238462643383279502884;
    return (z * z - 10.0 * cos_theta<CM>(two_pi * z) + 10.0);
}

// schwefel, cec2014.cpp:575-589, the three branches folded into one sin(sqrt(.)): f -= sub; f += pen
template <bool CM = false> __device__ __forceinline__ void schwefel_term(double zin, double inv_n, double &sub, double &pen)
{
    const double z = zin + 4.209687462275036e+002;
    const double az = fabs(z);
    const bool big = az > 500.0;
    // fm = fmod(|z|, 500) EXACTLY: q = nearest multiple, the fused remainder az - 500 q is exact and lies
    // in [-250, 250]; one fix-up brings it to [0, 500)
    const double q = round_magic(az * 0.002);
    double fm = fma(-q, 500.0, az);
    fm = (fm < 0.0) ? fm + 500.0 : fm;
    // z > 500: (500 - fmod(z,500)); z < -500: -(-500 + fmod(|z|,500)); else |z|
    const double m = big ? 500.0 - fm : az;
    const double t = (z - copysign(500.0, z)) * 0.01;
    // the separable kernel (CM) takes a branch-free square root so that the evaluations of a lane's coordinates interleave;
    // the IEEE sqrt's slow-path branch kept them in separate basic blocks (FP64 pipe 60 % busy, issue slots 61 %: latency bound)
    sub = copysign(m, z) * sin_theta<CM>(CM ? sqrt_nobranch(m) : sqrt(m));
    pen = big ? t * t * inv_n : 0.0;
}


// One primitive on n coordinates, evaluated by the kLPI lanes of an individual (lane q takes the terms
// j == q mod kLPI); every lane returns the full value.  Expressions follow cec2014.cpp term by term.
__device__ double eval_group(const GroupDesc &g, const Elem &v, const double *__restrict__ tab, int h)
{
    const int n = g.len;
    const double dn = static_cast<double>(static_cast<unsigned>(n));
    const int lo = h, hi = n;
    const double two_pi = 2.0 * 3.141592653589793238462643383279502884;
    const double *gt = tab + g.tab_off;
    switch (g.prim) {
        case P_ELLIPS: // :382-384
            return pair_add(ordered_sum(lo, hi, [&](int j) {
                const double z = v(j);
                return gt[j] * z * z;
            }));
        case P_BENT_CIGAR: // :395-398
            return pair_add(ordered_sum(lo, hi, [&](int j) {
                const double z = v(j);
                return (j == 0) ? z * z : 1.0e6 * z * z;
            }));
        case P_DISCUS: // :408-411
            return pair_add(ordered_sum(lo, hi, [&](int j) {
                const double z = v(j);
                return (j == 0) ? 1.0e6 * z * z : z * z;
            }));
        case P_ROSENBROCK: { // :438-444
            return pair_add(ordered_sum(h, n - 1, [&](int j) {
                const double zj = v(j) + 1.0, zn = v(j + 1) + 1.0;
                const double t1 = zj * zj - zn, t2 = zj - 1.0;
                return 100.0 * t1 * t1 + t2 * t2;
            }));
        }
        case P_ACKLEY: { // :476-482
            double s1 = ordered_sum(lo, hi, [&](int j) {
                const double z = v(j);
                return z * z;
            });
            double s2 = ordered_sum(lo, hi, [&](int j) { return cos_theta(two_pi * v(j)); });
            s1 = pair_add(s1);
            s2 = pair_add(s2);
            s1 = -0.2 * sqrt(s1 / dn);
            s2 /= dn;
            return 2.718281828459045235360287471352662498 - 20.0 * exp(s1) - exp(s2) + 20.0;
        }
        case P_WEIERSTRASS: { // :500-509
            // sum_k 0.5^k cos(theta_k), theta_k = fl(fl(2 pi 3^k) * u) as the reference forms it.  Term k = 0 is
            // evaluated directly from the reference's own argument (exact two-term reduction by 2 pi); the others come
            // from the angle-tripling map w -> w^3 on the unit circle (complex multiplication: errors grow exactly 3x
            // per step, so term k is off by ~3^k * 3e-16 and is weighted by 0.5^k: <= 1e-12 per coordinate at
            // k = 20).  That is below what separates ANY evaluation from the reference's: its theta_20 ~ 2e10 carries a
            // rounding error of ~2e-6 rad of its own, i.e. 2.4e-12 in the weighted term.  A second direct evaluation
            // at k = 10 (used until r1n) bought nothing measurable and cost 23 instructions per coordinate.
            // ~3.5x fewer FP64 instructions than 21 range-reduced cosines.
            return pair_add(ordered_sum(lo, hi, [&](int j) {
                       const double u = v(j) + 0.5;
                       double sum = 0.0, w = 1.0, sn = 0.0, cs = 1.0;
#pragma unroll
                       for (int k = 0; k <= 20; ++k) {
                           if (k == 0) {
                               sincos_turns(turns_of(gt[k] * u), sn, cs);
                           } else {
                               triple_angle(sn, cs);
                           }
                           sum = fma(w, cs, sum); // 0.5^k * cos is an exact scaling: == "sum += pow(a,k)*cos"
                           w *= 0.5;
                       }
                       return sum;
                   }))
                   - g.c0;
        }
        case P_GRIEWANK: { // :524-528; gt[j] = 1/sqrt(1+j) (the reference divides by sqrt(1+j): <= 1 ulp apart)
            double s = ordered_sum(lo, hi, [&](int j) {
                const double z = v(j);
                return z * z;
            });
            double p = 1.0;
            int j = lo;
            for (; j + kLPI < hi; j += 2 * kLPI) {
                const double a = cos_theta(v(j) * gt[j]), b = cos_theta(v(j + kLPI) * gt[j + kLPI]);
                p *= a;
                p *= b;
            }
            if (j < hi) p *= cos_theta(v(j) * gt[j]);
            s = pair_add(s);
            p = pair_mul(p);
            return 1.0 + s / 4000.0 - p;
        }
        case P_RASTRIGIN: // :541-543
            return pair_add(ordered_sum(lo, hi, [&](int j) { return rastrigin_term(v(j)); }));
        case P_SCHWEFEL: { // :575-589, the three branches folded into one sin(sqrt(.)) per coordinate
            const double inv_n = 1.0 / dn;
            double s = 0.0;
            auto term = [&](int j, double &sub, double &pen) { schwefel_term(v(j), inv_n, sub, pen); };
            int j = lo;
            for (; j + kLPI < hi; j += 2 * kLPI) {
                double s0, p0, s1, p1;
                term(j, s0, p0);
                term(j + kLPI, s1, p1);
                s -= s0;
                s += p0;
                s -= s1;
                s += p1;
            }
            if (j < hi) {
                double s0, p0;
                term(j, s0, p0);
                s -= s0;
                s += p0;
            }
            return pair_add(s) + g.c0;
        }
        case P_KATSUURA: { // :604-614
            // |2^k z - floor(2^k z + 0.5)| is the distance to the nearest integer: same value via round-to-nearest
            // (magic-constant add, valid for |2^k z| < 2^51).  prod_j b_j^c0 is taken as exp(c0 * log(prod_j b_j)): each
            // lane multiplies its own factors b_j in [1, 1 + n/2] (at most 25 of them: <= 5e42) and takes ONE log.
            double prod = ordered_prod(lo, hi, [&](int j) {
                const double temp = katsuura_inner(v(j));
                return 1.0 + static_cast<double>(j + 1) * temp;
            });
            double slog = pair_add(log(prod));
            const double p = exp(g.c0 * slog);
            return p * g.c1 - g.c1;
        }
        case P_HAPPYCAT:
        case P_HGBAT: { // :751-759, :774-782
            double r2 = ordered_sum(lo, hi, [&](int j) {
                const double z = v(j) - 1.0;
                return z * z;
            });
            double sz = ordered_sum(lo, hi, [&](int j) { return v(j) - 1.0; });
            r2 = pair_add(r2);
            sz = pair_add(sz);
            if (g.prim == P_HAPPYCAT) return sqrt(sqrt(fabs(r2 - dn))) + (0.5 * r2 + sz) / dn + 0.5;
            return sqrt(fabs(r2 * r2 - sz * sz)) + (0.5 * r2 + sz) / dn + 0.5;
        }
        case P_GRIE_ROSEN: // :702-713 (cyclic last term)
            return pair_add(ordered_sum(lo, hi, [&](int j) {
                const int jn = (j + 1 == n) ? 0 : j + 1;
                const double zj = v(j) + 1.0, zn = v(jn) + 1.0;
                const double t1 = zj * zj - zn, t2 = zj - 1.0;
                const double temp = 100.0 * t1 * t1 + t2 * t2;
                return (temp * temp) * 2.5e-4 - cos_theta(temp) + 1.0; // "/ 4000.0" as a multiplication (<= 1 ulp apart)
            }));
        case P_ESCAFFER6: // :727-736 (cyclic last term)
            return pair_add(ordered_sum(lo, hi, [&](int j) {
                const int jn = (j + 1 == n) ? 0 : j + 1;
                const double a = v(j), b = v(jn);
                const double ss = a * a + b * b;
                // sin^2(w) - 0.5 = -0.5 cos(2 w), 2 w = sqrt(4 ss); the quotient through a Newton reciprocal (~1 ulp)
                const double c2 = cos_theta(sqrt(4.0 * ss));
                const double t2 = 1.0 + 0.001 * ss;
                return 0.5 - (0.5 * c2) * fast_rcp(t2 * t2);
            }));
        default: return 0.0;
    }
}

template <int D, bool ROT>
__global__ void __launch_bounds__(kWarps * 32, 1) cec14_stage_kernel(const __grid_constant__ StageParams P)
{
    constexpr int DP = pad8(D);     // padded output count (rows of the rotation image)
    constexpr int KP = pad4(D);     // padded inner length
    constexpr int NT = DP / 8;      // 8-wide output tiles per warp
    constexpr int YS = ystride(D);  // row stride of Y tile and rotation image
    constexpr int WB = warp_buf_elems(D);
    constexpr int MR_ELEMS = ROT ? DP * YS : 0;
    // L phase geometry: a row (individual) is D/2 16-byte chunks, fetched by LPR lanes in PASS passes
    constexpr int CH = D / 2;
    constexpr int PASS = (CH + 31) / 32;
    constexpr int LPR = (CH + PASS - 1) / PASS;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sMr = reinterpret_cast<double *>(smem_raw);
    double *sBuf = sMr + MR_ELEMS;
    double *sOs = sBuf + kWarps * WB;
    double *sRate = sOs + D; // per coordinate of z: the sh_rate of the group that reads it (hybrids; 1.0 otherwise)
    int *sPerm = reinterpret_cast<int *>(sRate + DP);

    // ---- per-CTA preload: rotation image (built on the host), shift, permutation -----------------------------
    if (ROT) {
        const double2 *src = reinterpret_cast<const double2 *>(P.mr);
        double2 *dst = reinterpret_cast<double2 *>(sMr);
        for (int i = threadIdx.x; i < MR_ELEMS / 2; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        sOs[i] = P.os[i];
        sPerm[i] = P.perm ? P.perm[i] : i;
    }
    bool unit_rate = true; // every group scale is 1.0 (all basic functions): skip the multiplies altogether
    for (int gi = 0; gi < P.st.ngroups; ++gi) unit_rate = unit_rate && P.st.g[gi].rate == 1.0;
    for (int i = threadIdx.x; i < DP; i += blockDim.x) sRate[i] = 1.0;
    __syncthreads();
    if (!unit_rate) {
        for (int gi = 0; gi < P.st.ngroups; ++gi) {
            const GroupDesc &g = P.st.g[gi];
            for (int j = threadIdx.x; j < g.len; j += blockDim.x) sRate[P.st.permute ? sPerm[g.off + j] : g.off + j] = g.rate;
        }
        __syncthreads();
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *buf = sBuf + warp * WB;
    const int *perm = P.st.permute ? sPerm : nullptr;
    const long long ntiles = (P.n + kTileInd - 1) / kTileInd;
    const bool need_w = P.wout != nullptr;
    const double pre_rate = P.st.pre_rate;
    const bool unit_pre = pre_rate == 1.0;
    const int et = lane & (kTileInd - 1), eq = lane / kTileInd; // epilogue mapping: individual, term class

    // this lane's slice of the shift vector for the L phase (same columns for every row of every tile)
    double2 osv[PASS];
#pragma unroll
    for (int ps = 0; ps < PASS; ++ps) {
        const int c = ps * LPR + lane;
        osv[ps] = (lane < LPR && c < CH) ? make_double2(sOs[2 * c], sOs[2 * c + 1]) : make_double2(0.0, 0.0);
    }

    // The four warps of an SM sub-partition share its FP64 pipe (DMMA and DFMA issue to the same pipe).  Identical
    // warps started together stay in lock step - all loading, then all in the GEMM, then all in the epilogue - and
    // the pipe idles in between.  A one-off stagger of a quarter period per warp keeps the phases spread for the
    // whole launch (phase lengths are deterministic).
    if (ROT) {
        const long long wait = static_cast<long long>(warp >> 2) * (NT * (KP / 4) * 16 * kMT);
        const long long t_start = clock64();
        while (clock64() - t_start < wait) {
        }
    }

    for (long long tile = static_cast<long long>(blockIdx.x) * kWarps + warp; tile < ntiles;
         tile += static_cast<long long>(gridDim.x) * kWarps) {
        const long long t0 = tile * kTileInd;
        const int nt = (P.n - t0 < kTileInd) ? static_cast<int>(P.n - t0) : kTileInd;
        long long tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0, tp4 = 0, tp5 = 0, tp6 = 0;
        if (P.prof) tp0 = clock64();
        double wacc = 0.0;

        // ---- L: coalesced load (every 16-byte load of the tile in flight at once), shift (x - Os) and scale
        // (* sh_rate), cec2014.cpp:1245-1258.  ROT: row-major Y tile for the DMMA; otherwise straight into zT.
        {
            const double *src = P.x + t0 * D;
            const double scale = need_w ? 1.0 : pre_rate; // unaligned path: keep x-Os for the weight pass first
            if (P.aligned16) {
                double wrow[kTileInd];
#pragma unroll
                for (int t = 0; t < kTileInd; ++t) wrow[t] = 0.0;
                double2 xv[kTileInd][PASS];
#pragma unroll
                for (int t = 0; t < kTileInd; ++t)
#pragma unroll
                    for (int ps = 0; ps < PASS; ++ps) {
                        const int c = ps * LPR + lane;
                        xv[t][ps] = make_double2(0.0, 0.0);
                        if (lane < LPR && c < CH && t < nt)
                            xv[t][ps] = __ldcs(reinterpret_cast<const double2 *>(src + t * D) + c);
                    }
#pragma unroll
                for (int t = 0; t < kTileInd; ++t)
#pragma unroll
                    for (int ps = 0; ps < PASS; ++ps) {
                        const int c = ps * LPR + lane;
                        if (lane < LPR && c < CH) {
                            const double d0 = xv[t][ps].x - osv[ps].x, d1 = xv[t][ps].y - osv[ps].y;
                            if (need_w) wrow[t] += d0 * d0 + d1 * d1; // cf_cal weight sum_j (x_j - Os_j)^2, :1330-1332
                            // sh_rate = 1 (ellips, bent_cigar, discus, ackley, escaffer6, the hybrids): x * 1.0 is x, skip the multiply
                            double y0 = d0, y1 = d1;
                            if (!unit_pre) {
                                y0 = d0 * pre_rate;
                                y1 = d1 * pre_rate;
                            }
                            if (ROT) {
                                *reinterpret_cast<double2 *>(buf + t * YS + 2 * c) = make_double2(y0, y1);
                            } else {
                                buf[zt_index(2 * c, t)] = y0;
                                buf[zt_index(2 * c + 1, t)] = y1;
                            }
                        }
                    }
                if (need_w) {
#pragma unroll
                    for (int t = 0; t < kTileInd; ++t) {
                        double w = wrow[t];
#pragma unroll
                        for (int m = 16; m > 0; m >>= 1) w += __shfl_xor_sync(kFull, w, m);
                        if (t == et) wacc = w;
                    }
                }
            } else { // 8-byte aligned input only (a shard that starts mid-allocation): scalar loads
                for (int e = lane; e < kTileInd * D; e += 32) {
                    const int t = e / D, j = e - t * D;
                    const double xj = (t < nt) ? __ldcs(src + e) : 0.0;
                    const double y = (xj - sOs[j]) * scale;
                    if (ROT) buf[t * YS + j] = y;
                    else buf[zt_index(j, t)] = y;
                }
            }
            if (ROT && KP > D) { // zero the inner-index padding (zT of the previous tile lived there)
                if (lane < kTileInd)
                    for (int j = D; j < KP; ++j) buf[lane * YS + j] = 0.0;
            }
        }
        __syncwarp();
        if (P.prof) tp1 = clock64();

        if (need_w && !P.aligned16) { // unaligned path: weight pass over the tile, then the deferred scale
            for (int j = eq; j < D; j += kLPI) {
                double *pd = ROT ? buf + et * YS + j : buf + zt_index(j, et);
                const double d = *pd;
                wacc += d * d;
                *pd = d * pre_rate;
            }
            wacc = pair_add(wacc);
            __syncwarp();
        }
        if (P.prof) tp2 = clock64();

        // ---- G: z = Mr * y (:1224-1235) on DMMA.  A = Y tile (rows = individuals), B = Mr^T (rows of the image =
        // outputs); both fragments are "row g = lane/4, inner index 4u + lane%4": one 8-byte load per row and
        // 4-deep step, kMT + NT loads feeding kMT*NT independent m8n8k4 accumulator tiles.
        if (ROT) {
            if (P.prof) tp3 = clock64();
            const int g = lane >> 2, j = lane & 3;
            double acc[kMT][NT][2];
#pragma unroll
            for (int mt = 0; mt < kMT; ++mt)
#pragma unroll
                for (int nt2 = 0; nt2 < NT; ++nt2) acc[mt][nt2][0] = acc[mt][nt2][1] = 0.0;
            const double *ya = buf + g * YS + j;
            const double *mb = sMr + g * YS + j;
#pragma unroll kGemmUnroll
            for (int u = 0; u < KP / 4; ++u) {
                double a[kMT], b[NT];
#pragma unroll
                for (int mt = 0; mt < kMT; ++mt) a[mt] = ya[mt * 8 * YS + u * 4];
#pragma unroll
                for (int nt2 = 0; nt2 < NT; ++nt2) b[nt2] = mb[nt2 * 8 * YS + u * 4];
#pragma unroll
                for (int nt2 = 0; nt2 < NT; ++nt2)
#pragma unroll
                    for (int mt = 0; mt < kMT; ++mt) dmma(acc[mt][nt2][0], acc[mt][nt2][1], a[mt], b[nt2]);
            }
            __syncwarp();
            if (P.prof) tp4 = clock64();
            // accumulator tile (mt, nt): lane holds z[individual mt*8 + g][coordinate nt*8 + 2j + {0,1}] -> zT
#pragma unroll
            for (int mt = 0; mt < kMT; ++mt)
#pragma unroll
                for (int nt2 = 0; nt2 < NT; ++nt2) {
                    const int c0 = nt2 * 8 + 2 * j;
                    double z0 = acc[mt][nt2][0], z1 = acc[mt][nt2][1];
                    if (!unit_rate) { // the primitive's own sh_rate inside a hybrid (:381 etc. with s_flag = r_flag = 0)
                        z0 *= sRate[c0];
                        z1 *= sRate[c0 + 1];
                    }
                    buf[zt_index(c0, mt * 8 + g)] = z0;
                    buf[zt_index(c0 + 1, mt * 8 + g)] = z1;
                }
            __syncwarp();
        }

        if (!ROT && !unit_rate) { // un-rotated hybrid (not in the CEC2014 recipes; kept for completeness)
            for (int j = eq; j < D; j += kLPI) buf[zt_index(j, et)] *= sRate[j];
            __syncwarp();
        }
        if (P.prof) tp5 = clock64();
        // ---- E: primitives on z (kLPI lanes per individual) -------------------------------------------------------
        double val = 0.0;
        {
            Elem v;
            v.zt = buf;
            v.ind = et;
            v.idx = perm;
            for (int gi = 0; gi < P.st.ngroups; ++gi) {
                const GroupDesc &g = P.st.g[gi];
                v.off = g.off;
                val += eval_group(g, v, P.table, eq);
            }
        }
        if (eq == 0 && et < nt) {
            if (need_w) {
                if (P.st.scaled) val = P.st.mul * val / P.st.div; // e.g. :1047 fit = 10000 * fit / 1e+4
                P.out[t0 + et] = val;
                P.wout[t0 + et] = wacc;
            } else {
                P.out[t0 + et] = val + P.fbias; // :126 f[0] += 100.0 * func
            }
        }
        __syncwarp();
        if (P.prof && lane == 0) {
            tp6 = clock64();
            atomicAdd(P.prof + kPhLoad, static_cast<unsigned long long>(tp1 - tp0));
            atomicAdd(P.prof + kPhWeight, static_cast<unsigned long long>(tp2 - tp1));
            if (ROT) {
                atomicAdd(P.prof + kPhTokenWait, static_cast<unsigned long long>(tp3 - tp2));
                atomicAdd(P.prof + kPhGemm, static_cast<unsigned long long>(tp4 - tp3));
                atomicAdd(P.prof + kPhStoreZ, static_cast<unsigned long long>(tp5 - tp4));
            }
            atomicAdd(P.prof + kPhEpilogue, static_cast<unsigned long long>(tp6 - tp5));
            atomicAdd(P.prof + kPhTiles, 1ull);
        }
    }
}

// ---- separable fast path --------------------------------------------------------------------------------------------
// Un-rotated, un-permuted single-group stages whose primitive is a plain sum over coordinates (f8 rastrigin, f10 and
// cf02[0] schwefel, cf01[4] ellips): nothing needs the tile in shared memory.  G lanes share a row (individual): lane
// cl of the group takes the 16-byte chunks cl, cl + G, ... of the row straight from global memory, evaluates its
// coordinates in registers and the G partial sums are added in lane order - the same order for every row, so a
// result does not depend on the row's position in the batch.  HBM / FP64-instruction bound, no barriers.
__host__ __device__ constexpr int sep_lanes(int d)
{
    // lanes per row: maximise (used lanes / 32) x (used chunk slots / issued chunk slots) with at most 8 chunks per
    // lane (they are all in flight at once, in registers); ties -> wider group
    const int ch = d / 2;
    int best = 32;
    long long best_num = 0, best_den = 1;
    for (int g = 1; g <= 32; ++g) {
        const int rows = 32 / g, pass = (ch + g - 1) / g;
        if (pass > 8) continue;
        const long long num = static_cast<long long>(rows) * g * ch, den = 32LL * g * pass;
        if (num * best_den >= best_num * den) {
            best = g;
            best_num = num;
            best_den = den;
        }
    }
    return best;
}

constexpr int kSepThreads = 256;

#ifndef PGC_SEP_MINB
#define PGC_SEP_MINB 2
#endif
template <int D> __global__ void __launch_bounds__(kSepThreads, PGC_SEP_MINB) cec14_sep_kernel(const __grid_constant__ StageParams P)
{
    static_assert(D % 2 == 0, "rows are fetched in 16-byte chunks");
    constexpr int CH = D / 2, G = sep_lanes(D), ROWS = 32 / G, PASS = (CH + G - 1) / G;
    const int lane = threadIdx.x & 31;
    const int r = lane / G, cl = lane - r * G;
    const bool lane_ok = r < ROWS;
    const GroupDesc &g = P.st.g[0];
    const double *__restrict__ gt = P.table + g.tab_off;
    const bool need_w = P.wout != nullptr;
    const double pre_rate = P.st.pre_rate, rate = g.rate;
    const double inv_n = 1.0 / static_cast<double>(D);
    const int prim = g.prim;

    double2 osv[PASS], cf[PASS];
#pragma unroll
    for (int ps = 0; ps < PASS; ++ps) {
        const int c = ps * G + cl;
        const bool ok = lane_ok && c < CH;
        osv[ps] = ok ? make_double2(P.os[2 * c], P.os[2 * c + 1]) : make_double2(0.0, 0.0);
        cf[ps] = (ok && prim == P_ELLIPS) ? make_double2(gt[2 * c], gt[2 * c + 1]) : make_double2(0.0, 0.0);
    }

    const long long ntiles = (P.n + ROWS - 1) / ROWS;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    // the loads of tile t + 1 are issued before tile t is evaluated (register double buffer): ncu showed the HBM round trip of a
    // tile exposed with only ~4 warps per scheduler (long-scoreboard stalls 2.2 per issued instruction)
    auto fetch = [&](long long tile, double2(&dst)[PASS]) {
        const long long row = tile * ROWS + r;
        const bool act = lane_ok && tile < ntiles && row < P.n;
        const double2 *src = reinterpret_cast<const double2 *>(P.x + (act ? row : 0) * D);
#pragma unroll
        for (int ps = 0; ps < PASS; ++ps) {
            const int c = ps * G + cl;
            dst[ps] = (act && c < CH) ? __ldcs(src + c) : make_double2(0.0, 0.0);
        }
    };
    const long long tile0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    double2 xnext[PASS];
    fetch(tile0, xnext);
    for (long long tile = tile0; tile < ntiles; tile += nwarps) {
        const long long row = tile * ROWS + r;
        const bool active = lane_ok && row < P.n;
        double2 xv[PASS];
#pragma unroll
        for (int ps = 0; ps < PASS; ++ps) xv[ps] = xnext[ps];
        fetch(tile + nwarps, xnext);
        double s = 0.0, w = 0.0;
        constexpr bool FULL = CH % G == 0; // every pass is complete: no per-pass branch
        double z0[PASS], z1[PASS];
#pragma unroll
        for (int ps = 0; ps < PASS; ++ps) {
            const int c = ps * G + cl;
            const bool ok = FULL || c < CH;
            const double d0 = xv[ps].x - osv[ps].x, d1 = xv[ps].y - osv[ps].y; // :1245-1258
            if (ok) w += d0 * d0 + d1 * d1;                                      // cf_cal weight, :1330-1332
            z0[ps] = d0 * pre_rate * rate;
            z1[ps] = d1 * pre_rate * rate;
        }
        // the primitive is chosen OUTSIDE the pass loop: each branch is one straight-line block over all of the lane's coordinates,
        // so their (independent) evaluations interleave - with the switch inside the loop every pass was its own basic block and
        // the kernel ran at the latency of one FP64 dependency chain (ncu: FP64 pipe 56-60 % busy, issue slots 47-61 %)
        if (prim == P_RASTRIGIN) {
            double ta[PASS], tb[PASS];
#pragma unroll
            for (int ps = 0; ps < PASS; ++ps) {
                ta[ps] = rastrigin_term<true>(z0[ps]);
                tb[ps] = rastrigin_term<true>(z1[ps]);
            }
#pragma unroll
            for (int ps = 0; ps < PASS; ++ps)
                if (FULL || ps * G + cl < CH) {
                    s += ta[ps];
                    s += tb[ps];
                }
        } else if (prim == P_SCHWEFEL) {
            double s0[PASS], p0[PASS], s1[PASS], p1[PASS];
#pragma unroll
            for (int ps = 0; ps < PASS; ++ps) {
                schwefel_term<true>(z0[ps], inv_n, s0[ps], p0[ps]);
                schwefel_term<true>(z1[ps], inv_n, s1[ps], p1[ps]);
            }
#pragma unroll
            for (int ps = 0; ps < PASS; ++ps)
                if (FULL || ps * G + cl < CH) {
                    s -= s0[ps];
                    s += p0[ps];
                    s -= s1[ps];
                    s += p1[ps];
                }
        } else { // P_ELLIPS, :382-384
#pragma unroll
            for (int ps = 0; ps < PASS; ++ps)
                if (FULL || ps * G + cl < CH) {
                    s += cf[ps].x * z0[ps] * z0[ps];
                    s += cf[ps].y * z1[ps] * z1[ps];
                }
        }
        // the row's G partial sums, added in lane order
        double val = 0.0, wsum = 0.0;
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const int from = (lane_ok ? r * G : 0) + k;
            val += __shfl_sync(kFull, s, from);
            if (need_w) wsum += __shfl_sync(kFull, w, from);
        }
        if (prim == P_SCHWEFEL) val = val + g.c0;
        if (active && cl == 0) {
            if (need_w) {
                if (P.st.scaled) val = P.st.mul * val / P.st.div;
                P.out[row] = val;
                P.wout[row] = wsum;
            } else {
                P.out[row] = val + P.fbias;
            }
        }
    }
}

template <int D> int launch_sep(pgc_ctx *ctx, const StageParams &sp, cudaStream_t stream)
{
    constexpr int ROWS = 32 / sep_lanes(D);
    auto kern = cec14_sep_kernel<D>;
    int per_sm = 1;
    PGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSepThreads, 0));
    if (per_sm < 1) per_sm = 1;
    const long long ntiles = (sp.n + ROWS - 1) / ROWS;
    long long ctas = (ntiles + kSepThreads / 32 - 1) / (kSepThreads / 32);
    const long long cap = static_cast<long long>(ctx->sm_count) * per_sm;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    kern<<<static_cast<unsigned>(ctas), kSepThreads, 0, stream>>>(sp);
    PGC_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

inline bool stage_is_separable(const StageDesc &st, int dim)
{
    if (st.rotate || st.permute || st.ngroups != 1) return false;
    const GroupDesc &g = st.g[0];
    if (g.off != 0 || g.len != dim) return false;
    return g.prim == P_ELLIPS || g.prim == P_RASTRIGIN || g.prim == P_SCHWEFEL;
}

struct CombineParams {
    const double *fit; // [nstages][n]
    const double *w;   // [nstages][n]
    double *out;
    long long n;
    int nstages;
    int dim;
    double fbias;
    double delta[kMaxStages];
    double cbias[kMaxStages];
};

// cf_cal, cec2014.cpp:1319-1353
__global__ void cec14_combine_kernel(const __grid_constant__ CombineParams P)
{
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    double w[kMaxStages], fit[kMaxStages];
    double w_max = 0.0, w_sum = 0.0;
    const double nx = static_cast<double>(P.dim);
    for (int s = 0; s < P.nstages; ++s) {
        fit[s] = P.fit[s * P.n + i] + P.cbias[s];
        double ws = P.w[s * P.n + i];
        if (ws != 0.0)
            ws = sqrt(1.0 / ws) * exp(-ws / 2.0 / nx / (P.delta[s] * P.delta[s]));
        else
            ws = DBL_MAX;
        if (ws > w_max) w_max = ws;
        w[s] = ws;
    }
    for (int s = 0; s < P.nstages; ++s) w_sum = w_sum + w[s];
    if (w_max == 0.0) {
        for (int s = 0; s < P.nstages; ++s) w[s] = 1.0;
        w_sum = P.nstages;
    }
    double f = 0.0;
    for (int s = 0; s < P.nstages; ++s) f = f + w[s] / w_sum * fit[s];
    P.out[i] = f + P.fbias;
}

template <int D, bool ROT> size_t stage_smem_bytes()
{
    constexpr int DP = pad8(D);
    constexpr int YS = ystride(D);
    return sizeof(double) * ((ROT ? DP * YS : 0) + kWarps * warp_buf_elems(D) + D + DP) + sizeof(int) * D + 16;
}

template <int D, bool ROT> int launch_stage(pgc_ctx *ctx, const StageParams &sp, cudaStream_t stream)
{
    static thread_local int configured_dev = -1;
    const size_t smem = stage_smem_bytes<D, ROT>();
    auto kern = cec14_stage_kernel<D, ROT>;
    if (configured_dev != ctx->device) {
        PGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured_dev = ctx->device;
    }
    const long long ntiles = (sp.n + kTileInd - 1) / kTileInd;
    long long ctas = (ntiles + kWarps - 1) / kWarps;
    int per_sm = 1;
    if (!ROT) { // no big matrix: several CTAs fit per SM, ask the runtime how many
        PGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarps * 32, smem));
        if (per_sm < 1) per_sm = 1;
    }
    const long long cap = static_cast<long long>(ctx->sm_count) * per_sm;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    kern<<<static_cast<unsigned>(ctas), kWarps * 32, smem, stream>>>(sp);
    PGC_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

template <int D> int launch_stage_d(pgc_ctx *ctx, const StageParams &sp, bool rot, cudaStream_t stream)
{
    return rot ? launch_stage<D, true>(ctx, sp, stream) : launch_stage<D, false>(ctx, sp, stream);
}

// Shared-memory image of a row-major D x D rotation (Mr[i*D + j], z_i = sum_j Mr[i][j] y_j): DP rows (outputs,
// zero rows beyond D) of stride ystride(D) (zero beyond D), copied verbatim by the kernel.
size_t rotation_image_elems(int D) { return static_cast<size_t>(pad8(D)) * ystride(D); }

void make_rotation_image(const double *mr, int D, double *dst)
{
    const int YS = ystride(D);
    std::memset(dst, 0, sizeof(double) * rotation_image_elems(D));
    for (int i = 0; i < D; ++i)
        for (int k = 0; k < D; ++k) dst[static_cast<size_t>(i) * YS + k] = mr[static_cast<size_t>(i) * D + k];
}

} // namespace

int cec2014_create(pgc_problem *p, const pgc_problem_desc *d)
{
    Cec2014Recipe &r = p->cec14;
    int rc = build_cec2014_recipe(d->prob_id, d->dim, r);
    if (rc != PGC_OK) return rc;
    const size_t D = d->dim;
    int ncomp = 0;
    bool any_rot = false, any_perm = false;
    for (int s = 0; s < r.nstages; ++s) {
        if (r.st[s].comp + 1 > ncomp) ncomp = r.st[s].comp + 1;
        any_rot |= r.st[s].rotate != 0;
        any_perm |= r.st[s].permute != 0;
    }
    PGC_REQUIRE(d->shift && d->shift_len >= ncomp * D,
                "cec2014: shift table needs at least %zu values (m_origin_shift layout), got %zu", ncomp * D,
                d->shift_len);
    PGC_REQUIRE(!any_rot || (d->rotation && d->rotation_len >= ncomp * D * D),
                "cec2014: rotation table needs at least %zu values, got %zu", ncomp * D * D, d->rotation_len);
    PGC_REQUIRE(!any_perm || (d->shuffle && d->shuffle_len >= ncomp * D),
                "cec2014: shuffle table needs at least %zu values, got %zu", ncomp * D, d->shuffle_len);

    p->nx = D;
    p->nobj = 1;
    p->lb.assign(D, -100.0); // cec2014.cpp:103-109
    p->ub.assign(D, 100.0);
    static const char *names[31] = {"", "ellips_func", "bent_cigar_func", "discus_func", "rosenbrock_func", "ackley_func",
                                    "weierstrass_func", "griewank_func", "rastrigin_func_non_rotated", "rastrigin_func",
                                    "schwefel_func_non_rotated", "schwefel_func", "katsuura_func", "happycat_func",
                                    "hgbat_func", "grie_rosen_func", "escaffer6_func", "hf01", "hf02", "hf03", "hf04",
                                    "hf05", "hf06", "cf01", "cf02", "cf03", "cf04", "cf05", "cf06", "cf07", "cf08"};
    p->name = "CEC2014 - f" + std::to_string(d->prob_id) + "(" + names[d->prob_id] + ")"; // :253-350
    p->flops_per_eval = r.flops_per_eval;
    p->transc_per_eval = r.transc_per_eval;

    PGC_CUDA(cudaSetDevice(p->ctx->device));
    PGC_CUDA(cudaMalloc(&p->d_shift, sizeof(double) * ncomp * D));
    PGC_CUDA(cudaMemcpy(p->d_shift, d->shift, sizeof(double) * ncomp * D, cudaMemcpyHostToDevice));
    if (any_rot) {
        const size_t img = rotation_image_elems(static_cast<int>(D));
        std::vector<double> tiled(static_cast<size_t>(ncomp) * img);
        for (int c = 0; c < ncomp; ++c) make_rotation_image(d->rotation + c * D * D, static_cast<int>(D), tiled.data() + c * img);
        PGC_CUDA(cudaMalloc(&p->d_rotation, sizeof(double) * tiled.size()));
        PGC_CUDA(cudaMemcpy(p->d_rotation, tiled.data(), sizeof(double) * tiled.size(), cudaMemcpyHostToDevice));
    }
    if (any_perm) {
        std::vector<int> zero_based(ncomp * D);
        for (size_t i = 0; i < zero_based.size(); ++i) {
            const int s = d->shuffle[i];
            PGC_REQUIRE(s >= 1 && s <= static_cast<int>(D), "cec2014: shuffle entry %zu = %d is not in [1, %zu]", i, s, D);
            zero_based[i] = s - 1; // :808  m_z[S[j] - 1]
        }
        PGC_CUDA(cudaMalloc(&p->d_shuffle, sizeof(int) * zero_based.size()));
        PGC_CUDA(cudaMemcpy(p->d_shuffle, zero_based.data(), sizeof(int) * zero_based.size(), cudaMemcpyHostToDevice));
    }
    const size_t tab = r.table.size() ? r.table.size() : 2;
    PGC_CUDA(cudaMalloc(&p->d_table, sizeof(double) * tab));
    if (r.table.size())
        PGC_CUDA(cudaMemcpy(p->d_table, r.table.data(), sizeof(double) * r.table.size(), cudaMemcpyHostToDevice));
    return PGC_OK;
}

void cec2014_destroy(pgc_problem *p)
{
    cudaFree(p->d_shift);
    cudaFree(p->d_rotation);
    cudaFree(p->d_shuffle);
    cudaFree(p->d_table);
    p->d_shift = p->d_rotation = p->d_table = nullptr;
    p->d_shuffle = nullptr;
}

int cec2014_eval_impl(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream,
                      unsigned long long *d_prof);

int cec2014_eval(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream)
{
    return cec2014_eval_impl(p, d_dvs, n, d_fvs, stream, nullptr);
}

// Debug aid: run the evaluation with per-phase cycle counters (summed over warp-tiles and stages).
// out[0..6] = load, weight pass, token wait, GEMM, z store, epilogue cycles, warp-tiles.
int cec2014_phase_cycles(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, unsigned long long *out)
{
    unsigned long long *d_prof = nullptr;
    PGC_CUDA(cudaMalloc(&d_prof, sizeof(unsigned long long) * kPhCount));
    PGC_CUDA(cudaMemset(d_prof, 0, sizeof(unsigned long long) * kPhCount));
    int rc = cec2014_eval_impl(p, d_dvs, n, d_fvs, p->ctx->stream, d_prof);
    if (rc == PGC_OK) {
        cudaError_t e = cudaStreamSynchronize(p->ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpy(out, d_prof, sizeof(unsigned long long) * kPhCount, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "phase profile", __FILE__, __LINE__);
    }
    cudaFree(d_prof);
    return rc;
}

int cec2014_eval_impl(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream,
                      unsigned long long *d_prof)
{
    if (n == 0) return PGC_OK;
    const Cec2014Recipe &r = p->cec14;
    pgc_ctx *ctx = p->ctx;
    const size_t D = r.dim, img = rotation_image_elems(r.dim);
    double *fit = nullptr, *w = nullptr;
    if (r.composition) {
        int rc = ensure_scratch(ctx, sizeof(double) * 2 * kMaxStages * n);
        if (rc != PGC_OK) return rc;
        fit = static_cast<double *>(ctx->scratch);
        w = fit + kMaxStages * n;
    }
    for (int s = 0; s < r.nstages; ++s) {
        StageParams sp;
        sp.x = d_dvs;
        sp.mr = r.st[s].rotate ? p->d_rotation + r.st[s].comp * img : nullptr;
        sp.os = p->d_shift + r.st[s].comp * D;
        sp.perm = r.st[s].permute ? p->d_shuffle + r.st[s].comp * D : nullptr;
        sp.table = p->d_table;
        sp.out = r.composition ? fit + s * n : d_fvs;
        sp.wout = r.composition ? w + s * n : nullptr;
        sp.n = static_cast<long long>(n);
        sp.fbias = r.fbias;
        sp.aligned16 = (reinterpret_cast<uintptr_t>(d_dvs) & 15u) == 0;
        sp.prof = d_prof;
        sp.st = r.st[s];
        int rc;
        const bool rot = r.st[s].rotate != 0;
        if (!d_prof && sp.aligned16 && stage_is_separable(r.st[s], r.dim)) {
            switch (r.dim) {
                case 2: rc = launch_sep<2>(ctx, sp, stream); break;
                case 10: rc = launch_sep<10>(ctx, sp, stream); break;
                case 20: rc = launch_sep<20>(ctx, sp, stream); break;
                case 30: rc = launch_sep<30>(ctx, sp, stream); break;
                case 50: rc = launch_sep<50>(ctx, sp, stream); break;
                case 100: rc = launch_sep<100>(ctx, sp, stream); break;
                default: set_error("cec2014: unsupported dimension %d", r.dim); return PGC_ERR_INVALID_ARGUMENT;
            }
            if (rc != PGC_OK) return rc;
            continue;
        }
        switch (r.dim) {
            case 2: rc = launch_stage_d<2>(ctx, sp, rot, stream); break;
            case 10: rc = launch_stage_d<10>(ctx, sp, rot, stream); break;
            case 20: rc = launch_stage_d<20>(ctx, sp, rot, stream); break;
            case 30: rc = launch_stage_d<30>(ctx, sp, rot, stream); break;
            case 50: rc = launch_stage_d<50>(ctx, sp, rot, stream); break;
            case 100: rc = launch_stage_d<100>(ctx, sp, rot, stream); break;
            default: set_error("cec2014: unsupported dimension %d", r.dim); return PGC_ERR_INVALID_ARGUMENT;
        }
        if (rc != PGC_OK) return rc;
    }
    if (r.composition) {
        CombineParams cp;
        cp.fit = fit;
        cp.w = w;
        cp.out = d_fvs;
        cp.n = static_cast<long long>(n);
        cp.nstages = r.nstages;
        cp.dim = r.dim;
        cp.fbias = r.fbias;
        for (int s = 0; s < kMaxStages; ++s) {
            cp.delta[s] = r.delta[s];
            cp.cbias[s] = r.cbias[s];
        }
        const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
        cec14_combine_kernel<<<blocks, 256, 0, stream>>>(cp);
        PGC_CUDA(cudaGetLastError());
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
    }
    return PGC_OK;
}

} // namespace pgc
