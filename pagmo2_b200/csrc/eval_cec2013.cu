// eval_cec2013.cu - CEC2013 f1..f28 batch fitness on sm_100a.
//
// Replaces, for a whole batch, reference pagmo::cec2013::fitness (src/problems/cec2013.cpp:78-197; primitives :319-865,
// compositions :867-1036, helpers :1038-1126).
//
// Every CEC2013 primitive is a chain over the reference's two work vectors (m_y, m_z): shift, then up to three rotations
// with element-wise maps in between (oszfunc :1061, asyfunc :1053, coordinate-dependent scalings), then a reduction.  The host
// turns the chain into a PLAN of launches, cut so that each launch needs exactly one rotation matrix in shared memory:
//   launch = load (x - Os, or the state saved by the previous launch) -> element-wise steps -> rotation -> element-wise
//            steps -> save the state [n x D] for the next launch, or reduce to f.
// A warp owns tiles of 8 individuals and two private shared-memory buffers (the images of m_y and m_z, row-major, padded
// stride), so the reference's "destination keeps its old value" semantics of asyfunc carry over literally.  The rotation runs
// on the FP64 tensor path exactly as in eval_cec2014.cu (mma.sync.m8n8k4.f64, matrix image resident in shared memory);
// reductions use 4 lanes per individual.  Compositions run their components one after the other (fit_i and the cf_cal
// distance sum_j (x_j - Os_ij)^2 go to scratch) and finish with one cf_cal kernel.
// Warps per CTA are chosen per launch: as many as fit next to the matrix (<= 16), fewer when the batch is small so that the
// tiles spread over all SMs (island-sized populations of ~1000 individuals are latency-, not throughput-bound).
//
// Reference quirks that are kept (see also oracle/restate_cec2013.c): asyfunc writes only positive inputs; grie_rosen's
// rotation is dead code (:812-820 overwrite m_z from the un-rotated m_y), so no rotation is launched for it; component i of a
// composition reads its shift at Os[i*nx] and its second rotation at Mr[(i+1)*nx*nx]; cf_cal's zero-distance weight is 1e99.
#include <cfloat>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cec_device.cuh"
#include "pgc_internal.cuh"

namespace pgc
{

enum Op13 : int { O_MULDIV, O_MUL, O_ROT, O_COPY, O_OSZ, O_ASY, O_TAB, O_ADD, O_STEP, O_BISIGN };
enum Red13 : int { R_SPHERE, R_ELLIPS, R_BENT, R_DISCUS, R_DIFPOW, R_ROSEN, R_SCHAF7, R_ACKLEY, R_WEIER, R_GRIEW, R_RASTR, R_SCHWEF,
                   R_KATS, R_BIRAS, R_GRROS, R_ESCAF, R_STEP_RASTR /* program id only; reduces like R_RASTR */ };

struct Step13 {
    int op, src, dst, tab; // tab: offset into the constant table (O_TAB multiplier, O_ASY coefficient)
    double a, b;
};
constexpr int kMaxSteps13 = 8;

struct Launch13 {
    int from_x;  // 1: buffer[in_buf] = x - Os_comp; 0: buffer[in_buf] = saved state
    int in_buf;
    int nsteps;
    int rot;     // index of the matrix used by this launch's O_ROT step, -1: none
    int to_state; // 1: save buffer[end_buf] for the next launch; 0: reduce buffer[end_buf]
    int end_buf;
    int red;
    int comp;    // shift component
    int want_w;  // composition: also emit the cf_cal distance of this component
    int slot;    // composition slot (fit / w row) or -1: write f + fbias
    int scaled;
    double mul, div;
    Step13 st[kMaxSteps13];
};

struct Cec2013Plan {
    int func = 0, dim = 0, ncomp = 1;
    double fbias = 0;
    double delta[5] = {}, cbias[5] = {};
    std::vector<Launch13> launches;
    // constant-table offsets
    int t10 = 0, t100 = 0, tell = 0, tdif = 0, tgri = 0, twei = 0, tasy5 = 0, tasy2 = 0;
    double weier_c0 = 0, kats_c0 = 0, kats_c1 = 0, bi_s = 0, bi_mu1 = 0;
    bool needs_state = false;
};

namespace
{

using namespace cecdev;
constexpr int kMaxWarps13 = 16;
// island-sized batches: several islands share the GPU (one stream each), so the SMALL instance is built to co-reside - at most 8 warps
// per CTA and a register budget that lets kSmallMinBlocks13 CTAs of different islands sit on one SM
#ifndef PGC_CEC13_SMALL_MINBLOCKS
#define PGC_CEC13_SMALL_MINBLOCKS 3
#endif
constexpr int kSmallWarps13 = 8, kSmallMinBlocks13 = PGC_CEC13_SMALL_MINBLOCKS;

constexpr int kMaxMerge = 3; // sub-launches of one kernel launch (island-sized batches: all matrices of a chain in shared memory)

struct Params13 {
    const double *x;
    double *state;
    const double *mr[kMaxMerge]; // rotation image of every sub-launch
    const double *os;    // component shift
    const double *table;
    double *out;
    double *wout;
    long long n;
    double fbias;
    double weier_c0, kats_c0, kats_c1, bi_s, bi_mu1;
    int tell, tdif, tgri, twei;
    int strict; // rotations accumulate one product at a time in the reference's j order (pgc_problem_set_strict)
    int ti;     // individuals per warp tile: 8, or 4 / 2 / 1 for island-sized batches (the other rows of the tile stay zero)
    int nl;     // sub-launches executed back to back on a tile (the vector stays in the warp's buffers between them)
    Launch13 L[kMaxMerge];
};

struct Row { // one individual's row of a buffer
    const double *p;
    __device__ __forceinline__ double operator()(int j) const { return p[j]; }
};

// reductions of cec2013.cpp on the finished vector v (4 lanes per individual, lane q takes terms j == q mod 4)
template <int D>
__device__ double reduce13(const Params13 &P, const Launch13 &L, const Row &v, int h, const double *__restrict__ xrow,
                           const double *__restrict__ sOs)
{
    constexpr int n = D;
    const double dn = static_cast<double>(n);
    const double two_pi = 2.0 * 3.141592653589793238462643383279502884;
    const double *tab = P.table;
    switch (L.red) {
        case R_SPHERE: // :329-331
            return pair_add(ordered_sum(h, n, [&](int j) { return v(j) * v(j); }));
        case R_ELLIPS: // :345-348
            return pair_add(ordered_sum(h, n, [&](int j) { return tab[P.tell + j] * v(j) * v(j); }));
        case R_BENT: // :369-372
            return pair_add(ordered_sum(h, n, [&](int j) { return (j == 0) ? v(j) * v(j) : 1.0e6 * v(j) * v(j); }));
        case R_DISCUS: // :387-390
            return pair_add(ordered_sum(h, n, [&](int j) { return (j == 0) ? 1.0e6 * v(j) * v(j) : v(j) * v(j); }));
        case R_DIFPOW: // :403-407
            return sqrt(pair_add(ordered_sum(h, n, [&](int j) { return pow(fabs(v(j)), tab[P.tdif + j]); })));
        case R_ROSEN: // :430-435 (the +1 was applied by the chain)
            return pair_add(ordered_sum(h, n - 1, [&](int j) {
                const double t1 = v(j) * v(j) - v(j + 1), t2 = v(j) - 1.0;
                return 100.0 * t1 * t1 + t2 * t2;
            }));
        case R_SCHAF7: { // :458-465
            const double s = pair_add(ordered_sum(h, n - 1, [&](int j) {
                const double w = sqrt(v(j) * v(j) + v(j + 1) * v(j + 1));
                const double t = sin(50.0 * pow(w, 0.2));
                const double r = sqrt(w);
                return r + r * t * t;
            }));
            const double nm1 = static_cast<double>(n - 1);
            return s * s / nm1 / nm1;
        }
        case R_ACKLEY: { // :490-499
            double s1 = ordered_sum(h, n, [&](int j) { return v(j) * v(j); });
            double s2 = ordered_sum(h, n, [&](int j) { return cos(two_pi * v(j)); });
            s1 = pair_add(s1);
            s2 = pair_add(s2);
            s1 = -0.2 * sqrt(s1 / dn);
            s2 /= dn;
            return 2.718281828459045235360287471352662498 - 20.0 * exp(s1) - exp(s2) + 20.0;
        }
        case R_WEIER: { // :528-541; same angle-tripling evaluation as eval_cec2014.cu (terms k = 0, 10 restart exactly)
            const double *gt = tab + P.twei;
            return pair_add(ordered_sum(h, n, [&](int j) {
                       const double u = v(j) + 0.5;
                       double sum = 0.0, w = 1.0, sn = 0.0, cs = 1.0;
#pragma unroll
                       for (int k = 0; k <= 20; ++k) {
                           if (k % 10 == 0 && k < 20) {
                               sincos_turns(turns_of(gt[k] * u), sn, cs);
                           } else {
                               triple_angle(sn, cs);
                           }
                           sum = fma(w, cs, sum);
                           w *= 0.5;
                       }
                       return sum;
                   }))
                   - P.weier_c0;
        }
        case R_GRIEW: { // :564-571; table = 1/sqrt(1+j)
            double s = ordered_sum(h, n, [&](int j) { return v(j) * v(j); });
            double p = 1.0;
            for (int j = h; j < n; j += kLPI) p *= cos_theta(v(j) * tab[P.tgri + j]);
            s = pair_add(s);
            p = pair_mul(p);
            return 1.0 + s / 4000.0 - p;
        }
        case R_RASTR: // :611-614, :657-660
            return pair_add(ordered_sum(h, n, [&](int j) { return (v(j) * v(j) - 10.0 * cos_theta(two_pi * v(j)) + 10.0); }));
        case R_SCHWEF: { // :685-699 (the +420.96... was applied by the chain); branches folded as in eval_cec2014.cu
            const double inv_n = 1.0 / dn;
            double s = 0.0;
            for (int j = h; j < n; j += kLPI) {
                const double z = v(j);
                const double az = fabs(z);
                const bool big = az > 500.0;
                const double q = round_magic(az * 0.002);
                double fm = fma(-q, 500.0, az);
                fm = (fm < 0.0) ? fm + 500.0 : fm;
                const double m = big ? 500.0 - fm : az;
                const double t = (z - copysign(500.0, z)) * 0.01;
                s -= copysign(m, z) * sin_theta(sqrt(m));
                s += big ? t * t * inv_n : 0.0;
            }
            return 4.189828872724338e+002 * dn + pair_add(s);
        }
        case R_KATS: { // :726-738; prod_j b_j^c0 = exp(c0 * log(prod_j b_j)): every lane multiplies its own factors (each in
                       // [1, 1 + n/2], at most 25 of them) and takes one log - as eval_cec2014.cu
            const double prod = ordered_prod(h, n, [&](int j) { return 1.0 + static_cast<double>(j + 1) * katsuura_inner(v(j)); });
            double slog = pair_add(log(prod));
            return exp(P.kats_c0 * slog) * P.kats_c1 - P.kats_c1;
        }
        case R_BIRAS: { // :779-797; tmpx is rebuilt from x (:751-763): y = (x-Os)*0.1, tmpx = +-2y + mu0
            const double mu0 = 2.5;
            double t1 = 0.0, t2 = 0.0, tc = 0.0;
            for (int j = h; j < n; j += kLPI) {
                double t = 2.0 * ((xrow[j] - sOs[j]) * (10.0 / 100.0));
                if (sOs[j] < 0.0) t *= -1.0;
                t += mu0;
                const double q1 = t - mu0, q2 = t - P.bi_mu1;
                t1 += q1 * q1;
                t2 += q2 * q2;
                tc += cos_theta(two_pi * v(j));
            }
            t1 = pair_add(t1);
            t2 = pair_add(t2);
            tc = pair_add(tc);
            t2 *= P.bi_s;
            t2 += 1.0 * dn;
            const double f = (t1 < t2) ? t1 : t2;
            return f + 10.0 * (dn - tc);
        }
        case R_GRROS: // :822-832 (cyclic last term)
            return pair_add(ordered_sum(h, n, [&](int j) {
                const int jn = (j + 1 == n) ? 0 : j + 1;
                const double t1 = v(j) * v(j) - v(jn), t2 = v(j) - 1.0;
                const double temp = 100.0 * t1 * t1 + t2 * t2;
                return (temp * temp) / 4000.0 - cos(temp) + 1.0;
            }));
        case R_ESCAF: // :853-864 (cyclic last term)
            return pair_add(ordered_sum(h, n, [&](int j) {
                const int jn = (j + 1 == n) ? 0 : j + 1;
                const double ss = v(j) * v(j) + v(jn) * v(jn);
                double t1 = sin(sqrt(ss));
                t1 = t1 * t1;
                const double t2 = 1.0 + 0.001 * ss;
                return 0.5 + (t1 - 0.5) / (t2 * t2);
            }));
        default: return 0.0;
    }
}

// SMALL = island-sized batches: run-time tile size (P.ti individuals per warp) and up to kMaxMerge sub-launches per launch; the
// throughput instance keeps both as compile-time constants (8 individuals, one sub-launch)
template <int D, bool SMALL> __global__ void __launch_bounds__(SMALL ? kSmallWarps13 * 32 : kMaxWarps13 * 32, SMALL ? kSmallMinBlocks13 : 1) cec13_kernel(const __grid_constant__ Params13 P)
{
    constexpr int DP = pad8(D), KP = pad4(D), NT = DP / 8, YS = ystride(D);
    // island-sized batches: fewer individuals per warp tile, so that the batch spreads over many warps - every phase of a tile is
    // a dependent chain of one warp, and a row of the tile is computed independently of the other rows (same bits either way)
    const int TI = SMALL ? P.ti : kTileInd, TILE = TI * D;
    const int NL = SMALL ? P.nl : 1;
    bool rot = false; // any sub-launch rotates
    for (int q = 0; q < NL; ++q) rot = rot || P.L[q].rot >= 0;
    const int mslots = rot ? NL : 0; // one matrix slot per sub-launch
    const int W = blockDim.x >> 5;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sMr = reinterpret_cast<double *>(smem_raw);
    double *sBuf = sMr + mslots * DP * YS;
    double *sOs = sBuf + W * 2 * TI * YS; // a warp's two buffers hold TI rows
    unsigned short *sList = reinterpret_cast<unsigned short *>(sOs + D); // [W][TILE]: positions of the positive inputs of asyfunc

    for (int q = 0; q < NL; ++q) {
        if (P.L[q].rot < 0) continue;
        const double2 *src = reinterpret_cast<const double2 *>(P.mr[q]);
        double2 *dst = reinterpret_cast<double2 *>(sMr + q * DP * YS);
        for (int i = threadIdx.x; i < DP * YS / 2; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    for (int i = threadIdx.x; i < W * 2 * TI * YS; i += blockDim.x) sBuf[i] = 0.0; // the inner-index padding stays 0
    for (int i = threadIdx.x; i < D; i += blockDim.x) sOs[i] = P.os[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *bufs[2] = {sBuf + warp * 2 * TI * YS, sBuf + warp * 2 * TI * YS + TI * YS};
    unsigned short *list = sList + warp * TILE;
    const long long ntiles = (P.n + TI - 1) / TI;
    const int et = lane & (kTileInd - 1), eq = lane / kTileInd;
    const int er = et < TI ? et : 0; // lanes of absent rows read row 0 and drop their result
    const double *tab = P.table;

    if (rot && W > 4 && TI == kTileInd) { // spread the warps of an SM sub-partition over the phases (see eval_cec2014.cu)
        const long long wait = static_cast<long long>(warp >> 2) * (NT * (KP / 4) * 16);
        const long long t_start = clock64();
        while (clock64() - t_start < wait) {
        }
    }

    for (long long tile = static_cast<long long>(blockIdx.x) * W + warp; tile < ntiles; tile += static_cast<long long>(gridDim.x) * W) {
        const long long t0 = tile * TI;
        const int nt = (P.n - t0 < TI) ? static_cast<int>(P.n - t0) : TI;
        const int live = nt * D;
        double wacc = 0.0;
        int prev_end = 0;

        for (int q = 0; q < NL; ++q) {
        const Launch13 &L = P.L[q];
        const double *sM = sMr + q * DP * YS;
        { // ---- load: the 8 rows of a tile are one contiguous block of the input
            double *in = bufs[L.in_buf];
            if (q > 0 && !L.from_x) { // merged launch: the previous sub-launch left the vector in one of the warp's buffers
                if (prev_end != L.in_buf) {
                    const double *src = bufs[prev_end];
                    for (int e = lane; e < TILE; e += 32) {
                        const int t = e / D, a = t * YS + e - t * D;
                        in[a] = src[a];
                    }
                }
                __syncwarp();
            } else if (L.from_x) {
                const double *src = P.x + t0 * D;
                for (int e = lane; e < TILE; e += 32) {
                    const int t = e / D, i = e - t * D;
                    in[t * YS + i] = (e < live) ? __ldg(src + e) - sOs[i] : 0.0; // shiftfunc :1038-1044
                }
                __syncwarp();
                if (L.want_w) { // cf_cal distance sum_j (x_j - Os_j)^2, :1101-1103
                    const double *row = in + er * YS;
                    for (int j = eq; j < D; j += kLPI) wacc += row[j] * row[j];
                    wacc = pair_add(wacc);
                }
            } else {
                const double *src = P.state + t0 * D;
                for (int e = lane; e < TILE; e += 32) {
                    const int t = e / D, i = e - t * D;
                    in[t * YS + i] = (e < live) ? src[e] : 0.0;
                }
                __syncwarp();
            }
        }

        for (int si = 0; si < L.nsteps; ++si) {
            const Step13 &s = L.st[si];
            const double *in = bufs[s.src];
            double *out = bufs[s.dst];
            switch (s.op) {
                case O_MULDIV:
                    for (int e = lane; e < TILE; e += 32) {
                        const int t = e / D, a = t * YS + e - t * D;
                        out[a] = in[a] * s.a / s.b;
                    }
                    break;
                case O_MUL:
                    for (int e = lane; e < TILE; e += 32) {
                        const int t = e / D, a = t * YS + e - t * D;
                        out[a] = in[a] * s.a;
                    }
                    break;
                case O_COPY:
                    for (int e = lane; e < TILE; e += 32) {
                        const int t = e / D, a = t * YS + e - t * D;
                        out[a] = in[a];
                    }
                    break;
                case O_ADD:
                    for (int e = lane; e < TILE; e += 32) {
                        const int t = e / D, a = t * YS + e - t * D;
                        out[a] = in[a] + s.a;
                    }
                    break;
                case O_TAB: // m * pow(base, i/(nx-1)/2), table built on the host with the reference's expression
                    for (int e = lane; e < TILE; e += 32) {
                        const int t = e / D, i = e - t * D, a = t * YS + i;
                        out[a] = in[a] * tab[s.tab + i];
                    }
                    break;
                case O_STEP: // :633-635
                    for (int e = lane; e < TILE; e += 32) {
                        const int t = e / D, a = t * YS + e - t * D;
                        const double v = in[a];
                        if (fabs(v) > 0.5) out[a] = floor(2. * v + 0.5) / 2.;
                    }
                    break;
                case O_ASY: { // :1053-1059: only positive inputs are written
                    // about half of the inputs are positive; their positions are first compacted into a list (warp votes) so that
                    // the pow() below runs with full warps instead of half-empty ones
                    int cnt = 0;
                    for (int e0 = 0; e0 < TILE; e0 += 32) {
                        const int e = e0 + lane;
                        bool pos = false;
                        int a = 0, i = 0;
                        if (e < TILE) {
                            const int t = e / D;
                            i = e - t * D;
                            a = t * YS + i;
                            pos = in[a] > 0;
                        }
                        const unsigned m = __ballot_sync(kFull, pos);
                        if (pos) list[cnt + __popc(m & ((1u << lane) - 1u))] = static_cast<unsigned short>(a);
                        cnt += __popc(m);
                    }
                    __syncwarp();
                    for (int k = lane; k < cnt; k += 32) {
                        const int a = list[k], i = a - (a / YS) * YS;
                        const double v = in[a];
                        out[a] = pow(v, 1.0 + tab[s.tab + i] * sqrt(v));
                    }
                    break;
                }
                case O_OSZ: { // :1061-1089: end coordinates transformed, the rest copied
                    for (int e = lane; e < TILE; e += 32) {
                        const int t = e / D, i = e - t * D, a = t * YS + i;
                        if (i != 0 && i != D - 1) out[a] = in[a];
                    }
                    if (lane < 2 * kTileInd) {
                        const int t = lane & (kTileInd - 1), i = (lane < kTileInd) ? 0 : D - 1;
                        const double v = t < TI ? in[t * YS + i] : 0.0;
                        double r = 0.0; // v == 0: sx = 0 (the stale xx of the reference is finite, so 0 * exp(.) = 0)
                        if (v != 0) {
                            const double xx = log(fabs(v));
                            const double c1 = v > 0 ? 10 : 5.5, c2 = v > 0 ? 7.9 : 3.1;
                            r = exp(xx + 0.049 * (sin(c1 * xx) + sin(c2 * xx)));
                            r = v > 0 ? r : -r;
                        }
                        if (t < TI && (lane < kTileInd || D > 1)) out[t * YS + i] = r;
                    }
                    break;
                }
                case O_BISIGN: // :755-762: z = +-2y by the sign of the shift
                    for (int e = lane; e < TILE; e += 32) {
                        const int t = e / D, i = e - t * D, a = t * YS + i;
                        double v = 2 * in[a];
                        if (sOs[i] < 0.) v *= -1.;
                        out[a] = v;
                    }
                    break;
                case O_ROT: { // rotatefunc :1046-1051 on the FP64 tensor path (fragments as in eval_cec2014.cu)
                    __syncwarp();
                    if (P.strict) {
                        // the reference's own summation: xrot[i] = xrot[i] + x[j] * Mr[i * nx + j], j ascending, one rounding per
                        // multiply and per add (the library is built with -fmad=false) - bit-identical rotated vectors, for the
                        // functions whose later sin / cos / pow amplify the last bits (f7, f8, f20, f28)
                        for (int e = lane; e < TILE; e += 32) {
                            const int t = e / D, i = e - t * D;
                            const double *y = in + t * YS, *m = sM + i * YS;
                            double acc = 0.;
                            for (int k = 0; k < D; ++k) acc = acc + y[k] * m[k];
                            out[t * YS + i] = acc;
                        }
                        break;
                    }
                    const int g = lane >> 2, j = lane & 3;
                    double acc[NT][2];
#pragma unroll
                    for (int nt2 = 0; nt2 < NT; ++nt2) acc[nt2][0] = acc[nt2][1] = 0.0;
                    const double *ya = in + (g < TI ? g : 0) * YS + j; // absent rows: any finite operand, result dropped
                    const double *mb = sM + g * YS + j;
#pragma unroll 2
                    for (int u = 0; u < KP / 4; ++u) {
                        const double a = ya[u * 4];
                        double b[NT];
#pragma unroll
                        for (int nt2 = 0; nt2 < NT; ++nt2) b[nt2] = mb[nt2 * 8 * YS + u * 4];
#pragma unroll
                        for (int nt2 = 0; nt2 < NT; ++nt2) dmma(acc[nt2][0], acc[nt2][1], a, b[nt2]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int nt2 = 0; nt2 < NT; ++nt2) {
                        const int c = nt2 * 8 + 2 * j;
                        if (g < TI && c < D) out[g * YS + c] = acc[nt2][0];
                        if (g < TI && c + 1 < D) out[g * YS + c + 1] = acc[nt2][1];
                    }
                    break;
                }
                default: break;
            }
            __syncwarp();
        }

        if (L.to_state && q + 1 < NL) {
            prev_end = L.end_buf; // stays in shared memory for the next sub-launch
        } else if (L.to_state) {
            const double *src = bufs[L.end_buf];
            double *dst = P.state + t0 * D;
            for (int e = lane; e < live; e += 32) {
                const int t = e / D, i = e - t * D;
                dst[e] = src[t * YS + i];
            }
        } else {
            Row v{bufs[L.end_buf] + er * YS};
            const long long xr = (et < nt) ? t0 + et : t0;
            double val = reduce13<D>(P, L, v, eq, P.x + xr * D, sOs);
            if (eq == 0 && et < nt) {
                if (L.slot >= 0) {
                    if (L.scaled) val = L.mul * val / L.div; // e.g. :877 fit = 10000 * fit / 1e+4
                    P.out[L.slot * P.n + t0 + et] = val;
                } else {
                    P.out[t0 + et] = val + P.fbias; // :83 f[0] += bias
                }
            }
        }
        if (L.want_w && eq == 0 && et < nt) P.wout[L.slot * P.n + t0 + et] = wacc;
        __syncwarp();
        } // sub-launches
    }
}

struct Combine13 {
    const double *fit, *w; // [slots][n]
    double *out;
    long long n;
    int slots, dim;
    double fbias;
    double delta[5], cbias[5];
};

// cf_cal, cec2013.cpp:1091-1124
__global__ void cec13_combine_kernel(const __grid_constant__ Combine13 P)
{
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    double w[5], fit[5];
    double w_max = 0.0, w_sum = 0.0;
    const double nx = static_cast<double>(P.dim);
    for (int s = 0; s < P.slots; ++s) {
        fit[s] = P.fit[s * P.n + i] + P.cbias[s];
        double ws = P.w[s * P.n + i];
        if (ws != 0.0)
            ws = sqrt(1.0 / ws) * exp(-ws / 2.0 / nx / (P.delta[s] * P.delta[s]));
        else
            ws = 1.0e99;
        if (ws > w_max) w_max = ws;
        w[s] = ws;
    }
    for (int s = 0; s < P.slots; ++s) w_sum = w_sum + w[s];
    if (w_max == 0.0) {
        for (int s = 0; s < P.slots; ++s) w[s] = 1.0;
        w_sum = P.slots;
    }
    double f = 0.0;
    for (int s = 0; s < P.slots; ++s) f = f + w[s] / w_sum * fit[s];
    P.out[i] = f + P.fbias;
}

inline bool merge_enabled()
{
    const char *e = std::getenv("PGC_CEC13_MERGE"); // PGC_CEC13_MERGE=0: one launch per rotation at every batch size
    return !(e && e[0] == '0');
}

// island-sized batches (~1000 individuals) are latency-bound: shrink the tile until there are ~4 warps per SM
inline int tile_individuals(const pgc_ctx *ctx, long long n)
{
    // `sharers` contexts evaluate concurrently on this device (pgc_ctx_set_sharers: the islands of an archipelago that share a GPU):
    // together they fill the SMs, so each keeps somewhat fuller tiles - an m8n8k4 tile with one individual wastes 7/8 of every DMMA,
    // but a launch of few fat tiles is latency-bound: the target number of tiles shrinks with the square root of the sharers
    // (measured, cfg5 on one GPU, 8 islands of 1024: tiles of 1 / 2 / 4 / 8 individuals give 6.9e4 / 7.7e4 / 7.5e4 / 5.7e4 island-generations/s)
    const long long want = std::max(1ll, static_cast<long long>(static_cast<double>(ctx->sm_count) * 4. / std::sqrt(static_cast<double>(std::max(1, ctx->sharers)))));
    int ti = kTileInd;
    while (ti > 1 && (n + ti - 1) / ti < want) ti >>= 1;
    if (const char *e = std::getenv("PGC_CEC13_TI")) { // experiment switch: force the tile size of island-sized batches
        const int v = std::atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) ti = std::min(ti < kTileInd ? v : ti, kTileInd);
    }
    return ti;
}

template <int D> int launch13(pgc_ctx *ctx, const Params13 &pp, cudaStream_t stream)
{
    constexpr int DP = pad8(D), YS = ystride(D);
    static thread_local int configured_dev = -1;
    const int ti = tile_individuals(ctx, pp.n);
    const bool small = ti < kTileInd || pp.nl > 1;
    auto kern = small ? cec13_kernel<D, true> : cec13_kernel<D, false>;
    if (configured_dev != ctx->device) {
        PGC_CUDA(cudaFuncSetAttribute(cec13_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ctx->smem_optin)));
        PGC_CUDA(cudaFuncSetAttribute(cec13_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ctx->smem_optin)));
        configured_dev = ctx->device;
    }
    bool any_rot = false;
    for (int q = 0; q < pp.nl; ++q) any_rot = any_rot || pp.L[q].rot >= 0;
    const size_t fixed = sizeof(double) * ((any_rot ? pp.nl * DP * YS : 0) + D) + 16;
    const size_t per_warp = (sizeof(double) * 2 * ti * YS + sizeof(unsigned short) * ti * D + 15) / 16 * 16;
    int fit = static_cast<int>((ctx->smem_optin - fixed) / per_warp);
    if (fit > kMaxWarps13) fit = kMaxWarps13;
    if (small && fit > kSmallWarps13) fit = kSmallWarps13;
    PGC_REQUIRE(fit >= 1, "cec2013: shared memory too small for dimension %d", D);
    const long long ntiles = (pp.n + ti - 1) / ti;
    // small batches: fewer warps per CTA so that the tiles cover the SMs.  Alone, that is the lowest latency; but every CTA carries the
    // function's rotation matrices (80 KB for three at D = 50) whatever its number of warps, so two such CTAs fill an SM's shared memory
    // and a 128-CTA launch of 2-warp CTAs owns almost half the device: launches of different islands then queue behind each other
    // (measured, 8 streams of 1024-row f12 D=50 evaluations: 9.4 us per launch however many streams, profiles/r2q_concurrent_parts.json).
    // With `sharers` contexts on the device each launch takes its share of the CTA slots and fuller CTAs instead.
    long long spread = ctx->sm_count;
    if (small && ctx->sharers > 1) {
        int slots_per_sm = 2;
        if (const char *e = std::getenv("PGC_CEC13_SHARE_SLOTS")) slots_per_sm = std::max(1, std::atoi(e)); // experiment switch
        spread = std::max(1ll, static_cast<long long>(ctx->sm_count) * slots_per_sm / ctx->sharers);
    }
    long long w = (ntiles + spread - 1) / spread;
    if (w > fit) w = fit;
    if (w < 1) w = 1;
    long long ctas = (ntiles + w - 1) / w;
    if (ctas > ctx->sm_count) ctas = ctx->sm_count;
    Params13 launch_pp = pp;
    launch_pp.ti = ti;
    kern<<<static_cast<unsigned>(ctas), static_cast<unsigned>(w * 32), fixed + per_warp * w, stream>>>(launch_pp);
    PGC_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

// ---- plan construction (host) -----------------------------------------------------------------------------------
enum { Y = 0, Z = 1 };
struct LStep { int op, src, dst; double a, b; int which; }; // which: O_ROT matrix 0/1, O_TAB 10/100, O_ASY unused

std::vector<LStep> chain_of(int red)
{
    const double k420 = 4.209687462275036e+002;
    switch (red) {
        case R_SPHERE: return {{O_ROT, Y, Z, 0, 0, 0}};                                                           // :319-332
        case R_ELLIPS: return {{O_ROT, Y, Z, 0, 0, 0}, {O_OSZ, Z, Y}};                                            // :334-349
        case R_BENT: return {{O_ROT, Y, Z, 0, 0, 0}, {O_ASY, Z, Y, 0.5}, {O_ROT, Y, Z, 0, 0, 1}};                  // :351-373
        case R_DISCUS: return {{O_ROT, Y, Z, 0, 0, 0}, {O_OSZ, Z, Y}};                                            // :375-391
        case R_DIFPOW: return {{O_ROT, Y, Z, 0, 0, 0}};                                                           // :393-408
        case R_ROSEN: return {{O_MULDIV, Y, Y, 2.048, 100.}, {O_ROT, Y, Z, 0, 0, 0}, {O_ADD, Z, Z, 1.}};          // :410-436
        case R_SCHAF7:                                                                                            // :438-466
        case R_ACKLEY:                                                                                            // :468-500
            return {{O_ROT, Y, Z, 0, 0, 0}, {O_ASY, Z, Y, 0.5}, {O_TAB, Y, Z, 0, 0, 10}, {O_ROT, Z, Y, 0, 0, 1}};
        case R_WEIER:                                                                                             // :502-542
            return {{O_MULDIV, Y, Y, 0.5, 100.}, {O_ROT, Y, Z, 0, 0, 0}, {O_ASY, Z, Y, 0.5}, {O_TAB, Y, Z, 0, 0, 10}, {O_ROT, Z, Y, 0, 0, 1}};
        case R_GRIEW: return {{O_MULDIV, Y, Y, 600.0, 100.0}, {O_ROT, Y, Z, 0, 0, 0}, {O_TAB, Z, Z, 0, 0, 100}};  // :544-572
        case R_RASTR:                                                                                             // :574-615
            return {{O_MULDIV, Y, Y, 5.12, 100.}, {O_ROT, Y, Z, 0, 0, 0}, {O_OSZ, Z, Y}, {O_ASY, Y, Z, 0.2}, {O_ROT, Z, Y, 0, 0, 1},
                    {O_TAB, Y, Y, 0, 0, 10}, {O_ROT, Y, Z, 0, 0, 0}};
        case R_STEP_RASTR:                                                                                        // :617-661
            return {{O_MULDIV, Y, Y, 5.12, 100.}, {O_ROT, Y, Z, 0, 0, 0}, {O_STEP, Z, Z}, {O_OSZ, Z, Y}, {O_ASY, Y, Z, 0.2},
                    {O_ROT, Z, Y, 0, 0, 1}, {O_TAB, Y, Y, 0, 0, 10}, {O_ROT, Y, Z, 0, 0, 0}};
        case R_SCHWEF: return {{O_MUL, Y, Y, 1000. / 100.}, {O_ROT, Y, Z, 0, 0, 0}, {O_TAB, Z, Y, 0, 0, 10}, {O_ADD, Y, Z, k420}}; // :663-700
        case R_KATS: return {{O_MUL, Y, Y, 5.0 / 100.0}, {O_ROT, Y, Z, 0, 0, 0}, {O_TAB, Z, Z, 0, 0, 100}, {O_ROT, Z, Y, 0, 0, 1}}; // :702-739
        case R_BIRAS:                                                                                             // :741-799
            return {{O_MUL, Y, Y, 10.0 / 100.0}, {O_BISIGN, Y, Z}, {O_ROT, Z, Y, 0, 0, 0}, {O_TAB, Y, Y, 0, 0, 100}, {O_ROT, Y, Z, 0, 0, 1}};
        case R_GRROS: return {{O_MULDIV, Y, Y, 5., 100.}, {O_ADD, Y, Z, 1.}}; // :801-833, the rotation :812-816 is overwritten by :818-820
        case R_ESCAF: return {{O_ROT, Y, Z, 0, 0, 0}, {O_ASY, Z, Y, 0.5}, {O_ROT, Y, Z, 0, 0, 1}};                 // :835-865
        default: return {};
    }
}

int end_buffer_of(int red)
{
    switch (red) {
        case R_ELLIPS:
        case R_DISCUS:
        case R_SCHAF7:
        case R_ACKLEY:
        case R_WEIER:
        case R_KATS: return Y;
        default: return Z;
    }
}

// append the launches of one primitive (component `comp`, rotated or not) to the plan
int compile_chain(Cec2013Plan &pl, int red, bool rotated, int comp, int slot, double mul, double div, bool want_w)
{
    Launch13 cur{};
    auto fresh = [&](int from_x, int in_buf) {
        cur = Launch13{};
        cur.from_x = from_x;
        cur.in_buf = in_buf;
        cur.rot = -1;
        cur.comp = comp;
        cur.slot = slot;
    };
    fresh(1, Y);
    cur.want_w = want_w ? 1 : 0;
    auto push_step = [&](const Step13 &s) {
        if (cur.nsteps >= kMaxSteps13) return false;
        cur.st[cur.nsteps++] = s;
        return true;
    };
    for (const LStep &ls : chain_of(red)) {
        Step13 s{ls.op, ls.src, ls.dst, 0, ls.a, ls.b};
        if (ls.op == O_ROT) {
            if (!rotated) {
                s.op = O_COPY;
            } else {
                if (cur.rot >= 0) { // second matrix: cut here, the rotation's source is the only live vector
                    cur.to_state = 1;
                    cur.end_buf = ls.src;
                    pl.launches.push_back(cur);
                    pl.needs_state = true;
                    fresh(0, ls.src);
                }
                cur.rot = comp + ls.which;
                PGC_REQUIRE(cur.rot < 10, "cec2013: component %d needs rotation matrix %d, the table holds 10", comp, cur.rot);
            }
        } else if (ls.op == O_TAB) {
            s.tab = ls.which == 10 ? pl.t10 : pl.t100;
        } else if (ls.op == O_ASY) {
            s.tab = ls.a == 0.5 ? pl.tasy5 : pl.tasy2;
        }
        PGC_REQUIRE(push_step(s), "cec2013: internal error, chain of primitive %d exceeds %d steps per launch", red, kMaxSteps13);
    }
    cur.to_state = 0;
    cur.end_buf = end_buffer_of(red);
    cur.red = red == R_STEP_RASTR ? R_RASTR : red;
    cur.scaled = mul != 0.0;
    cur.mul = mul;
    cur.div = div;
    pl.launches.push_back(cur);
    return PGC_OK;
}

struct Part13 { int red; double mul, div; bool own_unrotated; };

} // namespace

int cec2013_create(pgc_problem *p, const pgc_problem_desc *d)
{
    const unsigned func = d->prob_id, dim = d->dim;
    PGC_REQUIRE(func >= 1 && func <= 28, "Error: CEC2013 Test functions are only defined for prob_id in [1, 28], a prob_id of %u was detected.",
                func); // cec2013.cpp:59-63
    PGC_REQUIRE(dim == 2 || dim == 5 || (dim >= 10 && dim <= 100 && dim % 10 == 0),
                "Error: CEC2013 Test functions are only defined for dimensions 2,5,10,20,30,40,50,60,70,80,90,100, a dimension of %u was detected.",
                dim); // :53-58
    const size_t D = dim;
    auto *pl = new Cec2013Plan;
    p->cec13 = pl;
    pl->func = static_cast<int>(func);
    pl->dim = static_cast<int>(dim);
    static const double FB[28] = {-1400, -1300, -1200, -1100, -1000, -900, -800, -700, -600, -500, -400, -300, -200, -100,
                                  100, 200, 300, 400, 500, 600, 700, 800, 900, 1000, 1100, 1200, 1300, 1400}; // :83-193
    pl->fbias = FB[func - 1];

    // constant table: the reference's own expressions, evaluated once on the host with the host libm
    std::vector<double> tab;
    auto add_table = [&](auto fn, int count) {
        const int off = static_cast<int>(tab.size());
        for (int i = 0; i < count; ++i) tab.push_back(fn(static_cast<unsigned>(i)));
        return off;
    };
    const unsigned nx = dim;
    pl->t10 = add_table([&](unsigned i) { return std::pow(10.0, (1. * i) / (nx - 1u) / 2.0); }, dim);
    pl->t100 = add_table([&](unsigned i) { return std::pow(100.0, (1. * i) / (nx - 1u) / 2.0); }, dim);
    pl->tell = add_table([&](unsigned i) { return std::pow(10.0, (6. * i) / (nx - 1u)); }, dim);
    pl->tdif = add_table([&](unsigned i) { return 2. + (4. * i) / (nx - 1u); }, dim);
    pl->tgri = add_table([&](unsigned i) { return 1.0 / std::sqrt(1.0 + i); }, dim);
    const double pi = 3.141592653589793238462643383279502884;
    pl->twei = add_table([&](unsigned j) { return 2.0 * pi * std::pow(3.0, j); }, 21);
    pl->tasy5 = add_table([&](unsigned i) { return (0.5 * i) / (nx - 1u); }, dim);
    pl->tasy2 = add_table([&](unsigned i) { return (0.2 * i) / (nx - 1u); }, dim);
    {
        double sum2 = 0.0; // :531-536
        for (unsigned j = 0; j <= 20; ++j) sum2 += std::pow(0.5, j) * std::cos(2.0 * pi * std::pow(3.0, j) * 0.5);
        pl->weier_c0 = nx * sum2;
        const double tmp3 = std::pow(1.0 * nx, 1.2); // :707
        pl->kats_c0 = 10.0 / tmp3;
        pl->kats_c1 = 10.0 / nx / nx; // :737
        pl->bi_s = 1.0 - 1.0 / (2.0 * std::pow(nx + 20.0, 0.5) - 8.2); // :748
        pl->bi_mu1 = -std::pow((2.5 * 2.5 - 1.0) / pl->bi_s, 0.5);      // :749
    }

    static const struct { int red, r; } F[20] = {{R_SPHERE, 0}, {R_ELLIPS, 1}, {R_BENT, 1}, {R_DISCUS, 1}, {R_DIFPOW, 0}, {R_ROSEN, 1},
                                                 {R_SCHAF7, 1}, {R_ACKLEY, 1}, {R_WEIER, 1}, {R_GRIEW, 1}, {R_RASTR, 0}, {R_RASTR, 1},
                                                 {R_STEP_RASTR, 1}, {R_SCHWEF, 0}, {R_SCHWEF, 1}, {R_KATS, 1}, {R_BIRAS, 0}, {R_BIRAS, 1},
                                                 {R_GRROS, 1}, {R_ESCAF, 1}}; // :82-161
    int rc = PGC_OK;
    if (func <= 20) {
        pl->ncomp = 1;
        rc = compile_chain(*pl, F[func - 1].red, F[func - 1].r != 0, 0, -1, 0.0, 0.0, false);
    } else {
        static const struct { int n; Part13 part[5]; double delta[5]; } C[8] = {
            {5, {{R_ROSEN, 10000, 1e+4}, {R_DIFPOW, 10000, 1e+10}, {R_BENT, 10000, 1e+30}, {R_DISCUS, 10000, 1e+10}, {R_SPHERE, 10000, 1e+5, true}},
             {10, 20, 30, 40, 50}},                                                                                   // cf01 :867-892
            {3, {{R_SCHWEF}, {R_SCHWEF}, {R_SCHWEF}}, {20, 20, 20}},                                                  // cf02 :894-905
            {3, {{R_SCHWEF}, {R_SCHWEF}, {R_SCHWEF}}, {20, 20, 20}},                                                  // cf03 :907-918
            {3, {{R_SCHWEF, 1000, 4e+3}, {R_RASTR, 1000, 1e+3}, {R_WEIER, 1000, 400}}, {20, 20, 20}},                 // cf04 :920-938
            {3, {{R_SCHWEF, 1000, 4e+3}, {R_RASTR, 1000, 1e+3}, {R_WEIER, 1000, 400}}, {10, 30, 50}},                 // cf05 :940-958
            {5, {{R_SCHWEF, 1000, 4e+3}, {R_RASTR, 1000, 1e+3}, {R_ELLIPS, 1000, 1e+10}, {R_WEIER, 1000, 400}, {R_GRIEW, 1000, 100}},
             {10, 10, 10, 10, 10}},                                                                                   // cf06 :960-984
            {5, {{R_GRIEW, 10000, 100}, {R_RASTR, 10000, 1e+3}, {R_SCHWEF, 10000, 4e+3}, {R_WEIER, 10000, 400}, {R_SPHERE, 10000, 1e+5, true}},
             {10, 10, 10, 20, 20}},                                                                                   // cf07 :986-1010
            {5, {{R_GRROS, 10000, 4e+3}, {R_SCHAF7, 10000, 4e+6}, {R_SCHWEF, 10000, 4e+3}, {R_ESCAF, 10000, 2e+7}, {R_SPHERE, 10000, 1e+5, true}},
             {10, 20, 30, 40, 50}},                                                                                   // cf08 :1012-1036
        };
        const auto &c = C[func - 21];
        pl->ncomp = c.n;
        const bool r_flag = func != 22; // :147
        for (int i = 0; i < c.n && rc == PGC_OK; ++i) {
            pl->delta[i] = c.delta[i];
            pl->cbias[i] = 100.0 * i;
            rc = compile_chain(*pl, c.part[i].red, r_flag && !c.part[i].own_unrotated, i, i, c.part[i].mul, c.part[i].div, true);
        }
    }
    if (rc != PGC_OK) return rc;

    int max_rot = -1;
    double flops = 0.0;
    for (const Launch13 &L : pl->launches) {
        if (L.rot > max_rot) max_rot = L.rot;
        flops += (L.rot >= 0 ? 2.0 * D * D : 0.0) + 2.0 * D * (L.nsteps + 1);
    }
    PGC_REQUIRE(d->shift && d->shift_len >= static_cast<size_t>(pl->ncomp) * D,
                "cec2013: shift table needs at least %zu values (component i at i*dim, cec2013.cpp:878), got %zu", pl->ncomp * D, d->shift_len);
    PGC_REQUIRE(max_rot < 0 || (d->rotation && d->rotation_len >= static_cast<size_t>(max_rot + 1) * D * D),
                "cec2013: rotation table needs at least %zu values, got %zu", (max_rot + 1) * D * D, d->rotation_len);

    p->nx = D;
    p->nobj = 1;
    p->lb.assign(D, -100.0); // :206-212
    p->ub.assign(D, 100.0);
    static const char *names[29] = {"", "sphere_func", "ellips_func", "bent_cigar_func", "discus_func", "dif_powers_func_non_rotated",
                                    "rosenbrock_func", "schaffer_F7_func", "ackley_func", "weierstrass_func", "griewank_func",
                                    "rastrigin_func_non_rotated", "rastrigin_func", "step_rastrigin_func", "schwefel_func_non_rotated",
                                    "schwefel_func", "katsuura_func", "bi_rastrigin_func_non_rotated", "bi_rastrigin_func", "grie_rosen_func",
                                    "escaffer6_func", "cf01", "cf02", "cf03", "cf04", "cf05", "cf06", "cf07", "cf08"};
    p->name = "CEC2013 - f" + std::to_string(func) + "(" + names[func] + ")"; // :218-310
    p->flops_per_eval = flops;
    p->transc_per_eval = 0;

    PGC_CUDA(cudaSetDevice(p->ctx->device));
    PGC_CUDA(cudaMalloc(&p->d_shift, sizeof(double) * pl->ncomp * D));
    PGC_CUDA(cudaMemcpy(p->d_shift, d->shift, sizeof(double) * pl->ncomp * D, cudaMemcpyHostToDevice));
    if (max_rot >= 0) {
        const size_t img = static_cast<size_t>(pad8(dim)) * ystride(dim);
        std::vector<double> tiled(static_cast<size_t>(max_rot + 1) * img, 0.0);
        const int YS = ystride(dim);
        for (int c = 0; c <= max_rot; ++c)
            for (size_t i = 0; i < D; ++i)
                for (size_t k = 0; k < D; ++k) tiled[c * img + i * YS + k] = d->rotation[c * D * D + i * D + k];
        PGC_CUDA(cudaMalloc(&p->d_rotation, sizeof(double) * tiled.size()));
        PGC_CUDA(cudaMemcpy(p->d_rotation, tiled.data(), sizeof(double) * tiled.size(), cudaMemcpyHostToDevice));
    }
    PGC_CUDA(cudaMalloc(&p->d_table, sizeof(double) * tab.size()));
    PGC_CUDA(cudaMemcpy(p->d_table, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice));
    return PGC_OK;
}

void cec2013_destroy(pgc_problem *p)
{
    delete p->cec13;
    p->cec13 = nullptr;
}

int cec2013_eval(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream)
{
    if (n == 0) return PGC_OK;
    const Cec2013Plan &pl = *p->cec13;
    pgc_ctx *ctx = p->ctx;
    const size_t D = pl.dim, img = static_cast<size_t>(pad8(pl.dim)) * ystride(pl.dim);
    const bool comp = pl.func > 20;
    const size_t state_elems = pl.needs_state ? n * D : 0, slot_elems = comp ? 10 * n : 0;
    double *state = nullptr, *fit = nullptr, *w = nullptr;
    if (state_elems + slot_elems) {
        int rc = ensure_scratch(ctx, sizeof(double) * (state_elems + slot_elems));
        if (rc != PGC_OK) return rc;
        state = static_cast<double *>(ctx->scratch);
        fit = state + state_elems;
        w = fit + 5 * n;
    }
    auto launch_group = [&](const Launch13 *Ls, int nl) -> int {
        Params13 pp{};
        pp.x = d_dvs;
        pp.state = state;
        pp.nl = nl;
        for (int q = 0; q < nl; ++q) {
            pp.mr[q] = Ls[q].rot >= 0 ? p->d_rotation + Ls[q].rot * img : nullptr;
            pp.L[q] = Ls[q];
        }
        pp.os = p->d_shift + Ls[0].comp * D;
        pp.table = p->d_table;
        pp.out = comp ? fit : d_fvs;
        pp.wout = w;
        pp.n = static_cast<long long>(n);
        pp.fbias = pl.fbias;
        pp.weier_c0 = pl.weier_c0;
        pp.kats_c0 = pl.kats_c0;
        pp.kats_c1 = pl.kats_c1;
        pp.bi_s = pl.bi_s;
        pp.bi_mu1 = pl.bi_mu1;
        pp.tell = pl.tell;
        pp.tdif = pl.tdif;
        pp.tgri = pl.tgri;
        pp.twei = pl.twei;
        pp.strict = p->strict;
        pp.ti = kTileInd; // launch13 picks the tile size from the batch size
        switch (pl.dim) {
            case 2: return launch13<2>(ctx, pp, stream);
            case 5: return launch13<5>(ctx, pp, stream);
            case 10: return launch13<10>(ctx, pp, stream);
            case 20: return launch13<20>(ctx, pp, stream);
            case 30: return launch13<30>(ctx, pp, stream);
            case 40: return launch13<40>(ctx, pp, stream);
            case 50: return launch13<50>(ctx, pp, stream);
            case 60: return launch13<60>(ctx, pp, stream);
            case 70: return launch13<70>(ctx, pp, stream);
            case 80: return launch13<80>(ctx, pp, stream);
            case 90: return launch13<90>(ctx, pp, stream);
            case 100: return launch13<100>(ctx, pp, stream);
            default: set_error("cec2013: unsupported dimension %d", pl.dim); return PGC_ERR_INVALID_ARGUMENT;
        }
    };
    // Island-sized batches of a single-component function whose matrices all fit in shared memory: the whole chain in ONE launch
    // (the vector stays in the warp's buffers between the rotations instead of a round trip through the state array; same bits).
    const size_t nl_all = pl.launches.size();
    const size_t merged_smem = sizeof(double) * (nl_all * img + D) + 16 + 4 * (sizeof(double) * 2 * ystride(pl.dim) + sizeof(unsigned short) * D + 16);
    const bool merge = !comp && nl_all > 1 && nl_all <= static_cast<size_t>(kMaxMerge) && tile_individuals(ctx, static_cast<long long>(n)) < kTileInd
                       && merged_smem <= ctx->smem_optin && merge_enabled();
    if (merge) {
        int rc = launch_group(pl.launches.data(), static_cast<int>(nl_all));
        if (rc != PGC_OK) return rc;
    } else {
        for (const Launch13 &L : pl.launches) {
            int rc = launch_group(&L, 1);
            if (rc != PGC_OK) return rc;
        }
    }
    if (comp) {
        Combine13 cp;
        cp.fit = fit;
        cp.w = w;
        cp.out = d_fvs;
        cp.n = static_cast<long long>(n);
        cp.slots = pl.ncomp;
        cp.dim = pl.dim;
        cp.fbias = pl.fbias;
        for (int s = 0; s < 5; ++s) {
            cp.delta[s] = pl.delta[s];
            cp.cbias[s] = pl.cbias[s];
        }
        cec13_combine_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(cp);
        PGC_CUDA(cudaGetLastError());
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
    }
    return PGC_OK;
}

} // namespace pgc
