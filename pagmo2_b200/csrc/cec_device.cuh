// pagmo2_b200/csrc/cec_device.cuh - device helpers shared by the CEC2014 and CEC2013 evaluators: operand-tile geometry for
// the FP64 tensor path (mma.sync.m8n8k4.f64), branch-free trigonometry, and the "4 lanes per individual" ordered reductions
// of the epilogues (8-individual warp tiles).
#pragma once
#include <cuda_runtime.h>

namespace pgc
{
namespace cecdev
{

constexpr int kTileInd = 8;         // individuals per warp tile (one m8n8k4 row tile)
constexpr int kLPI = 32 / kTileInd; // lanes per individual in the epilogue
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ constexpr int pad8(int d) { return (d + 7) / 8 * 8; }
__host__ __device__ constexpr int pad4(int d) { return (d + 3) / 4 * 4; }
// Row stride (doubles) of the row-major operand tiles - the warp's Y tile (rows = individuals) and the rotation
// image (rows = outputs): >= the padded inner length and == 4 or 12 (mod 16), so that the 16 lanes of a half
// warp (4 rows x 4 consecutive doubles of an m8n8k4 operand fragment) cover all 32 banks exactly once.
__host__ __device__ constexpr int ystride(int d)
{
    int s = pad4(d);
    while (s % 16 != 4 && s % 16 != 12) s += 4;
    return s;
}

// D(8x8) += A(8x4) * B(4x8), FP64 tensor path.  Lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2*(l%4)+{0,1}].
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ---- branch-free FP64 trigonometry for the epilogues ---------------------------------------------------------
// libdevice's sin/cos carry a data-dependent branch (Payne-Hanek slow path), which stops the compiler from
// interleaving independent evaluations; the epilogues then run at the latency of one dependent DFMA chain per
// coordinate.  These versions are straight-line code (selects only), ~1 ulp, valid for |theta| < 2^50, so two
// coordinates per lane overlap.  The argument is the reference's own rounded double (e.g. fl(2*pi*z)):
// turns = frac(theta / (2 pi)) with a two-term 1/(2 pi) (error ~1e-17 turns), then a quadrant fold to
// [-pi/4, pi/4] and the classic minimax kernels (fdlibm k_sin.c / k_cos.c coefficients).
__device__ __forceinline__ double round_magic(double x) // round to nearest integer, |x| < 2^51
{
    return (x + 6755399441055744.0) - 6755399441055744.0;
}

__device__ __forceinline__ double turns_of(double theta) // frac(theta / (2 pi)) in [-0.5, 0.5]
{
    const double I1 = 0x1.45f306dc9c883p-3, I2 = -0x1.6b01ec5417056p-57;
    const double p = theta * I1;
    const double e = fma(theta, I1, -p);
    return (p - rint(p)) + fma(theta, I2, e);
}

// sin and cos of 2*pi*r for |r| <~ 1
__device__ __forceinline__ void sincos_turns(double r, double &sn, double &cs)
{
    const double q = round_magic(4.0 * r);
    const double f = fma(-0.25, q, r); // exact, in [-1/8, 1/8]
    const int iq = __double2int_rn(q);
    const double t = fma(f, 6.283185307179586232, f * 2.4492935982947064e-16); // 2*pi*f, hi + lo
    const double z = t * t;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double st = fma(t * z, ps, t);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double ct = fma(z * z, pc, fma(-0.5, z, 1.0));
    // angle = t + iq*pi/2
    const bool odd = iq & 1;
    const double s0 = odd ? ct : st, c0 = odd ? st : ct;
    sn = (iq & 2) ? -s0 : s0;
    cs = ((iq + 1) & 2) ? -c0 : c0;
}

// cos(2*pi*r) alone, |r| < 2^49: fold to f = r - q/2 in [-1/4, 1/4] (q = nearest integer to 2r, exact), one even
// polynomial in f^2 over the whole half period (Chebyshev interpolant of degree 8 in f^2, truncation 4e-18; coefficients
// from scripts/gen_cos_turns_poly.py) and the sign (-1)^q taken from the parity bit the rounding constant leaves in the
// low word.  12 FP64 instructions, no conversion, no select; ABSOLUTE error ~3e-16 (not a relative bound near the zeros,
// which the sums and products of the primitives do not need).
// CM = true reads the coefficients from constant memory: in the issue-bound separable kernel the literals were
// re-materialised with two integer moves per use (46 of 86 instructions per coordinate, ncu r1m_sep_f8).  The tile kernel
// keeps literals (CM = false): there the issue slots are idle, and the constant-memory build measured 1.5-3 % slower.
static __constant__ double kCosTurns[9] = {0x1.0000000000000p+0,  -0x1.3bd3cc9be45dbp+4, 0x1.03c1f081b5992p+6,
                                           -0x1.55d3c7e3bfbf5p+6, 0x1.e1f506813a321p+5,  -0x1.a6d1efc8c38bep+4,
                                           0x1.f9d254582ac30p+2,  -0x1.b6957b54dd389p+0, 0x1.1678f9078a9b3p-2};

template <bool CM = false> __device__ __forceinline__ double cos_turns(double r)
{
    const double t = fma(2.0, r, 6755399441055744.0);
    const double q = t - 6755399441055744.0;
    const double f = fma(-0.5, q, r);
    const double u = f * f;
    double p;
    if (CM) {
        p = fma(u, kCosTurns[8], kCosTurns[7]);
#pragma unroll
        for (int k = 6; k >= 0; --k) p = fma(u, p, kCosTurns[k]);
    } else {
        p = fma(u, 0x1.1678f9078a9b3p-2, -0x1.b6957b54dd389p+0);
        p = fma(u, p, 0x1.f9d254582ac30p+2);
        p = fma(u, p, -0x1.a6d1efc8c38bep+4);
        p = fma(u, p, 0x1.e1f506813a321p+5);
        p = fma(u, p, -0x1.55d3c7e3bfbf5p+6);
        p = fma(u, p, 0x1.03c1f081b5992p+6);
        p = fma(u, p, -0x1.3bd3cc9be45dbp+4);
        p = fma(u, p, 1.0);
    }
    const int flip = __double2loint(t) << 31;
    return __hiloint2double(__double2hiint(p) ^ flip, __double2loint(p));
}

template <bool CM = false> __device__ __forceinline__ double cos_theta(double theta) { return cos_turns<CM>(turns_of(theta)); }

// sin(2 pi r) = cos(2 pi (r - 1/4)); the shift costs at most 2^-54 turns
template <bool CM = false> __device__ __forceinline__ double sin_theta(double theta)
{
    return cos_turns<CM>(turns_of(theta) - 0.25);
}

// 1/d for a normal, positive d: hardware seed (rcp.approx.ftz.f64, ~2^-20) + one cubic step - about 1 ulp, branch-free
__device__ __forceinline__ double fast_rcp(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);  // r (1 + e + e^2) = 1/d (1 - e^3): one cubic step takes the 2^-20 seed below 2^-53
    return fma(r, fma(e, e, e), r);
}

// sqrt(x) for x >= 0 without the IEEE routine's special-case branch: hardware seed (rsqrt.approx.ftz.f64), two coupled
// Goldschmidt steps on (g ~ sqrt x, h ~ 1 / (2 sqrt x)) and one residual correction - <= 1 ulp on normal inputs; x = 0 (and
// denormals) are lifted to the smallest magnitudes whose root the callers multiply by x anyway.
__device__ __forceinline__ double sqrt_nobranch(double x)
{
    const double xs = fmax(x, 1.0e-300);
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(xs));
    double g = xs * r, h = 0.5 * r;
    double e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    return fma(fma(-g, g, xs), h, g);
}

// katsuura's inner sum  sum_{k=1..32} |2^k z - floor(2^k z + 0.5)| / 2^k  (cec2014.cpp:604-611, cec2013.cpp:726-735).
// |2^k z - floor(2^k z + 0.5)| is the distance d_k of 2^k z to the nearest integer; it obeys the tent map
// d_{k+1} = 1/2 - |2 d_k - 1/2|, every step exact in FP64 for |z| >~ 2^-10 (below that a step may round by <= 2^-55, far
// inside the tolerance), and "/ 2^k" is an exact scaling: three instructions per term instead of a multiply, a rounding, a
// subtraction and a division.  Huge |z| (>= 2^18, outside every CEC domain) keeps the reference's own loop.
__device__ __forceinline__ double katsuura_inner(double z)
{
    double temp = 0.0;
    if (fabs(z) < 262144.0) {
        const double t2 = z + z;
        double d = fabs(t2 - round_magic(t2)), it1 = 0.5;
        temp = d * it1;
#pragma unroll
        for (int k = 2; k <= 32; ++k) {
            it1 *= 0.5;
            d = 0.5 - fabs(fma(2.0, d, -0.5));
            temp = fma(d, it1, temp);
        }
    } else {
        double t1 = 1.0;
        for (int k = 1; k <= 32; ++k) {
            t1 *= 2.0;
            const double t2 = t1 * z;
            temp += fabs(t2 - floor(t2 + 0.5)) / t1;
        }
    }
    return temp;
}

// one angle-tripling step w -> w^3 on the unit circle, w = cs + i sn: c (c^2 - 3 s^2) + i s (3 c^2 - s^2), six FP64 instructions
__device__ __forceinline__ void triple_angle(double &sn, double &cs)
{
    const double cc = cs * cs, ss = sn * sn;
    const double c3 = cs * fma(-3.0, ss, cc), s3 = sn * fma(3.0, cc, -ss);
    cs = c3;
    sn = s3;
}

// the kLPI lanes of an individual are kTileInd apart (lane = q * kTileInd + individual)
__device__ __forceinline__ double pair_add(double v)
{
#pragma unroll
    for (int m = kTileInd; m < 32; m <<= 1) v = v + __shfl_xor_sync(kFull, v, m);
    return v;
}
__device__ __forceinline__ double pair_mul(double v)
{
#pragma unroll
    for (int m = kTileInd; m < 32; m <<= 1) v = v * __shfl_xor_sync(kFull, v, m);
    return v;
}

// Sum term(j) for j = lo, lo + kLPI, ... < hi IN ORDER (lane q of an individual takes the terms j == q mod kLPI),
// evaluating two terms at a time so that their (independent, branch-free) dependency chains overlap.
template <class F> __device__ __forceinline__ double ordered_sum(int lo, int hi, F term)
{
    double s = 0.0;
    int j = lo;
    for (; j + kLPI < hi; j += 2 * kLPI) {
        const double a = term(j), b = term(j + kLPI);
        s += a;
        s += b;
    }
    if (j < hi) s += term(j);
    return s;
}

// Product of term(j) over the same index set, same two-at-a-time evaluation.
template <class F> __device__ __forceinline__ double ordered_prod(int lo, int hi, F term)
{
    double p = 1.0;
    int j = lo;
    for (; j + kLPI < hi; j += 2 * kLPI) {
        const double a = term(j), b = term(j + kLPI);
        p *= a;
        p *= b;
    }
    if (j < hi) p *= term(j);
    return p;
}

} // namespace cecdev
} // namespace pgc
