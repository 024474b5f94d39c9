// rot_i8.cu - see rot_i8.cuh: z = Mr * ((x - Os) * rate) through tcgen05.mma kind::i8 with FP64 accuracy (Ozaki digit planes).
// This file holds the digit-plane builder for Mr (host), the warp-specialised kernel, and the debug entry point that returns z
// itself (pgc_debug_rot_i8) so that the tensor-core rotation can be compared with an extended-precision product entry by entry.
#include <cmath>
#include <cstring>
#include <vector>

#include "pgc_internal.cuh"
#include "rot_i8.cuh"

namespace pgc
{
namespace i8rot
{

// ---- host: digit planes of a rotation matrix --------------------------------------------------------------------------------
// M: D x D row-major (reference layout, cec2014.cpp:1231: z[i] += x[j] * Mr[i * nx + j]); perm (optional, 0-based): row p of the
// image is row perm[p] of M (the hybrids' y[p] = z[S[p] - 1], cec2014.cpp:807-809).  img: kSlices planes in a_offset() layout;
// row_scale[p] = 2^(e_p - 54 + 48): what the epilogue multiplies the base-256 Horner value of row p by (the 2^48 = 256^6 of the
// digit weights is folded in here).
void build_matrix_planes(const double *M, int D, const int *perm, std::vector<int8_t> &img, std::vector<double> &row_scale)
{
    img.assign(static_cast<size_t>(kSlices) * kABytes, 0);
    row_scale.assign(kM, 0.);
    for (int p = 0; p < D; ++p) {
        const double *row = M + static_cast<size_t>(perm ? perm[p] : p) * D;
        double mx = 0.;
        for (int k = 0; k < D; ++k) mx = std::fmax(mx, std::fabs(row[k]));
        if (!(mx > 0.) || !std::isfinite(mx)) continue; // a zero row contributes nothing
        int e = 0;
        std::frexp(mx, &e); // mx = f * 2^e, f in [0.5, 1)  ->  |row| < 2^e
        row_scale[p] = std::ldexp(1.0, e - kScaleBits + 48);
        for (int k = 0; k < D; ++k) {
            long long Y = std::llrint(std::ldexp(row[k], kScaleBits - e));
            unsigned long long Yb = static_cast<unsigned long long>(Y + 0x0080808080808080LL);
            for (int s = 0; s < kSlices; ++s) {
                const unsigned byte = static_cast<unsigned>((Yb >> (8 * (kSlices - 1 - s))) & 0xffu);
                img[static_cast<size_t>(s) * kABytes + a_offset(p, k)] = static_cast<int8_t>(byte ^ 0x80u);
            }
        }
    }
}

namespace
{

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\t"
                 "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// COLL: 0 = plain, 1 = collector::a::fill (keep this A tile in the tensor core's operand collector), 2 = ::use, 3 = ::lastuse
template <int COLL>
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
#define PGC_MMA_I8(QUAL)                                                                                               \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                                  \
                 "tcgen05.mma.cta_group::1.kind::i8" QUAL " [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),  \
                 "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)                                        \
                 : "memory")
    if (COLL == 1) PGC_MMA_I8(".collector::a::fill");
    else if (COLL == 2) PGC_MMA_I8(".collector::a::use");
    else if (COLL == 3) PGC_MMA_I8(".collector::a::lastuse");
    else PGC_MMA_I8("");
#undef PGC_MMA_I8
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, no swizzle (layout type 0), descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return static_cast<uint64_t>((saddr >> 4) & 0x3fffu) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16)
           | (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor, kind::i8: D = S32, A = B = signed 8 bit, both K-major, N = 32, M = 128
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kTileN >> 3) << 17) | (static_cast<uint32_t>(kM >> 4) << 24);

constexpr int kEpiWarps = 4, kProdWarps = 8, kWarps = kEpiWarps + 1 + kProdWarps; // 13 warps
constexpr int kThreads = kWarps * 32;

struct Smem {
    int8_t a[kSlices][kABytes];            // 114 688 B
    int8_t b[kStages][kSlices][kBBytes];   //  86 016 B
    double os[kK];                         // shift (zero padded)
    double row_scale[kM];
    double coef[kM];                       // per-coordinate coefficient of the sum epilogue
    double ind_scale[8][kTileN];           // 2^(e - 54) of every decision vector of tile `it & 7` (read by the epilogue, which
                                           // runs at most 4 tiles behind the producers: a ring of 8 never collides)
    double part[2][kEpiWarps][kTileN];     // per-warp partial sums, double buffered by tile parity
    uint64_t full[kStages], empty[kStages], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

struct Params {
    const double *x;          // [n x D]
    const int8_t *planes;     // kSlices * kABytes (device)
    const double *row_scale;  // kM
    const double *os;         // D
    const double *coef;       // D (MODE 1)
    double *out;              // MODE 0: z [n x D]; MODE 1: f [n]
    long long n;
    double rate, fbias;
    int D;
    long long *prof; // optional: block 0 writes {producer wait, producer work, mma wait, mma issue, epilogue wait, epilogue work} cycles
};

// MODE 0: write z; MODE 1: f = sum_j coef[j] * z_j^2 + fbias
template <int MODE> __global__ void __launch_bounds__(kThreads, 1) rot_i8_kernel(const __grid_constant__ Params P)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = P.D;
    const long long ntiles = (P.n + kTileN - 1) / kTileN;

    // ---- one-off: digit planes of Mr, shift, scales; barriers; TMEM ----
    {
        const int4 *src = reinterpret_cast<const int4 *>(P.planes);
        int4 *dst = reinterpret_cast<int4 *>(&S.a[0][0]);
        for (int i = threadIdx.x; i < kSlices * kABytes / 16; i += kThreads) dst[i] = src[i];
        for (int i = threadIdx.x; i < kK; i += kThreads) S.os[i] = i < D ? P.os[i] : 0.;
        for (int i = threadIdx.x; i < kM; i += kThreads) {
            S.row_scale[i] = P.row_scale[i];
            S.coef[i] = (MODE == 1 && i < D) ? P.coef[i] : 0.;
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&S.full[s], 4 * 32);  // every lane of the 4 producer warps of a tile arrives
            mbar_init(&S.empty[s], 1);      // tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&S.acc_full[b], 1);           // tcgen05.commit
            mbar_init(&S.acc_empty[b], kEpiWarps);  // one lane per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == kEpiWarps) { // the MMA warp owns the tensor memory: all 512 columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async(); // the plain stores of the digit planes must be visible to the tensor core's (async proxy) reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    if (warp > kEpiWarps) {
        // ===== producers: two groups of 4 warps, alternating tiles ======================================================
        const int pw = warp - kEpiWarps - 1, group = pw >> 2, g8 = pw & 3; // g8: the 8-individual row group of this warp
        const int nloc = lane & 7, kq = lane >> 3;                          // individual in the row group, k-block within a pass
        long long it = group, t_wait = 0, t_work = 0, c0 = clock64();
        for (long long tile = blockIdx.x + static_cast<long long>(group) * gridDim.x; tile < ntiles; tile += 2ll * gridDim.x, it += 2) {
            const int s = static_cast<int>(it % kStages);
            mbar_wait(&S.empty[s], static_cast<unsigned>(((it / kStages) & 1) ^ 1));
            { const long long c1 = clock64(); t_wait += c1 - c0; c0 = c1; }
            const long long ind = tile * kTileN + g8 * 8 + nloc;
            const bool live = ind < P.n;
            const double *xr = P.x + ind * D;
            double y[2][16];
            int mxh = 0; // max over the vector of the high word of |y|: monotone in |y|, and only the exponent is needed
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int k0 = (pass * 4 + kq) * 16;
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    double2 v = make_double2(0., 0.);
                    const int k = k0 + c;
                    if (live && k < D) v = *reinterpret_cast<const double2 *>(xr + k); // D is even: rows are 16-byte aligned
                    const double a = (k < D) ? (v.x - S.os[k]) * P.rate : 0.;
                    const double b = (k + 1 < D) ? (v.y - S.os[k + 1]) * P.rate : 0.;
                    y[pass][c] = a;
                    y[pass][c + 1] = b;
                    mxh = max(mxh, max(__double2hiint(a) & 0x7fffffff, __double2hiint(b) & 0x7fffffff));
                }
            }
            mxh = max(mxh, __shfl_xor_sync(0xffffffffu, mxh, 8));
            mxh = max(mxh, __shfl_xor_sync(0xffffffffu, mxh, 16));
            // |y| < 2^e with e = exponent field - 1022; scale up by 2^(54 - e), remember 2^(e - 54)
            const int ebits = (mxh >> 20) & 0x7ff;
            const bool zero = ebits < 64 || ebits == 0x7ff; // an all-zero (or non-finite) vector: digits 0, scale 0 -> z = 0
            const double up = zero ? 0. : __hiloint2double((2099 - ebits) << 20, 0);
            const double up_hi = zero ? 0. : __hiloint2double((2099 - 24 - ebits) << 20, 0); // up * 2^-24
            if (kq == 0) S.ind_scale[it & 7][g8 * 8 + nloc] = zero ? 0. : __hiloint2double((ebits - 53) << 20, 0);
            // The 54-bit integer Y = rint(y * up) is formed WITHOUT 64-bit conversions (F2I.S64.F64 runs on the quarter-rate XU pipe,
            // which saturated in the first version of this kernel): H = rint(y * up * 2^-24) and L = rint(y * up - H * 2^24) come out
            // of the low mantissa word of (value + 1.5 * 2^52); Y = H * 2^24 + L, |H| < 2^30, |L| <= 2^23.  Balanced base-256 digits:
            // Y + 0x80808080808080 = (H + 0x80808080 + carry) * 2^24 + ((L + 0x808080) mod 2^24); digit = byte ^ 0x80.
            const double magic = 6755399441055744.0; // 1.5 * 2^52
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                unsigned lo[16], hi[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const double tm = y[pass][c] * up_hi + magic;
                    const double hd = tm - magic;
                    const double r = fma(-hd, 16777216.0, y[pass][c] * up);
                    const unsigned lw = static_cast<unsigned>(__double2loint(r + magic)) + 0x808080u;
                    lo[c] = lw;
                    hi[c] = static_cast<unsigned>(__double2loint(tm)) + 0x80808080u + (lw >> 24);
                }
                const int kb = pass * 4 + kq;
#pragma unroll
                for (int sl = 0; sl < kSlices; ++sl) {
                    const int byte = kSlices - 1 - sl; // digit plane 0 = most significant byte (6)
                    const unsigned *src = byte < 3 ? lo : hi;
                    const unsigned sel = static_cast<unsigned>(byte < 3 ? byte : byte - 3);
                    uint4 w;
                    unsigned *wp = reinterpret_cast<unsigned *>(&w);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const unsigned t01 = __byte_perm(src[4 * q], src[4 * q + 1], sel | ((4u + sel) << 4));
                        const unsigned t23 = __byte_perm(src[4 * q + 2], src[4 * q + 3], sel | ((4u + sel) << 4));
                        wp[q] = __byte_perm(t01, t23, 0x5410) ^ 0x80808080u;
                    }
                    *reinterpret_cast<uint4 *>(&S.b[s][sl][kb * (kTileN * 16) + g8 * 128 + nloc * 16]) = w;
                }
            }
            fence_proxy_async();
            mbar_arrive(&S.full[s]);
            { const long long c1 = clock64(); t_work += c1 - c0; c0 = c1; }
        }
        if (P.prof && blockIdx.x == 0 && pw == 0 && lane == 0) { P.prof[0] = t_wait; P.prof[1] = t_work; }
    } else if (warp == kEpiWarps) {
        // ===== MMA issuer: one thread ================================================================================
        if (lane == 0) {
            const uint32_t a_base = smem_u32(&S.a[0][0]), b_base = smem_u32(&S.b[0][0][0]);
            long long it = 0, t_wait = 0, t_work = 0, t_wacc = 0, c0 = clock64();
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int s = static_cast<int>(it % kStages), buf = static_cast<int>(it & 1);
                mbar_wait(&S.acc_empty[buf], static_cast<unsigned>(((it >> 1) & 1) ^ 1));
                { const long long c1 = clock64(); t_wacc += c1 - c0; c0 = c1; }
                mbar_wait(&S.full[s], static_cast<unsigned>((it / kStages) & 1));
                { const long long c1 = clock64(); t_wait += c1 - c0; c0 = c1; }
                tc_fence_after();
                // loop order: one A tile (Mr plane j, k-step ks) serves the 7 - j products with y planes i = 0 .. 6 - j, which go to
                // different accumulators g = i + j; the tile is read from shared memory once (collector::a::fill) and reused
                // (::use / ::lastuse), which cuts the A-side shared-memory traffic of a tile from 112 to 28 tile reads
                const uint32_t dbase = tmem + static_cast<uint32_t>(buf * kAccCols);
#pragma unroll
                for (int ks = 0; ks < kK / 32; ++ks) {
#pragma unroll
                    for (int j = 0; j <= kMaxG; ++j) {
                        const uint64_t da = make_desc(a_base + static_cast<uint32_t>(j * kABytes + ks * 2 * (kM * 16)), kM * 16, 128);
#pragma unroll
                        for (int i = 0; i + j <= kMaxG; ++i) {
                            const uint64_t db = make_desc(b_base + static_cast<uint32_t>((s * kSlices + i) * kBBytes + ks * 2 * (kTileN * 16)), kTileN * 16, 128);
                            const uint32_t d = dbase + static_cast<uint32_t>((i + j) * kTileN);
                            // accumulator g = i + j is first written by (ks = 0, j = 0, i = g)
                            const uint32_t acc = (ks == 0 && j == 0) ? 0u : 1u;
                            const int last = kMaxG - j;
                            if (last == 0) tc_mma_i8<0>(d, da, db, kIdesc, acc);
                            else if (i == 0) tc_mma_i8<1>(d, da, db, kIdesc, acc);
                            else if (i == last) tc_mma_i8<3>(d, da, db, kIdesc, acc);
                            else tc_mma_i8<2>(d, da, db, kIdesc, acc);
                        }
                    }
                }
                tc_commit(&S.empty[s]);      // the stage's digit planes may be overwritten once these MMAs have read them
                tc_commit(&S.acc_full[buf]); // ... and the accumulators are complete
                { const long long c1 = clock64(); t_work += c1 - c0; c0 = c1; }
            }
            if (P.prof && blockIdx.x == 0) { P.prof[2] = t_wait; P.prof[3] = t_work; P.prof[6] = t_wacc; }
        }
    } else {
        // ===== epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 = output coordinates ==================================
        const int j = warp * 32 + lane;
        const double rs = S.row_scale[j];
        long long it = 0, t_wait = 0, t_work = 0, c0 = clock64();
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int buf = static_cast<int>(it & 1);
            mbar_wait(&S.acc_full[buf], static_cast<unsigned>((it >> 1) & 1));
            { const long long c1 = clock64(); t_wait += c1 - c0; c0 = c1; }
            tc_fence_after();
            const uint32_t t0 = tmem + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(buf * kAccCols);
#pragma unroll 1
            for (int q = 0; q < kTileN / 8; ++q) {
                uint32_t r[kMaxG + 1][8];
#pragma unroll
                for (int g = 0; g <= kMaxG; ++g) tc_ld8(t0 + static_cast<uint32_t>(g * kTileN + q * 8), r[g]);
                tc_wait_ld();
                double z[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    // base-256 Horner over the 7 accumulators, again without XU conversions (I2F.F64): neighbours are combined
                    // exactly in 64-bit integers (a_g * 256 + a_{g+1}, |.| < 2^34) and turned into doubles by adding the integer to
                    // the bit pattern of 1.5 * 2^52 and subtracting that constant
                    double pr[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        long long v = static_cast<int>(r[2 * t][c]);
                        if (2 * t + 1 <= kMaxG) v = v * 256 + static_cast<int>(r[2 * t + 1][c]);
                        pr[t] = __longlong_as_double(0x4338000000000000LL + v) - 6755399441055744.0;
                    }
                    // groups (0,1) (2,3) (4,5) (6): weights 256^5, 256^3, 256^1, 256^0 relative to group 6
                    double h = fma(pr[0], 65536.0, pr[1]);
                    h = fma(h, 65536.0, pr[2]);
                    h = fma(h, 256.0, pr[3]);
                    z[c] = (h * rs) * S.ind_scale[it & 7][q * 8 + c];
                }
                if (MODE == 0) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const long long ind = tile * kTileN + q * 8 + c;
                        if (ind < P.n && j < D) P.out[ind * D + j] = z[c];
                    }
                } else {
                    const double cj = S.coef[j];
                    double t[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) t[c] = cj * z[c] * z[c];
                    // butterfly over the 32 coordinates of this warp: 8 values per lane -> lane c (c < 8, in bits 4..2) holds
                    // the sum of value c; a fixed tree, so a row's result does not depend on its position in the batch
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const bool up = lane & 16;
                        const double send = up ? t[c] : t[c + 4], keep = up ? t[c + 4] : t[c];
                        t[c] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const bool up = lane & 8;
                        const double send = up ? t[c] : t[c + 2], keep = up ? t[c + 2] : t[c];
                        t[c] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
                    {
                        const bool up = lane & 4;
                        const double send = up ? t[0] : t[1], keep = up ? t[1] : t[0];
                        t[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                    t[0] += __shfl_xor_sync(0xffffffffu, t[0], 2);
                    t[0] += __shfl_xor_sync(0xffffffffu, t[0], 1);
                    // lane bits (4, 3, 2) = (value bit 2, bit 1, bit 0)
                    if ((lane & 3) == 0) {
                        const int c = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                        S.part[buf][warp][q * 8 + c] = t[0];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.acc_empty[buf]);
            if (MODE == 1) {
                asm volatile("bar.sync 1, 128;" ::: "memory"); // the four epilogue warps: partial sums complete
                if (warp == 0) {
                    const long long ind = tile * kTileN + lane;
                    if (ind < P.n) P.out[ind] = ((S.part[buf][0][lane] + S.part[buf][1][lane]) + (S.part[buf][2][lane] + S.part[buf][3][lane])) + P.fbias;
                }
                // part[buf] is written again two tiles later, after another bar.sync of all four warps: no second barrier needed
            }
            { const long long c1 = clock64(); t_work += c1 - c0; c0 = c1; }
        }
        if (P.prof && blockIdx.x == 0 && warp == 0 && lane == 0) { P.prof[4] = t_wait; P.prof[5] = t_work; P.prof[7] = it; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

} // namespace

// ---- probe: how long does ONE tcgen05.mma kind::i8 (M = 128, K = 32) take as a function of N and of the issue pattern? ----------
// pattern 0: all MMAs accumulate into one D; 1: D rotates over 7 accumulators; 2: as 1 with collector::a fill / use / lastuse in
// groups of 7 (one A tile, seven B tiles); 3: as 0 but A and B tiles rotate through shared memory (fresh operands every time)
template <int N, int PATTERN> __global__ void __launch_bounds__(128, 1) mma_probe_kernel(long long *out, int iters)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (64 << 10) / 4; i += blockDim.x) reinterpret_cast<unsigned *>(smem_raw)[i] = 0x01010101u * (i & 3);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem_raw), b0 = a0 + (32 << 10);
        const long long c0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 7; ++u) {
                constexpr int kBT = (32 << 10) / (N * 32); // B tiles that fit in the 32 KB behind the A tiles
                constexpr int kAcc = 512 / N < 7 ? 512 / N : 7;
                const uint32_t aoff = (PATTERN == 3) ? static_cast<uint32_t>(((it * 7 + u) & 7) * 4096) : 0u;
                const uint32_t boff = static_cast<uint32_t>((((PATTERN == 3) ? it * 7 + u : u) % kBT) * (N * 32));
                const uint64_t da = make_desc(a0 + aoff, 128 * 16, 128), db = make_desc(b0 + boff, N * 16, 128);
                const uint32_t d = tmem + ((PATTERN == 1 || PATTERN == 2) ? static_cast<uint32_t>((u % kAcc) * N) : 0u);
                if (PATTERN == 2) {
                    if (u == 0) tc_mma_i8<1>(d, da, db, idesc, 1);
                    else if (u == 6) tc_mma_i8<3>(d, da, db, idesc, 1);
                    else tc_mma_i8<2>(d, da, db, idesc, 1);
                } else
                    tc_mma_i8<0>(d, da, db, idesc, 1);
            }
        }
        const long long c1 = clock64();
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        const long long c2 = clock64();
        out[0] = c1 - c0;
        out[1] = c2 - c0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, int PATTERN> int run_probe(cudaStream_t st, long long *d_out, int iters, double *cyc_issue, double *cyc_done)
{
    PGC_CUDA(cudaFuncSetAttribute(mma_probe_kernel<N, PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 << 10));
    mma_probe_kernel<N, PATTERN><<<1, 128, 64 << 10, st>>>(d_out, iters);
    PGC_CUDA(cudaGetLastError());
    long long h[2];
    PGC_CUDA(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, st));
    PGC_CUDA(cudaStreamSynchronize(st));
    *cyc_issue = static_cast<double>(h[0]) / (7.0 * iters);
    *cyc_done = static_cast<double>(h[1]) / (7.0 * iters);
    return PGC_OK;
}

int mma_probe(pgc_ctx *ctx, double *out /* [4 N values][4 patterns][2] */, cudaStream_t st)
{
    long long *d = nullptr;
    PGC_CUDA(cudaMallocAsync(&d, 16, st));
    const int iters = 400;
    int rc = PGC_OK;
#define PGC_PROBE(ni, NN)                                                                                              \
    if (rc == PGC_OK) rc = run_probe<NN, 0>(st, d, iters, &out[(ni * 4 + 0) * 2], &out[(ni * 4 + 0) * 2 + 1]);          \
    if (rc == PGC_OK) rc = run_probe<NN, 1>(st, d, iters, &out[(ni * 4 + 1) * 2], &out[(ni * 4 + 1) * 2 + 1]);          \
    if (rc == PGC_OK) rc = run_probe<NN, 2>(st, d, iters, &out[(ni * 4 + 2) * 2], &out[(ni * 4 + 2) * 2 + 1]);          \
    if (rc == PGC_OK) rc = run_probe<NN, 3>(st, d, iters, &out[(ni * 4 + 3) * 2], &out[(ni * 4 + 3) * 2 + 1]);
    PGC_PROBE(0, 32)
    PGC_PROBE(1, 64)
    PGC_PROBE(2, 128)
    PGC_PROBE(3, 256)
#undef PGC_PROBE
    cudaFreeAsync(d, st);
    (void)ctx;
    return rc;
}

// debug / measurement entry: MODE 0 (d_coef == nullptr) writes z [n x D], MODE 1 writes f [n] = sum coef_j z_j^2 + fbias
int rot_i8_debug(pgc_ctx *ctx, const double *h_M, int D, const double *h_os, const double *h_coef, double rate, double fbias, const double *d_x,
                 size_t n, double *d_out, int reps, float *ms_per_launch, cudaStream_t st)
{
    PGC_REQUIRE(D >= 2 && D <= kM && D % 2 == 0, "rot_i8: the tensor-core rotation handles even dimensions up to %d (got %d)", kM, D);
    std::vector<int8_t> img;
    std::vector<double> rs;
    build_matrix_planes(h_M, D, nullptr, img, rs);
    int8_t *d_img = nullptr;
    double *d_small = nullptr;
    PGC_CUDA(cudaMallocAsync(&d_img, img.size(), st));
    PGC_CUDA(cudaMallocAsync(&d_small, sizeof(double) * (kM + 2 * D), st));
    PGC_CUDA(cudaMemcpyAsync(d_img, img.data(), img.size(), cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(d_small, rs.data(), sizeof(double) * kM, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(d_small + kM, h_os, sizeof(double) * D, cudaMemcpyHostToDevice, st));
    if (h_coef) PGC_CUDA(cudaMemcpyAsync(d_small + kM + D, h_coef, sizeof(double) * D, cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaStreamSynchronize(st)); // the host vectors go out of scope
    long long *d_prof = nullptr;
    PGC_CUDA(cudaMallocAsync(&d_prof, 8 * sizeof(long long), st));
    PGC_CUDA(cudaMemsetAsync(d_prof, 0, 8 * sizeof(long long), st));
    Params P{d_x, d_img, d_small, d_small + kM, d_small + kM + D, d_out, static_cast<long long>(n), rate, fbias, D, d_prof};
    const size_t smem = sizeof(Smem) + 1024;
    const long long ntiles = (static_cast<long long>(n) + kTileN - 1) / kTileN;
    const unsigned grid = static_cast<unsigned>(std::min<long long>(ntiles, ctx->sm_count));
    if (grid) {
        PGC_CUDA(cudaFuncSetAttribute(rot_i8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        PGC_CUDA(cudaFuncSetAttribute(rot_i8_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        cudaEvent_t e0, e1;
        PGC_CUDA(cudaEventCreate(&e0));
        PGC_CUDA(cudaEventCreate(&e1));
        if (reps < 1) reps = 1;
        for (int r = 0; r < reps + 1; ++r) { // launch 0 warms up; the events bracket the other `reps`
            if (r == 1) PGC_CUDA(cudaEventRecord(e0, st));
            if (r == reps && reps == 1 && false) break;
            if (h_coef) rot_i8_kernel<1><<<grid, kThreads, smem, st>>>(P);
            else rot_i8_kernel<0><<<grid, kThreads, smem, st>>>(P);
            if (reps == 1) break;
        }
        PGC_CUDA(cudaEventRecord(e1, st));
        PGC_CUDA(cudaGetLastError());
        PGC_CUDA(cudaStreamSynchronize(st));
        if (ms_per_launch) {
            float ms = 0.f;
            if (reps > 1) PGC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            *ms_per_launch = reps > 1 ? ms / static_cast<float>(reps) : 0.f;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        ctx->launches.fetch_add(static_cast<uint64_t>(reps > 1 ? reps + 1 : 1), std::memory_order_relaxed);
    }
    if (std::getenv("PGC_ROT_I8_PROF")) {
        long long h[8];
        PGC_CUDA(cudaMemcpyAsync(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost, st));
        PGC_CUDA(cudaStreamSynchronize(st));
        std::fprintf(stderr, "[rot_i8 prof, block 0, last launch] tiles %lld | producer wait %lld work %lld | mma wait(full) %lld wait(acc) %lld issue %lld | "
                             "epilogue wait %lld work %lld (cycles)\n", h[7], h[0], h[1], h[2], h[6], h[3], h[4], h[5]);
    }
    PGC_CUDA(cudaFreeAsync(d_prof, st));
    PGC_CUDA(cudaFreeAsync(d_img, st));
    PGC_CUDA(cudaFreeAsync(d_small, st));
    return PGC_OK;
}

} // namespace i8rot
} // namespace pgc

extern "C" int pgc_debug_mma_i8_probe(pgc_ctx *ctx, double *out32)
{
    using namespace pgc;
    PGC_REQUIRE(ctx && out32, "pgc_debug_mma_i8_probe: null argument");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return i8rot::mma_probe(ctx, out32, ctx->stream);
}

extern "C" int pgc_debug_rot_i8(pgc_ctx *ctx, const double *M, size_t D, const double *os, const double *coef, double rate, double fbias,
                                const double *d_x, size_t n, double *d_out, int reps, float *ms_per_launch, void *stream)
{
    using namespace pgc;
    PGC_REQUIRE(ctx && M && os && d_x && d_out, "pgc_debug_rot_i8: null argument");
    PGC_CUDA(cudaSetDevice(ctx->device));
    return i8rot::rot_i8_debug(ctx, M, static_cast<int>(D), os, coef, rate, fbias, d_x, n, d_out, reps, ms_per_launch, stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}
