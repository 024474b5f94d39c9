// rot_i8.cuh - the FP64-accurate rotation z = Mr * y on the 5th-generation tensor cores (tcgen05.mma kind::i8, TMEM accumulators).
//
// Why: the CEC rotations are 2*D^2 flop per evaluation on the FP64 pipe, which bounds the headline benchmark at ~1e9 evals/s/GPU
// (DESIGN.md 3.1).  sm_100a has no FP64 tensor operand, but it multiplies int8 matrices EXACTLY (int32 accumulation) at 4.5 POPS.
// The Ozaki scheme turns one FP64 product into a few exact integer products: every row of Mr and every decision vector y is
// scaled by its own power of two, rounded to a 54-bit integer and written as 7 balanced base-256 digits d_s in [-128, 127]
// (value = sum_s d_s 256^(6-s)); the integer dot product is sum_{i,j} 256^(12-i-j) <y_i, m_j>, of which the 28 digit pairs with
// i + j <= 6 are kept (the dropped ones are below 2^-56 of |row| * |y|: measured 5e-16 relative to the largest entry of z, the
// same as a plain FP64 dot product, scripts/ozaki_emulation.py).  All pairs with the same i + j share one int32 accumulator
// (<= 7 * 128 * 2^14 < 2^25), so a tile needs 7 accumulators per output.
//
// Mapping (one CTA per SM, persistent, warp-specialised):
//   A operand = the digit planes of Mr, 7 x [128 outputs x 128 k] int8, K-major, resident in shared memory (112 KB);
//   B operand = the digit planes of a tile of 32 decision vectors, 7 x [32 x 128 k] int8 per pipeline stage (3 stages);
//   D         = 7 accumulators [128 lanes = outputs] x [32 columns = individuals] int32 in TMEM, double buffered (448 columns);
//   producer warps  : load x, y = (x - Os) * rate, per-vector scale, digits, canonical no-swizzle core-matrix layout, mbarrier;
//   MMA warp        : one thread issues 28 x 4 tcgen05.mma (K = 32 each) per tile, tcgen05.commit -> mbarriers;
//   epilogue warps  : tcgen05.ld, Horner in base 256 -> double, scales, then the primitive's per-coordinate term and a butterfly
//                     reduction over the coordinates (lane = output coordinate).
#pragma once

#include <cstdint>

namespace pgc
{
namespace i8rot
{

constexpr int kSlices = 7;          // balanced base-256 digits per operand
constexpr int kMaxG = 6;            // digit pairs (i, j) with i + j <= kMaxG are multiplied
constexpr int kM = 128;             // MMA M: output coordinates (TMEM lanes), D <= 128
constexpr int kK = 128;             // padded inner length (4 MMAs of K = 32)
constexpr int kTileN = 32;          // individuals per tile (MMA N)
constexpr int kStages = 3;          // B-operand pipeline depth
constexpr int kABytes = kM * kK;    // one digit plane of Mr
constexpr int kBBytes = kTileN * kK;
constexpr int kScaleBits = 54;      // |v| < 2^e  ->  integer rint(v * 2^(54 - e)), |.| < 2^54: always 7 balanced digits
constexpr int kAccCols = (kMaxG + 1) * kTileN; // TMEM columns of one accumulator set

// canonical K-major, no-swizzle operand layout (UMMA "INTERLEAVE"): core matrix = 8 rows x 16 bytes, contiguous (128 B);
// 8-row groups follow at SBO = 128 B, the next 16 k-values at LBO = rows * 16 B
__host__ __device__ constexpr int a_offset(int row, int k) { return (k / 16) * (kM * 16) + (row / 8) * 128 + (row % 8) * 16 + (k % 16); }
__host__ __device__ constexpr int b_offset(int row, int k) { return (k / 16) * (kTileN * 16) + (row / 8) * 128 + (row % 8) * 16 + (k % 16); }

} // namespace i8rot
} // namespace pgc
