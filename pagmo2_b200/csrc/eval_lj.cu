// eval_lj.cu - Lennard-Jones cluster energy (reference src/problems/lennard_jones.cpp:72-92, coordinate map _r :132-151,
// bounds :99-110) for a batch of decision vectors on sm_100a.
//
// N(N-1)/2 pair terms per individual (11 175 at 150 atoms) of ~14 FP64 flops + one division: FP64-pipe bound
// (arithmetic intensity ~60 flop/B at D = 444).  One warp per individual: the atom coordinates are staged in shared
// memory (SoA), the pair list (i, j) - in the reference's loop order - is linearised and strided over the lanes, lane
// partial sums are combined with shuffles (row i broadcast, 32 consecutive j per step).  Differences to the reference: summation order (lane-strided instead of
// sequential), fused multiply-adds in the distance, and a Newton reciprocal of d*d*d instead of libm pow(d, -3): ~1e-15 relative.  A coincident pair makes the reference assign
// DBL_MAX and keep summing, then multiply by 4 (:81-91): the result is +inf, reproduced explicitly.
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "pgc_internal.cuh"

namespace pgc
{

namespace
{

constexpr int kLjWarps = 8;

// 1/d for a normal, positive d: hardware seed (rcp.approx.ftz.f64, ~2^-20) + one cubic step (three FMAs) - about
// 1 ulp, no special-case branch.  d = 0 gives +inf (the caller flags coincident atoms separately).
__device__ __forceinline__ double fast_rcp(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);  // r (1 + e + e^2) = 1/d (1 - e^3): one cubic step takes the 2^-20 seed below 2^-53
    return fma(r, fma(e, e, e), r);
}

__global__ void __launch_bounds__(kLjWarps * 32) lj_kernel(const double *__restrict__ x, double *__restrict__ f, long long n, int atoms,
                                                           const ushort2 *__restrict__ pairs, int npairs)
{
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *px = smem + static_cast<size_t>(warp) * 3 * atoms, *py = px + atoms, *pz = py + atoms;
    const int D = 3 * atoms - 6;
    for (long long ind = static_cast<long long>(blockIdx.x) * kLjWarps + warp; ind < n; ind += static_cast<long long>(gridDim.x) * kLjWarps) {
        const double *xi = x + ind * D;
        // coordinate map _r, :132-151: atom 0 at the origin, atom 1 on the z axis, atom 2 in the yz plane
        for (int a = lane; a < atoms; a += 32) {
            double cx, cy, cz;
            if (a == 0) {
                cx = cy = cz = 0.0;
            } else if (a == 1) {
                cx = cy = 0.0;
                cz = xi[0];
            } else if (a == 2) {
                cx = 0.0;
                cy = xi[1];
                cz = xi[2];
            } else {
                const double *q = xi + 3 * (a - 2);
                cx = q[0];
                cy = q[1];
                cz = q[2];
            }
            px[a] = cx;
            py[a] = cy;
            pz[a] = cz;
        }
        __syncwarp();
        double s = 0.0;
        bool coincident = false;
        // pairs (i, j > i): atom i is a broadcast read, the lanes take consecutive j (conflict-free), 32 at a time starting
        // at the chunk that contains i + 1.  ~14 FP64-pipe instructions per pair: 3 sub, mul + 2 fma, 2 mul, reciprocal
        // (seed + 2 Newton steps = 4 fma), fma, add; a coincident pair only raises a flag (its inf - inf term is discarded).
        for (int i = 0; i + 1 < atoms; ++i) {
            const double xi0 = px[i], yi0 = py[i], zi0 = pz[i];
#pragma unroll 2
            for (int j0 = (i + 1) & ~31; j0 < atoms; j0 += 32) { // straight-line body: masked lanes redo pair (i, i+1) and drop it
                const int j = j0 + lane;
                const bool valid = j > i && j < atoms;
                const int jj = valid ? j : i + 1;
                const double dx = xi0 - px[jj], dy = yi0 - py[jj], dz = zi0 - pz[jj];
                const double dist = fma(dz, dz, fma(dy, dy, dx * dx)); // rij^2, :78-80
                coincident |= valid && dist == 0.0;
                const double sixth = fast_rcp(dist * dist * dist);     // rij^-6, :84
                const double term = fma(sixth, sixth, -sixth);         // :85
                s += valid ? term : 0.0;
            }
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
        coincident = __any_sync(0xffffffffu, coincident);
        if (lane == 0) f[ind] = coincident ? INFINITY : 4 * s; // :90
        __syncwarp();
    }
}


// Register-tiled variant for clusters of up to 32*NC atoms: lane l keeps atoms l, l+32, ... (NC of them) in registers for the
// whole individual, row atom i is a broadcast read, and the chunks that lie entirely at or below i are skipped by a warp-uniform
// test - no per-pair address arithmetic, no shared-memory traffic for the j side.  ~16 FP64-pipe instructions per pair slot.
template <int NC>
__global__ void __launch_bounds__(kLjWarps * 32) lj_reg_kernel(const double *__restrict__ x, double *__restrict__ f, long long n, int atoms)
{
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *px = smem + static_cast<size_t>(warp) * 3 * atoms, *py = px + atoms, *pz = py + atoms;
    const int D = 3 * atoms - 6;
    for (long long ind = static_cast<long long>(blockIdx.x) * kLjWarps + warp; ind < n; ind += static_cast<long long>(gridDim.x) * kLjWarps) {
        const double *xi = x + ind * D;
        double xj[NC], yj[NC], zj[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) { // coordinate map _r, :132-151; atoms beyond the cluster are parked far away, one per slot
            const int a = c * 32 + lane;
            double cx = 1.0e8 * (a + 1), cy = 0.0, cz = 0.0;
            if (a < atoms) {
                cx = (a >= 3) ? xi[3 * (a - 2)] : 0.0;
                cy = (a >= 3) ? xi[3 * (a - 2) + 1] : (a == 2 ? xi[1] : 0.0);
                cz = (a >= 3) ? xi[3 * (a - 2) + 2] : (a == 2 ? xi[2] : (a == 1 ? xi[0] : 0.0));
                px[a] = cx;
                py[a] = cy;
                pz[a] = cz;
            }
            xj[c] = cx;
            yj[c] = cy;
            zj[c] = cz;
        }
        __syncwarp();
        double s = 0.0;
        bool coincident = false;
        for (int i = 0; i + 1 < atoms; ++i) {
            const double xi0 = px[i], yi0 = py[i], zi0 = pz[i];
            const int cmin = (i + 1) >> 5; // chunks below hold only j <= i
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (c >= cmin) {
                    const int j = c * 32 + lane;
                    const bool valid = j > i && j < atoms;
                    const double dx = xi0 - xj[c], dy = yi0 - yj[c], dz = zi0 - zj[c];
                    const double dist = fma(dz, dz, fma(dy, dy, dx * dx)); // rij^2, :78-80
                    coincident |= valid && dist == 0.0;
                    const double sixth = fast_rcp(dist * dist * dist);     // rij^-6, :84
                    const double term = fma(sixth, sixth, -sixth);         // :85
                    s += valid ? term : 0.0;
                }
            }
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
        coincident = __any_sync(0xffffffffu, coincident);
        if (lane == 0) f[ind] = coincident ? INFINITY : 4 * s; // :90
        __syncwarp();
    }
}

// Circulant pair schedule: atom i meets atoms i + 1 ... i + (N - 1) / 2 (indices mod N; even N adds the offset N / 2 for the first
// N / 2 atoms), so every unordered pair is computed once and, unlike the triangular i < j loop, no lane idles on the diagonal:
// slot efficiency N / (32 NC) instead of ~75 %.  Lane l keeps its own atoms l, l + 32, ... (NC chunks) in registers.  The partner
// side is read from a wrapped copy of the coordinates in shared memory (position k holds atom k mod N): at step dp = 1 ... N the
// lane loads the ONE atom at position l + dp and uses it against every own chunk c whose offset (dp - 32 c) mod N lies in
// 1 ... (N - 1) / 2 - two or three chunks at a time - so each individual costs N loads per lane instead of one per pair (shared
// memory moves 128 B/clk/SM: at one 3 x 8 B load per pair it would run at 92 % of the FP64 pipe's time).  The live chunks are a
// cyclic range [LO, LO + A) of the chunk ring that changes ~2 NC times per individual (the segment table is built once per
// block); each (LO, A) is its own unrolled instance so the chunk arrays stay in registers, and the body is written stage by
// stage across the live chunks so that their dependent FP64 chains (15 deep) interleave.
// 13 FP64-pipe instructions per pair: 3 sub, mul + 2 fma, 2 mul, 3 fma (cubic step on the reciprocal seed), fma, add.
#ifndef PGC_LJ_CHAINS
#define PGC_LJ_CHAINS 6
#endif
#ifndef PGC_LJ_MINBLOCKS
#define PGC_LJ_MINBLOCKS 1
#endif

template <int NC, int LO, int A>
__device__ __forceinline__ void lj_span(const double *__restrict__ pos, int k0, int k1, const double (&xa)[NC], const double (&ya)[NC],
                                        const double (&za)[NC], double (&s)[NC])
{
    constexpr int U = (PGC_LJ_CHAINS + A - 1) / A; // ~PGC_LJ_CHAINS independent chains per warp
#pragma unroll U
    for (int k = k0; k < k1; ++k) {
        const double qx = pos[3 * k], qy = pos[3 * k + 1], qz = pos[3 * k + 2]; // one address register, immediate offsets
        double dist[A], cube[A], r[A], e[A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const int c = (LO + a) % NC;
            const double dx = xa[c] - qx, dy = ya[c] - qy, dz = za[c] - qz;
            dist[a] = fma(dz, dz, fma(dy, dy, dx * dx)); // rij^2, :78-80
        }
#pragma unroll
        for (int a = 0; a < A; ++a) cube[a] = dist[a] * dist[a] * dist[a];
#pragma unroll
        for (int a = 0; a < A; ++a) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[a]) : "d"(cube[a]));
#pragma unroll
        for (int a = 0; a < A; ++a) e[a] = fma(-cube[a], r[a], 1.0);
#pragma unroll
        for (int a = 0; a < A; ++a) e[a] = fma(e[a], e[a], e[a]);
#pragma unroll
        for (int a = 0; a < A; ++a) r[a] = fma(r[a], e[a], r[a]);     // rij^-6, :84
#pragma unroll
        for (int a = 0; a < A; ++a) e[a] = fma(r[a], r[a], -r[a]);    // :85
        // a coincident pair gives 1 / 0 = inf and inf * inf - inf = NaN, which poisons the lane's sum: detected once at the end
        // instead of one FP64 compare per pair
#pragma unroll
        for (int a = 0; a < A; ++a) s[(LO + a) % NC] += e[a]; // (parked lanes add ~1e-48 terms to sums that are dropped at the end)
    }
}

__host__ __device__ constexpr int lj_max_live(int nc) { return nc / 2 + 1; } // most own chunks one partner position can serve (checked for every N)

template <int NC, int LO, int A = 1>
__device__ __forceinline__ void lj_dispatch_count(int cnt, const double *__restrict__ pos, int k0, int k1, const double (&xa)[NC],
                                                  const double (&ya)[NC], const double (&za)[NC], double (&s)[NC])
{
    if (cnt == A) {
        lj_span<NC, LO, A>(pos, k0, k1, xa, ya, za, s);
        return;
    }
    if constexpr (A < lj_max_live(NC) && A < NC) lj_dispatch_count<NC, LO, A + 1>(cnt, pos, k0, k1, xa, ya, za, s);
}

template <int NC, int LO = 0>
__device__ __forceinline__ void lj_dispatch(int lo, int cnt, const double *__restrict__ pos, int k0, int k1, const double (&xa)[NC],
                                            const double (&ya)[NC], const double (&za)[NC], double (&s)[NC])
{
    if (lo == LO) {
        lj_dispatch_count<NC, LO>(cnt, pos, k0, k1, xa, ya, za, s);
        return;
    }
    if constexpr (LO + 1 < NC) lj_dispatch<NC, LO + 1>(lo, cnt, pos, k0, k1, xa, ya, za, s);
}

constexpr int kLjMaxSegments = 40; // the live set changes at most 2 NC + 1 times; gaps (clusters below 64 atoms) add NC more

// entries of the wrapped coordinate copy one warp needs: positions up to 31 + N
__host__ __device__ inline int lj_circ_len(int atoms) { return atoms + 32; }

template <int NC>
__global__ void __launch_bounds__(kLjWarps * 32, PGC_LJ_MINBLOCKS) lj_circ_kernel(const double *__restrict__ x, double *__restrict__ f, long long n, int atoms)
{
    extern __shared__ double smem[];
    __shared__ int seg_k0[kLjMaxSegments], seg_lo[kLjMaxSegments], seg_cnt[kLjMaxSegments], n_seg;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = atoms, W = lj_circ_len(N);
    double *pos = smem + static_cast<size_t>(warp) * 3 * W; // (x, y, z) of position k at pos[3 k ...]: 24-byte stride, conflict free
    const int D = 3 * atoms - 6;
    const int full = (N - 1) / 2;          // offsets at which every atom has a partner
    const bool half = (N % 2) == 0;        // even N: one more offset for the first N / 2 atoms
    if (threadIdx.x == 0) { // segments of dp = 1 ... N with a constant set of live chunks: c is live when (dp - 32 c) mod N is in [1, full]
        int ns = 0, prev_lo = -1, prev_cnt = -1;
        for (int dp = 1; dp <= N; ++dp) {
            unsigned mask = 0;
            for (int c = 0; c < NC; ++c) {
                const int off = ((dp - 32 * c) % N + N) % N;
                if (off >= 1 && off <= full) mask |= 1u << c;
            }
            const int cnt = __popc(mask);
            int lo = 0; // first chunk of the cyclic run: live, predecessor not live
            for (int c = 0; c < NC; ++c)
                if ((mask >> c & 1u) && !(mask >> ((c + NC - 1) % NC) & 1u)) lo = c;
            if (lo != prev_lo || cnt != prev_cnt) {
                seg_k0[ns] = dp;
                seg_lo[ns] = lo;
                seg_cnt[ns] = cnt;
                ++ns;
                prev_lo = lo;
                prev_cnt = cnt;
            }
        }
        seg_k0[ns] = N + 1;
        n_seg = ns;
    }
    __syncthreads();
    const int nseg = n_seg;
    for (long long ind = static_cast<long long>(blockIdx.x) * kLjWarps + warp; ind < n; ind += static_cast<long long>(gridDim.x) * kLjWarps) {
        const double *xi = x + ind * D;
        double xa[NC], ya[NC], za[NC];
        for (int a = 2 * N + lane; a < W; a += 32) pos[3 * a] = pos[3 * a + 1] = pos[3 * a + 2] = 0.0; // clusters below 32 atoms: read by parked lanes only
#pragma unroll
        for (int c = 0; c < NC; ++c) { // coordinate map _r, :132-151; lanes beyond the cluster are parked far away (their terms are dropped)
            const int a = c * 32 + lane;
            double cx = 1.0e8 * (a + 1), cy = 0.0, cz = 0.0;
            if (a < N) {
                cx = (a >= 3) ? xi[3 * (a - 2)] : 0.0;
                cy = (a >= 3) ? xi[3 * (a - 2) + 1] : (a == 2 ? xi[1] : 0.0);
                cz = (a >= 3) ? xi[3 * (a - 2) + 2] : (a == 2 ? xi[2] : (a == 1 ? xi[0] : 0.0));
                pos[3 * a] = cx;
                pos[3 * a + 1] = cy;
                pos[3 * a + 2] = cz;
                if (a + N < W) { // wrapped copy: position k holds atom k mod N
                    pos[3 * (a + N)] = cx;
                    pos[3 * (a + N) + 1] = cy;
                    pos[3 * (a + N) + 2] = cz;
                }
            }
            xa[c] = cx;
            ya[c] = cy;
            za[c] = cz;
        }
        __syncwarp();
        double s[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) s[c] = 0.0;
        for (int g = 0; g < nseg; ++g) {
            const int cnt = seg_cnt[g];
            if (cnt > 0) lj_dispatch<NC>(seg_lo[g], cnt, pos + 3 * lane, seg_k0[g], seg_k0[g + 1], xa, ya, za, s);
        }
        if (half) {
            const int d = N / 2;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (c * 32 < d) { // warp-uniform: chunks that hold atoms below N / 2
                    const int i = c * 32 + lane, j = i + d;
                    const bool valid = i < d;
                    const double dx = xa[c] - pos[3 * j], dy = ya[c] - pos[3 * j + 1], dz = za[c] - pos[3 * j + 2];
                    const double dist = fma(dz, dz, fma(dy, dy, dx * dx));
                    const double sixth = fast_rcp(dist * dist * dist);
                    const double term = fma(sixth, sixth, -sixth);
                    s[c] += valid ? term : 0.0;
                }
            }
        }
        if ((NC - 1) * 32 + lane >= N) s[NC - 1] = 0.0; // parked lanes of the last chunk: their (tiny) terms are dropped here
        double tot = s[0];
#pragma unroll
        for (int c = 1; c < NC; ++c) tot += s[c];
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, m);
        // coincident atoms: the reference assigns DBL_MAX to the pair and the final 4 * v overflows to +inf (:81-91)
        if (lane == 0) f[ind] = (tot != tot || isinf(tot)) ? INFINITY : 4 * tot; // :90
        __syncwarp();
    }
}

} // namespace

int lj_create(pgc_problem *p)
{
    const unsigned atoms = p->desc.dim;
    PGC_REQUIRE(atoms >= 3u,
                "The number of atoms in a Lennard Jones Clusters problem must be positive and greater than 2, while a number of %u "
                "was detected.",
                atoms); // lennard_jones.cpp:52-56
    PGC_REQUIRE(atoms <= 65535u, "Lennard-Jones device evaluator: at most 65535 atoms, %u requested", atoms);
    const size_t D = 3u * atoms - 6u;
    p->nx = D;
    p->nobj = 1;
    p->lb.assign(D, -3.0); // :99-110
    p->ub.assign(D, 3.0);
    for (size_t i = 0; i < D; ++i)
        if (i != 0 && i % 3 == 0) {
            p->lb[i] = 0.0;
            p->ub[i] = 6.0;
        }
    p->name = "Lennard Jones Cluster (" + std::to_string(atoms) + " atoms)"; // :113-116
    std::vector<ushort2> pairs;
    pairs.reserve(static_cast<size_t>(atoms) * (atoms - 1) / 2);
    for (unsigned i = 0; i + 1 < atoms; ++i)
        for (unsigned j = i + 1; j < atoms; ++j) pairs.push_back(make_ushort2(static_cast<unsigned short>(i), static_cast<unsigned short>(j)));
    PGC_CUDA(cudaSetDevice(p->ctx->device));
    PGC_CUDA(cudaMalloc(&p->d_shuffle, sizeof(ushort2) * pairs.size())); // reuses the int table slot
    PGC_CUDA(cudaMemcpy(p->d_shuffle, pairs.data(), sizeof(ushort2) * pairs.size(), cudaMemcpyHostToDevice));
    const double np = static_cast<double>(pairs.size());
    p->flops_per_eval = 14.0 * np; // 3 sub, 3 mul, 2 add, 2 mul, 1 div, 1 mul, 1 sub, 1 add
    p->transc_per_eval = 0;
    return PGC_OK;
}

int lj_eval(pgc_problem *p, const double *d_dvs, size_t n, double *d_fvs, cudaStream_t stream)
{
    if (n == 0) return PGC_OK;
    const int atoms = static_cast<int>(p->desc.dim);
    const int npairs = atoms * (atoms - 1) / 2;
    const size_t smem = sizeof(double) * 3 * atoms * kLjWarps;
    PGC_REQUIRE(smem <= 200 * 1024, "Lennard-Jones device evaluator: %d atoms do not fit the shared-memory tile", atoms);
    using reg_fn = void (*)(const double *, double *, long long, int);
    reg_fn reg = nullptr;
    switch ((atoms + 31) / 32) { // register-tiled kernel up to 256 atoms
        case 1: reg = lj_reg_kernel<1>; break;
        case 2: reg = lj_reg_kernel<2>; break;
        case 3: reg = lj_reg_kernel<3>; break;
        case 4: reg = lj_reg_kernel<4>; break;
        case 5: reg = lj_reg_kernel<5>; break;
        case 6: reg = lj_reg_kernel<6>; break;
        case 7: reg = lj_reg_kernel<7>; break;
        case 8: reg = lj_reg_kernel<8>; break;
        default: break;
    }
    reg_fn circ = nullptr;
    switch ((atoms + 31) / 32) { // circulant schedule (default) up to 256 atoms; PGC_LJ_ROWS=1 selects the row-by-row kernel (A/B)
        case 1: circ = lj_circ_kernel<1>; break;
        case 2: circ = lj_circ_kernel<2>; break;
        case 3: circ = lj_circ_kernel<3>; break;
        case 4: circ = lj_circ_kernel<4>; break;
        case 5: circ = lj_circ_kernel<5>; break;
        case 6: circ = lj_circ_kernel<6>; break;
        case 7: circ = lj_circ_kernel<7>; break;
        case 8: circ = lj_circ_kernel<8>; break;
        default: break;
    }
    const char *rows_env = std::getenv("PGC_LJ_ROWS");
    if (rows_env && rows_env[0] == '1') circ = nullptr;
    long long blocks = (static_cast<long long>(n) + kLjWarps - 1) / kLjWarps;
    int per_sm = 1;
    if (circ) {
        const size_t smem_c = sizeof(double) * 3 * static_cast<size_t>(lj_circ_len(atoms)) * kLjWarps;
        PGC_CUDA(cudaFuncSetAttribute(circ, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_c)));
        PGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, circ, kLjWarps * 32, smem_c));
        blocks = std::min<long long>(blocks, static_cast<long long>(p->ctx->sm_count) * std::max(per_sm, 1));
        circ<<<static_cast<unsigned>(blocks), kLjWarps * 32, smem_c, stream>>>(d_dvs, d_fvs, static_cast<long long>(n), atoms);
    } else if (reg) {
        PGC_CUDA(cudaFuncSetAttribute(reg, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        PGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reg, kLjWarps * 32, smem));
        blocks = std::min<long long>(blocks, static_cast<long long>(p->ctx->sm_count) * std::max(per_sm, 1));
        reg<<<static_cast<unsigned>(blocks), kLjWarps * 32, smem, stream>>>(d_dvs, d_fvs, static_cast<long long>(n), atoms);
    } else {
        PGC_CUDA(cudaFuncSetAttribute(lj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        PGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lj_kernel, kLjWarps * 32, smem));
        blocks = std::min<long long>(blocks, static_cast<long long>(p->ctx->sm_count) * std::max(per_sm, 1));
        lj_kernel<<<static_cast<unsigned>(blocks), kLjWarps * 32, smem, stream>>>(d_dvs, d_fvs, static_cast<long long>(n), atoms,
                                                                                  reinterpret_cast<const ushort2 *>(p->d_shuffle), npairs);
    }
    PGC_CUDA(cudaGetLastError());
    p->ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return PGC_OK;
}

} // namespace pgc
