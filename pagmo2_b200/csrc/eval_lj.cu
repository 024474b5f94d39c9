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
    long long blocks = (static_cast<long long>(n) + kLjWarps - 1) / kLjWarps;
    int per_sm = 1;
    if (reg) {
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
