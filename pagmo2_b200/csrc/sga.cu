// sga.cu - pagmo's simple genetic algorithm as a generational device loop (sm_100a).
//
// Replaces reference sga::evolve (src/algorithms/sga.cpp:184-292): perform_selection :341-383, perform_crossover :385-457
// (sbx via genetic_operators.cpp:71-144), perform_mutation :459-547, reinsertion = best NP of children + parents :282-289.
// One thread builds one offspring (one pair for sbx) from Philox streams, children are evaluated in one batch, the survivor
// sort is a stable radix sort on order-preserving keys (NaN last, detail::less_than_f).  Continuous decision vectors only
// (every UDP of the device path has nix = 0).
// How the reference's sequential random logic is restated, draw for draw, per individual (stream (seed, kTagSga, generation, .)):
//   * tournament selection: the reference keeps ONE index array that it partially Fisher-Yates-shuffles for every offspring
//     (:359-368); each offspring still sees a uniformly random subset of size param_s, so every offspring j shuffles its own
//     (virtual) identity array here - index j of the stream;
//   * mating partner (:411-412): after swap(all_idx[0], all_idx[i]) position p >= 1 holds p - 1 if p <= i, else p; so the partner
//     of i is that function of one uniform integer in [1, NP-1] - exactly the reference's value;
//   * mutation (:489-491): a uniform shuffle and N ~ Binomial(dim, m) mutated genes == every gene mutated independently with
//     probability m; one Bernoulli draw per gene here;
//   * sbx shuffles the selected individuals (:400) - a Philox-keyed permutation here (stable argsort of keys).
// Integer results of the survivor selection are exact; FP64 genes follow the reference's expressions (compiled with -fmad=false).
#include <cmath>
#include <vector>

#include "pgc_internal.cuh"
#include "philox.cuh"

namespace pgc
{

namespace
{

enum { kCrossExp = 0, kCrossBin = 1, kCrossSingle = 2, kCrossSbx = 3 };
enum { kMutGaussian = 0, kMutUniform = 1, kMutPolynomial = 2 };
enum { kSelTournament = 0, kSelTruncated = 1 };

struct SgaParams {
    const double *x, *f;      // parents [NP x nx], [NP]
    const unsigned *sel;      // selected parent of offspring slot j
    const unsigned *perm;     // sbx: shuffle of the slots
    const unsigned *order;    // truncated selection: indices by ascending fitness
    const double *lb, *ub;
    double *xnew;             // offspring [NP x nx]
    unsigned *sel_out;
    unsigned NP, nx;
    double cr, eta_c, m, param_m;
    unsigned param_s, crossover, mutation, selection;
    unsigned long long seed;
    unsigned generation;
};

__global__ void sga_selection_kernel(const SgaParams P)
{
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.NP) return;
    if (P.selection == kSelTruncated) { // :347-354
        P.sel_out[j] = P.order[j % P.param_s];
        return;
    }
    // :356-379: param_s distinct individuals (partial Fisher-Yates on a virtual identity array), the first minimum wins
    PhiloxStream rs(P.seed, kTagSga, P.generation, j);
    unsigned pos[16], val[16]; // swapped entries of the virtual array (param_s is small; beyond 16 picks fall back to rescans)
    unsigned nsw = 0, winner = 0;
    double fw = 0.0;
    for (unsigned i = 0; i < P.param_s; ++i) {
        unsigned index = i + static_cast<unsigned>(rs.next() * static_cast<double>(P.NP - i));
        if (index >= P.NP) index = P.NP - 1;
        // value currently at `index` and at `i`
        unsigned vi = i, vx = index;
        for (unsigned s = 0; s < nsw; ++s) {
            if (pos[s] == index) vx = val[s];
            if (pos[s] == i) vi = val[s];
        }
        // swap: position i now holds vx (never touched again), position index holds vi
        bool found = false;
        for (unsigned s = 0; s < nsw; ++s)
            if (pos[s] == index) {
                val[s] = vi;
                found = true;
            }
        if (!found && nsw < 16) {
            pos[nsw] = index;
            val[nsw] = vi;
            ++nsw;
        }
        const double fv = P.f[vx];
        if (i == 0 || fv < fw) {
            winner = vx;
            fw = fv;
        }
    }
    P.sel_out[j] = winner;
}

__device__ __forceinline__ double sga_betaq(double beta, double eta_c, double rand01) // genetic_operators.cpp:49-57
{
    const double alpha = 2. - pow(beta, -(eta_c + 1.));
    if (rand01 < (1. / alpha)) return pow(rand01 * alpha, 1. / (eta_c + 1.));
    return pow(1. / (2. - rand01 * alpha), 1. / (eta_c + 1.));
}

__device__ void sga_mutate(double *c, const SgaParams &P, PhiloxStream &rs) // :493-545 + force_bounds_stick
{
    for (unsigned g = 0; g < P.nx; ++g) {
        const double lb = P.lb[g], ub = P.ub[g];
        if (rs.next() < P.m) {
            if (P.mutation == kMutUniform) {
                c[g] = (lb == ub) ? lb : (ub - lb) * rs.next() + lb;
            } else if (P.mutation == kMutGaussian) {
                const double sd = (ub - lb) * P.param_m;
                const double u1 = 1.0 - rs.next(), u2 = rs.next();
                c[g] += (sqrt(-2.0 * log(u1)) * cos(2.0 * 3.141592653589793238462643383279502884 * u2)) * sd;
            } else {
                const double u = rs.next();
                if (u <= 0.5) {
                    const double delta_l = pow(2. * u, 1. / (1. + P.param_m)) - 1.;
                    c[g] += delta_l * (c[g] - lb);
                } else {
                    const double delta_r = 1 - pow(2. * (1. - u), 1. / (1. + P.param_m));
                    c[g] += delta_r * (ub - c[g]);
                }
            }
        }
        if (c[g] < lb) c[g] = lb;
        if (c[g] > ub) c[g] = ub;
    }
}

__global__ void sga_variation_kernel(const SgaParams P)
{
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned nx = P.nx;
    if (P.crossover == kCrossSbx) { // :398-405: pairs of the shuffled selection
        if (t >= P.NP / 2) return;
        PhiloxStream rs(P.seed, kTagSga, P.generation, P.NP + t);
        const double *p1 = P.x + static_cast<size_t>(P.sel[P.perm[2 * t]]) * nx, *p2 = P.x + static_cast<size_t>(P.sel[P.perm[2 * t + 1]]) * nx;
        double *c1 = P.xnew + static_cast<size_t>(2 * t) * nx, *c2 = c1 + nx;
        for (unsigned i = 0; i < nx; ++i) {
            c1[i] = p1[i];
            c2[i] = p2[i];
        }
        if (rs.next() < P.cr) { // genetic_operators.cpp:91-131
            for (unsigned i = 0; i < nx; ++i) {
                const double a = p1[i], b = p2[i], yl = P.lb[i], yu = P.ub[i];
                if ((rs.next() < 0.5) && (fabs(a - b)) > 1e-14 && yl != yu) {
                    const double y1 = (a < b) ? a : b, y2 = (a < b) ? b : a;
                    const double rand01 = rs.next();
                    double beta = 1. + (2. * (y1 - yl) / (y2 - y1));
                    double betaq = sga_betaq(beta, P.eta_c, rand01);
                    double v1 = 0.5 * ((y1 + y2) - betaq * (y2 - y1));
                    beta = 1. + (2. * (yu - y2) / (y2 - y1));
                    betaq = sga_betaq(beta, P.eta_c, rand01);
                    double v2 = 0.5 * ((y1 + y2) + betaq * (y2 - y1));
                    if (v1 < yl) v1 = yl;
                    if (v2 < yl) v2 = yl;
                    if (v1 > yu) v1 = yu;
                    if (v2 > yu) v2 = yu;
                    if (rs.next() < .5) {
                        c1[i] = v1;
                        c2[i] = v2;
                    } else {
                        c1[i] = v2;
                        c2[i] = v1;
                    }
                }
            }
        }
        sga_mutate(c1, P, rs);
        sga_mutate(c2, P, rs);
        return;
    }
    if (t >= P.NP) return;
    PhiloxStream rs(P.seed, kTagSga, P.generation, P.NP + t);
    double *child = P.xnew + static_cast<size_t>(t) * nx;
    const double *own = P.x + static_cast<size_t>(P.sel[t]) * nx;
    for (unsigned g = 0; g < nx; ++g) child[g] = own[g];
    // :411-416 mating partner among the other slots
    unsigned pidx = 1 + static_cast<unsigned>(rs.next() * static_cast<double>(P.NP - 1));
    if (pidx > P.NP - 1) pidx = P.NP - 1;
    const unsigned partner = (pidx <= t) ? pidx - 1 : pidx;
    const double *parent2 = P.x + static_cast<size_t>(P.sel[partner]) * nx;
    auto gene = [&]() {
        unsigned n = static_cast<unsigned>(rs.next() * static_cast<double>(nx));
        return n < nx ? n : nx - 1;
    };
    if (P.crossover == kCrossExp) { // :419-427
        unsigned n = gene(), L = 0;
        do {
            child[n] = parent2[n];
            n = (n + 1u) % nx;
            ++L;
        } while ((rs.next() < P.cr) && (L < nx));
    } else if (P.crossover == kCrossBin) { // :429-437
        unsigned n = gene();
        for (unsigned L = 0; L < nx; ++L) {
            if ((rs.next() < P.cr) || L + 1 == nx) child[n] = parent2[n];
            n = (n + 1) % nx;
        }
    } else { // single point :439-446
        if (rs.next() < P.cr) {
            const unsigned n = gene();
            for (unsigned g = n; g < nx; ++g) child[g] = parent2[g];
        }
    }
    sga_mutate(child, P, rs);
}

__global__ void gather2_kernel(const unsigned *sel, unsigned rows, unsigned n0, unsigned nx, const double *xA, const double *fA, const double *xB,
                               const double *fB, double *x_out, double *f_out)
{
    const unsigned r = blockIdx.x;
    if (r >= rows) return;
    const unsigned s = sel[r];
    const bool a = s < n0;
    const unsigned src = a ? s : s - n0;
    const double *x = (a ? xA : xB) + static_cast<size_t>(src) * nx;
    for (unsigned j = threadIdx.x; j < nx; j += blockDim.x) x_out[static_cast<size_t>(r) * nx + j] = x[j];
    if (threadIdx.x == 0) f_out[r] = (a ? fA : fB)[src];
}

} // namespace

int sga_evolve_device(pgc_problem *prob, double *d_x, double *d_f, unsigned NP, unsigned gens, double cr, double eta_c, double m,
                      double param_m, unsigned param_s, unsigned crossover, unsigned mutation, unsigned selection, unsigned long long seed,
                      unsigned first_generation, int (*eval)(pgc_problem *, const double *, size_t, double *, cudaStream_t), cudaStream_t st)
{
    // constructor checks, sga.cpp:116-165
    PGC_REQUIRE(cr <= 1. && cr >= 0., "The crossover probability must be in the [0,1] range, while a value of %g was detected", cr);
    PGC_REQUIRE(eta_c >= 1. && eta_c <= 100., "The distribution index for SBX crossover must be in [1, 100], while a value of %g was detected", eta_c);
    PGC_REQUIRE(m >= 0. && m <= 1., "The mutation probability must be in the [0,1] range, while a value of %g was detected", m);
    PGC_REQUIRE(param_s >= 1u, "The selection parameter must be at least 1, while a value of %u was detected", param_s);
    PGC_REQUIRE(mutation <= 2u, "The mutation type must either be \"gaussian\" or \"uniform\" or \"polynomial\": unknown type requested: %u", mutation);
    PGC_REQUIRE(selection <= 1u, "The selection type must either be \"roulette\" or \"truncated\" or \"tournament\": unknown type requested: %u",
                selection);
    PGC_REQUIRE(crossover <= 3u,
                "The crossover type must either be \"exponential\" or \"binomial\" or \"sbx\" or \"single\": unknown type requested: %u", crossover);
    if (mutation == kMutPolynomial)
        PGC_REQUIRE(param_m >= 1. && param_m <= 100.,
                    "Polynomial mutation was selected, the mutation parameter (distribution index) must be in [1, 100], while a value of %g was "
                    "detected",
                    param_m);
    else
        PGC_REQUIRE(param_m >= 0. && param_m <= 1., "The mutation parameter must be in [0,1], while a value of %g was detected", param_m);
    // evolve() checks, :194-216
    PGC_REQUIRE(prob->nobj == 1, "Multiple objectives detected in %s instance. SGA: Genetic Algorithm cannot deal with them", prob->name.c_str());
    PGC_REQUIRE(NP >= 2u, "%s needs at least 2 individuals in the population, %u detected", prob->name.c_str(), NP);
    PGC_REQUIRE(param_s <= NP, "The parameter for selection must be smaller than the population size, while a value of: %u was detected in a "
                               "population of size: %u", param_s, NP);
    PGC_REQUIRE(!(crossover == kCrossSbx && NP % 2u), "Population size must be even if sbx crossover is selected. Detected pop size is: %u", NP);
    PGC_REQUIRE(selection == kSelTruncated || param_s <= 16u, "sga on the device: tournament sizes above 16 are not implemented (got %u)", param_s);
    if (gens == 0) return PGC_OK;
    pgc_ctx *ctx = prob->ctx;
    const unsigned nx = static_cast<unsigned>(prob->nx);
    double *d_b = nullptr, *xnew = nullptr, *fboth = nullptr, *xo = nullptr, *fo = nullptr;
    unsigned *sel = nullptr, *perm = nullptr, *order = nullptr, *keep = nullptr;
    StreamScratch scratch(st); // released in stream order on every path out of this function
    PGC_CUDA(scratch.get(&d_b, 2 * nx * sizeof(double)));
    PGC_CUDA(scratch.get(&xnew, sizeof(double) * NP * nx));
    PGC_CUDA(scratch.get(&fboth, sizeof(double) * 2 * NP));
    PGC_CUDA(scratch.get(&xo, sizeof(double) * NP * nx));
    PGC_CUDA(scratch.get(&fo, sizeof(double) * NP));
    PGC_CUDA(scratch.get(&sel, sizeof(unsigned) * NP));
    PGC_CUDA(scratch.get(&perm, sizeof(unsigned) * NP));
    PGC_CUDA(scratch.get(&order, sizeof(unsigned) * NP));
    PGC_CUDA(scratch.get(&keep, sizeof(unsigned) * NP));
    PGC_CUDA(cudaMemcpyAsync(d_b, prob->lb.data(), nx * sizeof(double), cudaMemcpyHostToDevice, st));
    PGC_CUDA(cudaMemcpyAsync(d_b + nx, prob->ub.data(), nx * sizeof(double), cudaMemcpyHostToDevice, st));
    SgaParams P{};
    P.lb = d_b;
    P.ub = d_b + nx;
    P.NP = NP;
    P.nx = nx;
    P.cr = cr;
    P.eta_c = eta_c;
    P.m = m;
    P.param_m = param_m;
    P.param_s = param_s;
    P.crossover = crossover;
    P.mutation = mutation;
    P.selection = selection;
    P.seed = seed;
    int rc = PGC_OK;
    for (unsigned g = 0; g < gens && rc == PGC_OK; ++g) {
        P.generation = first_generation + g;
        P.x = d_x;
        P.f = d_f;
        P.xnew = xnew;
        P.sel = sel;
        P.sel_out = sel;
        P.perm = perm;
        P.order = order;
        if (selection == kSelTruncated && (rc = so_best_indices_device(ctx, d_f, NP, param_s, order, st))) break;
        sga_selection_kernel<<<(NP + 127) / 128, 128, 0, st>>>(P);
        if (crossover == kCrossSbx && (rc = philox_permutation_device(ctx, NP, seed, kTagShuffle1, P.generation, perm, st))) break;
        const unsigned work = crossover == kCrossSbx ? NP / 2 : NP;
        sga_variation_kernel<<<(work + 127) / 128, 128, 0, st>>>(P);
        if (cudaGetLastError() != cudaSuccess) {
            rc = PGC_ERR_CUDA;
            break;
        }
        ctx->launches.fetch_add(2, std::memory_order_relaxed);
        // children's fitness into the first half of the pool, parents' into the second (:275-281)
        if ((rc = eval(prob, xnew, NP, fboth, st))) break;
        if ((rc = log_sga_device(ctx, d_f, fboth, NP, g + 1u, static_cast<unsigned long long>(g + 1u) * NP, st))) break; // sga.cpp:252-274
        if (cudaMemcpyAsync(fboth + NP, d_f, sizeof(double) * NP, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            rc = PGC_ERR_CUDA;
            break;
        }
        if ((rc = so_best_indices_device(ctx, fboth, 2 * NP, NP, keep, st))) break;
        gather2_kernel<<<NP, 64, 0, st>>>(keep, NP, NP, nx, xnew, fboth, d_x, d_f, xo, fo);
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaMemcpyAsync(d_x, xo, sizeof(double) * NP * nx, cudaMemcpyDeviceToDevice, st) != cudaSuccess
            || cudaMemcpyAsync(d_f, fo, sizeof(double) * NP, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            rc = PGC_ERR_CUDA;
            break;
        }
    }
    cudaStreamSynchronize(st); // lb/ub staging came from pageable host vectors; the scratch is released by its destructor
    if (rc == PGC_ERR_CUDA) set_error("sga_evolve_device: CUDA failure: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

} // namespace pgc
