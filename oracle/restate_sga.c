/* oracle/restate_sga.c - plain-C restatement of pagmo::sga::evolve as the generational loop the device runs.  TEST INFRASTRUCTURE ONLY.
 * Follows reference src/algorithms/sga.cpp: evolve :184-292, perform_selection :341-383, perform_crossover :385-457 (sbx:
 * src/utils/genetic_operators.cpp:49-57,71-144), perform_mutation :459-547, force_bounds_stick src/utils/generic.cpp:171-183,
 * reinsertion (best NP of children followed by parents, stable) :275-289.
 * PARITY UNPINNED for the random stream (the reference draws from one sequential mt19937 and its tests only check determinism,
 * tests/sga.cpp): every draw comes from the Philox substream (seed, TAG_SGA, generation, index) in the reference's per-individual
 * order - index j for the tournament of offspring j, NP + i for the variation of offspring (or sbx pair) i.  Where the reference
 * threads one mutable index array through all individuals (tournament shuffle :359-368, mutation shuffle + binomial count
 * :489-491) the restatement uses the per-individual equivalent described in pagmo2_b200/csrc/sga.cu; the mating partner :411-412 is
 * reproduced exactly.  Continuous decision vectors only.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "philox.h"

enum { X_EXP = 0, X_BIN = 1, X_SINGLE = 2, X_SBX = 3 };
enum { M_GAUSS = 0, M_UNIFORM = 1, M_POLY = 2 };

static int less_f(double a, double b)
{
    if (!isnan(a)) return !isnan(b) ? a < b : 1;
    return 0;
}

static void stable_order(const double *f, size_t n, size_t *idx)
{
    size_t *tmp = (size_t *)malloc(n * sizeof(size_t));
    for (size_t i = 0; i < n; ++i) idx[i] = i;
    for (size_t w = 1; w < n; w *= 2) {
        for (size_t lo = 0; lo < n; lo += 2 * w) {
            size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n, a = lo, b = mid, k = lo;
            while (a < mid && b < hi) tmp[k++] = less_f(f[idx[b]], f[idx[a]]) ? idx[b++] : idx[a++];
            while (a < mid) tmp[k++] = idx[a++];
            while (b < hi) tmp[k++] = idx[b++];
        }
        memcpy(idx, tmp, n * sizeof(size_t));
    }
    free(tmp);
}

static double betaq_of(double beta, double eta_c, double rand01)
{
    const double alpha = 2. - pow(beta, -(eta_c + 1.));
    if (rand01 < (1. / alpha)) return pow(rand01 * alpha, 1. / (eta_c + 1.));
    return pow(1. / (2. - rand01 * alpha), 1. / (eta_c + 1.));
}

static void mutate(double *c, size_t nx, const double *lb, const double *ub, double m, double param_m, unsigned mutation, oracle_stream *rs)
{
    for (size_t g = 0; g < nx; ++g) {
        if (oracle_next(rs) < m) {
            if (mutation == M_UNIFORM) {
                c[g] = (lb[g] == ub[g]) ? lb[g] : (ub[g] - lb[g]) * oracle_next(rs) + lb[g];
            } else if (mutation == M_GAUSS) {
                const double sd = (ub[g] - lb[g]) * param_m;
                const double u1 = 1.0 - oracle_next(rs), u2 = oracle_next(rs);
                c[g] += (sqrt(-2.0 * log(u1)) * cos(2.0 * 3.141592653589793238462643383279502884 * u2)) * sd;
            } else {
                const double u = oracle_next(rs);
                if (u <= 0.5) {
                    const double delta_l = pow(2. * u, 1. / (1. + param_m)) - 1.;
                    c[g] += delta_l * (c[g] - lb[g]);
                } else {
                    const double delta_r = 1 - pow(2. * (1. - u), 1. / (1. + param_m));
                    c[g] += delta_r * (ub[g] - c[g]);
                }
            }
        }
        if (c[g] < lb[g]) c[g] = lb[g];
        if (c[g] > ub[g]) c[g] = ub[g];
    }
}

int oracle_sga_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t nx, unsigned gens,
                      double cr, double eta_c, double m, double param_m, unsigned param_s, unsigned crossover, unsigned mutation,
                      unsigned selection, uint64_t seed, uint32_t first_generation)
{
    if (NP < 2 || param_s < 1 || param_s > NP || crossover > 3 || mutation > 2 || selection > 1) return -1;
    if (crossover == X_SBX && NP % 2) return -1;
    size_t *sel = (size_t *)malloc(NP * sizeof(size_t)), *order = (size_t *)malloc(2 * NP * sizeof(size_t));
    size_t *perm = (size_t *)malloc(NP * sizeof(size_t)), *virt = (size_t *)malloc(NP * sizeof(size_t));
    double *xnew = (double *)malloc(NP * nx * sizeof(double)), *fboth = (double *)malloc(2 * NP * sizeof(double));
    double *xo = (double *)malloc(NP * nx * sizeof(double)), *fo = (double *)malloc(NP * sizeof(double));
    int rc = 0;
    for (unsigned g = 0; g < gens && !rc; ++g) {
        const uint32_t generation = first_generation + g;
        /* selection */
        if (selection == 1) {
            stable_order(f, NP, order);
            for (size_t j = 0; j < NP; ++j) sel[j] = order[j % param_s];
        } else {
            for (size_t j = 0; j < NP; ++j) {
                oracle_stream rs = {seed, ORACLE_TAG_SGA, generation, (uint32_t)j, 0};
                for (size_t k = 0; k < NP; ++k) virt[k] = k;
                for (unsigned i = 0; i < param_s; ++i) {
                    size_t index = i + (size_t)(oracle_next(&rs) * (double)(NP - i));
                    if (index >= NP) index = NP - 1;
                    const size_t t = virt[index];
                    virt[index] = virt[i];
                    virt[i] = t;
                }
                size_t winner = virt[0];
                for (unsigned i = 1; i < param_s; ++i)
                    if (f[virt[i]] < f[winner]) winner = virt[i];
                sel[j] = winner;
            }
        }
        /* crossover + mutation */
        if (crossover == X_SBX) {
            if (oracle_philox_perm(NP, seed, ORACLE_TAG_SHUFFLE1, generation, perm)) { rc = -1; break; }
            for (size_t t = 0; t < NP / 2; ++t) {
                oracle_stream rs = {seed, ORACLE_TAG_SGA, generation, (uint32_t)(NP + t), 0};
                const double *p1 = x + sel[perm[2 * t]] * nx, *p2 = x + sel[perm[2 * t + 1]] * nx;
                double *c1 = xnew + 2 * t * nx, *c2 = c1 + nx;
                memcpy(c1, p1, nx * sizeof(double));
                memcpy(c2, p2, nx * sizeof(double));
                if (oracle_next(&rs) < cr) {
                    for (size_t i = 0; i < nx; ++i) {
                        const double a = p1[i], b = p2[i], yl = lb[i], yu = ub[i];
                        if ((oracle_next(&rs) < 0.5) && (fabs(a - b)) > 1e-14 && yl != yu) {
                            const double y1 = (a < b) ? a : b, y2 = (a < b) ? b : a;
                            const double rand01 = oracle_next(&rs);
                            double beta = 1. + (2. * (y1 - yl) / (y2 - y1));
                            double betaq = betaq_of(beta, eta_c, rand01);
                            double v1 = 0.5 * ((y1 + y2) - betaq * (y2 - y1));
                            beta = 1. + (2. * (yu - y2) / (y2 - y1));
                            betaq = betaq_of(beta, eta_c, rand01);
                            double v2 = 0.5 * ((y1 + y2) + betaq * (y2 - y1));
                            if (v1 < yl) v1 = yl;
                            if (v2 < yl) v2 = yl;
                            if (v1 > yu) v1 = yu;
                            if (v2 > yu) v2 = yu;
                            if (oracle_next(&rs) < .5) { c1[i] = v1; c2[i] = v2; }
                            else { c1[i] = v2; c2[i] = v1; }
                        }
                    }
                }
                mutate(c1, nx, lb, ub, m, param_m, mutation, &rs);
                mutate(c2, nx, lb, ub, m, param_m, mutation, &rs);
            }
        } else {
            for (size_t t = 0; t < NP; ++t) {
                oracle_stream rs = {seed, ORACLE_TAG_SGA, generation, (uint32_t)(NP + t), 0};
                double *child = xnew + t * nx;
                memcpy(child, x + sel[t] * nx, nx * sizeof(double));
                size_t pidx = 1 + (size_t)(oracle_next(&rs) * (double)(NP - 1));
                if (pidx > NP - 1) pidx = NP - 1;
                const size_t partner = (pidx <= t) ? pidx - 1 : pidx; /* all_idx after swap(all_idx[0], all_idx[t]), :411-412 */
                const double *parent2 = x + sel[partner] * nx;
#define GENE(n)                                                 \
    do {                                                        \
        n = (size_t)(oracle_next(&rs) * (double)nx);            \
        if (n >= nx) n = nx - 1;                                \
    } while (0)
                if (crossover == X_EXP) {
                    size_t n, L = 0;
                    GENE(n);
                    do {
                        child[n] = parent2[n];
                        n = (n + 1u) % nx;
                        ++L;
                    } while ((oracle_next(&rs) < cr) && (L < nx));
                } else if (crossover == X_BIN) {
                    size_t n;
                    GENE(n);
                    for (size_t L = 0; L < nx; ++L) {
                        if ((oracle_next(&rs) < cr) || L + 1 == nx) child[n] = parent2[n];
                        n = (n + 1) % nx;
                    }
                } else if (oracle_next(&rs) < cr) {
                    size_t n;
                    GENE(n);
                    for (size_t k = n; k < nx; ++k) child[k] = parent2[k];
                }
                mutate(child, nx, lb, ub, m, param_m, mutation, &rs);
            }
        }
        /* evaluation and reinsertion */
        if ((rc = oracle_problem_eval(prob, xnew, NP, fboth))) break;
        memcpy(fboth + NP, f, NP * sizeof(double));
        stable_order(fboth, 2 * NP, order);
        for (size_t j = 0; j < NP; ++j) {
            const size_t s = order[j];
            memcpy(xo + j * nx, (s < NP ? xnew + s * nx : x + (s - NP) * nx), nx * sizeof(double));
            fo[j] = fboth[s];
        }
        memcpy(x, xo, NP * nx * sizeof(double));
        memcpy(f, fo, NP * sizeof(double));
    }
    free(sel); free(order); free(perm); free(virt); free(xnew); free(fboth); free(xo); free(fo);
    return rc;
}
