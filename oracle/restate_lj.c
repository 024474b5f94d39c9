/* oracle/restate_lj.c - plain-C restatement of pagmo::lennard_jones::fitness.  TEST INFRASTRUCTURE ONLY.
 * Follows reference src/problems/lennard_jones.cpp:72-92 (fitness), :132-151 (_r), :99-110 (bounds) operation by
 * operation; pinned bit-exactly against oracle/_ref and against the reference's known answers
 * (tests/lennard_jones.cpp:59-60) in tests/test_oracle.py. */
#include <float.h>
#include <math.h>
#include <stddef.h>

#include "oracle.h"

static double r_(unsigned atom, unsigned coord, const double *x)
{
    if (atom == 0u) return 0.0;
    if (atom == 1u) return coord < 2u ? 0.0 : x[0];
    if (atom == 2u) return coord == 0u ? 0.0 : x[coord];
    return x[3u * (atom - 2u) + coord];
}

int oracle_lj_fitness(unsigned atoms, const double *x, double *f)
{
    if (atoms < 3) return -1;
    double acc = 0.;
    for (unsigned i = 0u; i < (atoms - 1u); ++i)
        for (unsigned j = (i + 1u); j < atoms; ++j) {
            double sixth, dist;
            dist = pow(r_(i, 0u, x) - r_(j, 0u, x), 2) + pow(r_(i, 1u, x) - r_(j, 1u, x), 2) + pow(r_(i, 2u, x) - r_(j, 2u, x), 2);
            if (dist == 0.0) {
                acc = DBL_MAX;
            } else {
                sixth = pow(dist, -3);
                acc += (pow(sixth, 2) - sixth);
            }
        }
    *f = 4 * acc;
    return 0;
}

int oracle_lj_batch(unsigned atoms, const double *xs, size_t n, double *fs)
{
    const size_t D = 3u * atoms - 6u;
    for (size_t i = 0; i < n; ++i)
        if (oracle_lj_fitness(atoms, xs + i * D, fs + i)) return -1;
    return 0;
}
