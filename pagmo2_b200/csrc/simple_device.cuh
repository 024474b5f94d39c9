// simple_device.cuh - per-coordinate arithmetic of pagmo's simple single-objective UDPs, shared by the batch evaluator
// (eval_simple.cu) and the resident differential-evolution loop (de.cu) so that both produce the same bits:
//   rastrigin (src/problems/rastrigin.cpp:62-72), ackley (ackley.cpp:61-76), griewank (griewank.cpp:60-75),
//   schwefel (schwefel.cpp:60-69), rosenbrock (rosenbrock.cpp:59-66).
// A coordinate contributes a TERM (independent of the other coordinates, so terms can be computed in parallel) that is then
// FOLDED into the running sums in the reference's order, j ascending.  Compiled with -fmad=false: no contraction.
#pragma once

#include "../../include/pagmo_cuda/pgc.h"

namespace pgc
{
namespace simple
{

struct Acc {
    double a, b;
};

template <int FAM> __device__ __forceinline__ void acc_init(Acc &s)
{
    s.a = 0.0;
    s.b = (FAM == PGC_GRIEWANK) ? 1.0 : 0.0;
}

// term of coordinate j (value x); `has_next` / `xn` give x[j+1] for rosenbrock
template <int FAM> __device__ __forceinline__ void acc_term(double x, int j, bool has_next, double xn, double &ta, double &tb)
{
    const double omega = 2.0 * 3.141592653589793238462643383279502884;
    ta = 0.0;
    tb = 0.0;
    if (FAM == PGC_RASTRIGIN) {
        ta = x * x - 10.0 * cos(omega * x);
    } else if (FAM == PGC_ACKLEY) {
        ta = x * x;
        tb = cos(omega * x);
    } else if (FAM == PGC_GRIEWANK) {
        ta = x * x;
        tb = cos(x / sqrt(static_cast<double>(j) + 1.0));
    } else if (FAM == PGC_SCHWEFEL) {
        ta = x * sin(sqrt(fabs(x)));
    } else if (FAM == PGC_ROSENBROCK) {
        if (has_next) ta = 100.0 * (x * x - xn) * (x * x - xn) + (x - 1.0) * (x - 1.0);
    }
}

template <int FAM> __device__ __forceinline__ void acc_fold(Acc &s, double ta, double tb, bool has_next)
{
    if (FAM == PGC_GRIEWANK) {
        s.a += ta;
        s.b *= tb;
    } else if (FAM == PGC_ACKLEY) {
        s.a += ta;
        s.b += tb;
    } else if (FAM == PGC_ROSENBROCK) {
        if (has_next) s.a += ta;
    } else {
        s.a += ta;
    }
}

template <int FAM> __device__ __forceinline__ void acc_step(Acc &s, double x, int j, bool has_next, double xn)
{
    double ta, tb;
    acc_term<FAM>(x, j, has_next, xn, ta, tb);
    acc_fold<FAM>(s, ta, tb, has_next);
}

template <int FAM> __device__ __forceinline__ double acc_final(const Acc &s, int D)
{
    const double n = static_cast<double>(D);
    if (FAM == PGC_RASTRIGIN) return s.a + 10.0 * n;
    if (FAM == PGC_ACKLEY)
        return -20.0 * exp(-0.2 * sqrt(1.0 / n * s.a)) - exp(1.0 / n * s.b) + 20.0 + 2.718281828459045235360287471352662498; // nepero = std::exp(1.0)
    if (FAM == PGC_GRIEWANK) return (s.a / 4000.0 - s.b + 1.0);
    if (FAM == PGC_SCHWEFEL) return 418.9828872724338 * n - s.a;
    return s.a;
}

} // namespace simple
} // namespace pgc
