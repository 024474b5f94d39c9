// oracle/ref_algos.cpp - C entry points over the UNMODIFIED reference algorithms (nsga2, de, de1220, sade, pso, pso_gen).
// TEST INFRASTRUCTURE ONLY (CPU baseline timing and behavioural comparison); see ref_capi.h.
#include <chrono>
#include <cstring>
#include <stdexcept>
#include <string>

#include <pagmo/algorithm.hpp>
#include <pagmo/algorithms/de.hpp>
#include <pagmo/algorithms/de1220.hpp>
#include <pagmo/algorithms/nsga2.hpp>
#include <pagmo/algorithms/pso.hpp>
#include <pagmo/algorithms/pso_gen.hpp>
#include <pagmo/algorithms/sade.hpp>
#include <pagmo/batch_evaluators/thread_bfe.hpp>
#include <pagmo/bfe.hpp>
#include <pagmo/population.hpp>
#include <pagmo/problem.hpp>

#include "ref_capi.h"

struct ref_problem {
    pagmo::problem prob;
};

extern "C" {

// Runs `algo` for `gens` generations on a fresh population(prob, pop_size, pop_seed) and returns the wall time of
// evolve() alone (seconds) plus the final population (x [pop_size x nx], f [pop_size x nf], either may be NULL).
// algo: "nsga2" (bfe = thread_bfe when use_bfe), "de", "de1220", "sade", "pso", "pso_gen" - reference default parameters.
int ref_evolve(ref_problem *p, const char *algo, unsigned pop_size, unsigned gens, unsigned pop_seed, unsigned algo_seed,
               int use_bfe, double *seconds, double *x_out, double *f_out, unsigned long long *fevals)
{
    try {
        const std::string a(algo);
        pagmo::population pop(p->prob, pop_size, pop_seed);
        const auto f0 = pop.get_problem().get_fevals();
        pagmo::algorithm alg;
        if (a == "nsga2") {
            pagmo::nsga2 u(gens, 0.95, 10., 0.01, 50., algo_seed);
            if (use_bfe) u.set_bfe(pagmo::bfe{pagmo::thread_bfe{}});
            alg = pagmo::algorithm{u};
        } else if (a == "de") alg = pagmo::algorithm{pagmo::de(gens, 0.8, 0.9, 2u, 1e-6, 1e-6, algo_seed)};
        else if (a == "de1220") alg = pagmo::algorithm{pagmo::de1220(gens, pagmo::de1220_statics<void>::allowed_variants, 1u, 1e-6, 1e-6, false, algo_seed)};
        else if (a == "sade") alg = pagmo::algorithm{pagmo::sade(gens, 2u, 1u, 1e-6, 1e-6, false, algo_seed)};
        else if (a == "pso") alg = pagmo::algorithm{pagmo::pso(gens, 0.7298, 2.05, 2.05, 0.5, 5u, 2u, 4u, false, algo_seed)};
        else if (a == "pso_gen") alg = pagmo::algorithm{pagmo::pso_gen(gens, 0.7298, 2.05, 2.05, 0.5, 5u, 2u, 4u, false, algo_seed)};
        else throw std::invalid_argument("ref_evolve: unknown algorithm '" + a + "'");
        const auto t0 = std::chrono::steady_clock::now();
        pop = alg.evolve(pop);
        const auto t1 = std::chrono::steady_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
        if (fevals) *fevals = pop.get_problem().get_fevals() - f0;
        const auto nx = p->prob.get_nx(), nf = p->prob.get_nf();
        for (unsigned i = 0; i < pop_size; ++i) {
            if (x_out) std::memcpy(x_out + i * nx, pop.get_x()[i].data(), nx * sizeof(double));
            if (f_out) std::memcpy(f_out + i * nf, pop.get_f()[i].data(), nf * sizeof(double));
        }
        return 0;
    } catch (const std::exception &e) {
        extern void ref_set_error(const char *);
        ref_set_error(e.what());
        return 1;
    }
}

} // extern "C"
