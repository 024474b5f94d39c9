// oracle/ref_pin.cpp - C entry points that run the UNMODIFIED reference operators and algorithms from caller-given
// starting populations and seeds, so that the plain-C restatements (oracle/restate_*.c in sequential-mt19937 mode) can be
// compared with them bit for bit.  TEST INFRASTRUCTURE ONLY; see ref_capi.h.
#include <algorithm>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include <pagmo/algorithm.hpp>
#include <pagmo/algorithms/de.hpp>
#include <pagmo/algorithms/de1220.hpp>
#include <pagmo/algorithms/gaco.hpp>
#include <pagmo/algorithms/maco.hpp>
#include <pagmo/algorithms/moead_gen.hpp>
#include <pagmo/algorithms/nsga2.hpp>
#include <pagmo/algorithms/nspso.hpp>
#include <pagmo/algorithms/pso_gen.hpp>
#include <pagmo/algorithms/sade.hpp>
#include <pagmo/algorithms/sga.hpp>
#include <pagmo/population.hpp>
#include <pagmo/problem.hpp>
#include <pagmo/rng.hpp>
#include <pagmo/utils/generic.hpp>
#include <pagmo/utils/genetic_operators.hpp>
#include <pagmo/utils/multi_objective.hpp>

#include "ref_capi.h"

struct ref_problem {
    pagmo::problem prob;
};

extern "C" void ref_set_error(const char *);

template <typename F> static int guarded(F &&f)
{
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        ref_set_error(e.what());
        return 1;
    }
}

extern "C" {

// The real std:: classes on a std::mt19937 (= pagmo::detail::random_engine_type) seeded with `seed`:
// kind 0 raw words, 1 uniform_real_distribution<double>(0,1), 2 uniform_int_distribution<size_t>(a,b),
// 3 ONE normal_distribution<double>(0,1) object, 4 uniform_real_distribution<double>(-a,b).
int ref_std_sequence(unsigned seed, int kind, unsigned long long a, unsigned long long b, size_t n, double *out_real,
                     unsigned long long *out_int)
{
    return guarded([&] {
        pagmo::detail::random_engine_type e(seed);
        std::uniform_real_distribution<double> u01(0., 1.);
        std::normal_distribution<double> nd(0., 1.);
        for (size_t i = 0; i < n; ++i) switch (kind) {
                case 0: out_int[i] = e(); break;
                case 1: out_real[i] = u01(e); break;
                case 2: out_int[i] = std::uniform_int_distribution<size_t>(a, b)(e); break;
                case 3: out_real[i] = nd(e); break;
                case 4: out_real[i] = std::uniform_real_distribution<double>(-static_cast<double>(a), static_cast<double>(b))(e); break;
                default: throw std::invalid_argument("ref_std_sequence: kind");
            }
    });
}

// `rounds` successive std::shuffle calls on one iota vector (what nsga2.cpp:138-139,180-181 does with shuffle1)
int ref_std_shuffles(unsigned seed, size_t n, size_t rounds, size_t *perm)
{
    return guarded([&] {
        pagmo::detail::random_engine_type e(seed);
        std::vector<size_t> v(n);
        for (size_t i = 0; i < n; ++i) v[i] = i;
        for (size_t r = 0; r < rounds; ++r) std::shuffle(v.begin(), v.end(), e);
        std::copy(v.begin(), v.end(), perm);
    });
}

// std::binomial_distribution<size_t>(t, p) - one object, n draws
int ref_std_binomial(unsigned seed, unsigned long long t, double p, size_t n, unsigned long long *out)
{
    return guarded([&] {
        pagmo::detail::random_engine_type e(seed);
        std::binomial_distribution<size_t> d(t, p);
        for (size_t i = 0; i < n; ++i) out[i] = d(e);
    });
}

// detail::sbx_crossover_impl (genetic_operators.cpp:71-144), then detail::polynomial_mutation_impl (:148-197) on both
// children, then n_pairs calls of detail::mo_tournament_selection_impl (:200-211) on (2i, 2i+1) - all on one engine.
int ref_genetic_operators(const double *p1, const double *p2, size_t nx, const double *lb, const double *ub, double p_cr, double eta_c,
                          double p_m, double eta_m, const size_t *rank, const double *cd, size_t n_pairs, unsigned seed, double *c1,
                          double *c2, size_t *winners)
{
    return guarded([&] {
        pagmo::detail::random_engine_type e(seed);
        const pagmo::vector_double a(p1, p1 + nx), b(p2, p2 + nx);
        const std::pair<pagmo::vector_double, pagmo::vector_double> bounds{pagmo::vector_double(lb, lb + nx), pagmo::vector_double(ub, ub + nx)};
        auto ch = pagmo::detail::sbx_crossover_impl(a, b, bounds, 0u, p_cr, eta_c, e);
        pagmo::detail::polynomial_mutation_impl(ch.first, bounds, 0u, p_m, eta_m, e);
        pagmo::detail::polynomial_mutation_impl(ch.second, bounds, 0u, p_m, eta_m, e);
        std::copy(ch.first.begin(), ch.first.end(), c1);
        std::copy(ch.second.begin(), ch.second.end(), c2);
        const std::vector<pagmo::vector_double::size_type> r(rank, rank + 2 * n_pairs);
        const std::vector<double> c(cd, cd + 2 * n_pairs);
        for (size_t i = 0; i < n_pairs; ++i) winners[i] = pagmo::detail::mo_tournament_selection_impl(2 * i, 2 * i + 1, r, c, e);
    });
}

// Runs an unmodified reference UDA on the population whose decision vectors are x0 [n x nx] (fitness computed by the problem,
// population.cpp:570-596).  par[] per algorithm, in constructor order without gen/seed:
//   "nsga2"   cr, eta_c, m, eta_m                                   (nsga2.hpp)
//   "pso_gen" omega, eta1, eta2, max_vel, variant, neighb_type, neighb_param   (pso_gen.hpp:109)
//   "de"      F, CR, variant, ftol, xtol                            (de.hpp:103)
//   "sade"    variant, variant_adptv, ftol, xtol                    (sade.hpp:138)
//   "de1220"  variant_adptv, ftol, xtol, n_allowed, allowed...      (de1220.hpp:155)
//   "sga"     cr, eta_c, m, param_m, param_s; strategies "crossover,mutation,selection" in `strategies`  (sga.hpp:166)
int ref_evolve_from(ref_problem *p, const char *algo, const double *par, size_t npar, const char *strategies, const double *x0, size_t n,
                    unsigned gens, unsigned seed, double *x_out, double *f_out)
{
    return guarded([&] {
        const std::string a(algo);
        const auto nx = p->prob.get_nx(), nf = p->prob.get_nf();
        pagmo::population pop(p->prob, 0u, 0u);
        for (size_t i = 0; i < n; ++i) pop.push_back(pagmo::vector_double(x0 + i * nx, x0 + (i + 1) * nx));
        auto need = [&](size_t k) {
            if (npar < k) throw std::invalid_argument("ref_evolve_from: too few parameters for '" + a + "'");
        };
        pagmo::algorithm alg;
        unsigned calls = 1; // evolve() calls of the one algorithm object (gaco with memory)
        if (a == "nsga2") {
            need(4);
            alg = pagmo::algorithm{pagmo::nsga2(gens, par[0], par[1], par[2], par[3], seed)};
        } else if (a == "pso_gen") {
            need(7);
            alg = pagmo::algorithm{pagmo::pso_gen(gens, par[0], par[1], par[2], par[3], static_cast<unsigned>(par[4]),
                                                  static_cast<unsigned>(par[5]), static_cast<unsigned>(par[6]), false, seed)};
        } else if (a == "de") {
            need(5);
            alg = pagmo::algorithm{pagmo::de(gens, par[0], par[1], static_cast<unsigned>(par[2]), par[3], par[4], seed)};
        } else if (a == "sade") {
            need(4);
            alg = pagmo::algorithm{pagmo::sade(gens, static_cast<unsigned>(par[0]), static_cast<unsigned>(par[1]), par[2], par[3], false, seed)};
        } else if (a == "de1220") {
            need(4);
            const auto na = static_cast<size_t>(par[3]);
            need(4 + na);
            std::vector<unsigned> allowed;
            for (size_t i = 0; i < na; ++i) allowed.push_back(static_cast<unsigned>(par[4 + i]));
            alg = pagmo::algorithm{pagmo::de1220(gens, allowed, static_cast<unsigned>(par[0]), par[1], par[2], false, seed)};
        } else if (a == "sga") {
            need(5);
            const std::string s(strategies ? strategies : "exponential,polynomial,tournament");
            const auto c1 = s.find(','), c2 = s.find(',', c1 + 1);
            if (c1 == std::string::npos || c2 == std::string::npos) throw std::invalid_argument("ref_evolve_from: strategies");
            alg = pagmo::algorithm{pagmo::sga(gens, par[0], par[1], par[2], par[3], static_cast<unsigned>(par[4]), s.substr(0, c1),
                                              s.substr(c1 + 1, c2 - c1 - 1), s.substr(c2 + 1), seed)};
        } else if (a == "moead_gen") { // strategies = "<weight generation>,<decomposition>"
            need(7);
            const std::string s(strategies ? strategies : "grid,tchebycheff");
            const auto c1 = s.find(',');
            if (c1 == std::string::npos) throw std::invalid_argument("ref_evolve_from: strategies");
            alg = pagmo::algorithm{pagmo::moead_gen(gens, s.substr(0, c1), s.substr(c1 + 1), static_cast<unsigned>(par[0]), par[1], par[2], par[3],
                                                    par[4], static_cast<unsigned>(par[5]), par[6] != 0., seed)};
        } else if (a == "nspso") {
            need(6);
            alg = pagmo::algorithm{pagmo::nspso(gens, par[0], par[1], par[2], par[3], par[4], static_cast<unsigned>(par[5]),
                                                std::string(strategies ? strategies : "crowding distance"), false, seed)};
        } else if (a == "gaco") { // ker, q, oracle, acc, threshold, n_gen_mark, impstop, evalstop, focus (gaco.hpp:104-107) [, memory, calls]
            if (npar != 9 && npar != 11) throw std::invalid_argument("ref_evolve_from: gaco takes 9 or 11 parameters");
            const bool memory = npar == 11 && par[9] != 0.;
            if (npar == 11) calls = static_cast<unsigned>(par[10]);
            alg = pagmo::algorithm{pagmo::gaco(gens, static_cast<unsigned>(par[0]), par[1], par[2], par[3], static_cast<unsigned>(par[4]),
                                               static_cast<unsigned>(par[5]), static_cast<unsigned>(par[6]), static_cast<unsigned>(par[7]), par[8],
                                               memory, seed)};
        } else if (a == "maco") { // ker, q, threshold, n_gen_mark, evalstop, focus (maco.hpp:107-109)
            need(6);
            alg = pagmo::algorithm{pagmo::maco(gens, static_cast<unsigned>(par[0]), par[1], static_cast<unsigned>(par[2]),
                                               static_cast<unsigned>(par[3]), static_cast<unsigned>(par[4]), par[5], false, seed)};
        } else
            throw std::invalid_argument("ref_evolve_from: unknown algorithm '" + a + "'");
        for (unsigned c = 0; c < calls; ++c) pop = alg.evolve(pop); // the same algorithm object every time: its members carry over
        for (size_t i = 0; i < n; ++i) {
            if (x_out) std::memcpy(x_out + i * nx, pop.get_x()[i].data(), nx * sizeof(double));
            if (f_out) std::memcpy(f_out + i * nf, pop.get_f()[i].data(), nf * sizeof(double));
        }
    });
}

// decomposition_weights (multi_objective.hpp:128-213) with a fresh std::mt19937(seed), and kNN of the rows (generic.cpp:107-148)
int ref_decomposition_weights(size_t n_f, size_t n_w, const char *method, unsigned seed, double *out)
{
    return guarded([&] {
        std::mt19937 e(seed);
        const auto w = pagmo::decomposition_weights(n_f, n_w, std::string(method), e);
        for (size_t i = 0; i < w.size(); ++i) std::memcpy(out + i * n_f, w[i].data(), n_f * sizeof(double));
    });
}
int ref_knn(const double *points, size_t n, size_t m, size_t k, size_t *out)
{
    return guarded([&] {
        std::vector<pagmo::vector_double> p(n);
        for (size_t i = 0; i < n; ++i) p[i].assign(points + i * m, points + (i + 1) * m);
        const auto r = pagmo::kNN(p, k);
        for (size_t i = 0; i < n; ++i)
            for (size_t j = 0; j < k; ++j) out[i * k + j] = r[i][j];
    });
}

// population::population(prob, n, seed) (population.cpp:62-80 -> random_decision_vector, utils/generic.cpp) : the decision
// vectors a seeded reference population starts from
int ref_population_init(ref_problem *p, size_t n, unsigned seed, double *x_out, unsigned long long *ids_out)
{
    return guarded([&] {
        pagmo::population pop(p->prob, n, seed);
        const auto nx = p->prob.get_nx();
        for (size_t i = 0; i < n; ++i) {
            std::memcpy(x_out + i * nx, pop.get_x()[i].data(), nx * sizeof(double));
            if (ids_out) ids_out[i] = pop.get_ID()[i];
        }
    });
}

} // extern "C"

// std::sort of iota(n) by keys (ascending with operator<, or descending with operator>): the tie order of the library
extern "C" int ref_std_argsort(const double *keys, size_t n, int desc, size_t *out)
{
    return guarded([&] {
        std::vector<size_t> v(n);
        for (size_t i = 0; i < n; ++i) v[i] = i;
        if (desc) std::sort(v.begin(), v.end(), [keys](size_t a, size_t b) { return keys[a] > keys[b]; });
        else std::sort(v.begin(), v.end(), [keys](size_t a, size_t b) { return keys[a] < keys[b]; });
        std::copy(v.begin(), v.end(), out);
    });
}
