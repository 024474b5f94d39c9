// cuda_algorithms.hpp - header-only user-defined algorithms (UDAs, reference include/pagmo/algorithm.hpp:122-160) whose evolve()
// runs whole generations on the device: trial construction / variation, batch fitness, selection.
//
//   pagmo_cuda::cuda_de      pagmo::de      (de.hpp:119      gen, F, CR, variant, ftol, xtol, seed)
//   pagmo_cuda::cuda_sade    pagmo::sade    (sade.hpp:138    gen, variant, variant_adptv, ftol, xtol, memory, seed)
//   pagmo_cuda::cuda_de1220  pagmo::de1220  (de1220.hpp:158  gen, allowed_variants, variant_adptv, ftol, xtol, memory, seed)
//   pagmo_cuda::cuda_pso_gen pagmo::pso_gen (pso_gen.hpp:127 gen, omega, eta1, eta2, max_vel, variant, neighb_type, neighb_param, memory, seed)
//   pagmo_cuda::cuda_pso     pagmo::pso     (pso.hpp:110     same arguments; batched per generation like pso_gen)
//   pagmo_cuda::cuda_nsga2   pagmo::nsga2   (nsga2.hpp:103   gen, cr, eta_c, m, eta_m, seed)
//   pagmo_cuda::cuda_sga     pagmo::sga     (sga.hpp:166     gen, cr, eta_c, m, param_m, param_s, crossover, mutation, selection, seed)
//   pagmo_cuda::cuda_cmaes   pagmo::cmaes   (cmaes.hpp:110   gen, cc, cs, c1, cmu, sigma0, ftol, xtol, memory, force_bounds, seed)
//   pagmo_cuda::cuda_xnes    pagmo::xnes    (xnes.hpp:107    gen, eta_mu, eta_sigma, eta_b, sigma0, ftol, xtol, memory, force_bounds, seed)
//   pagmo_cuda::cuda_gaco    pagmo::gaco    (gaco.hpp:104    gen, ker, q, oracle, acc, threshold, n_gen_mark, impstop, evalstop, focus, memory, seed)
//   pagmo_cuda::cuda_maco    pagmo::maco    (maco.hpp:107    gen, ker, q, threshold, n_gen_mark, evalstop, focus, memory, seed)
//
// Same constructor arguments as the reference UDAs (plus the device), so `algorithm{cuda_sade{50u}}` drops into an island of a
// stock pagmo::archipelago: thread_island (thread_island.cpp:79-159) runs it unchanged, and pagmo's own migration machinery
// (island.cpp:461-641) keeps working on the host populations.  The population's problem must have a device evaluator: a
// pagmo_cuda:: UDP or a stock UDP that cuda_bfe recognises; otherwise evolve() throws (no CPU fallback).
// Differences from the reference algorithms, as documented in DESIGN.md: the DE family and PSO are GENERATIONAL (all trial
// vectors of a generation are built from the previous population and evaluated in one batch; the reference updates in place,
// individual by individual), random draws come from Philox streams keyed by (seed, generation, individual).  `memory = true`
// (sade, de1220, pso_gen, pso, nspso, cmaes, xnes) keeps the adaptation state / velocities / archive / distribution between evolve()
// calls as the reference does.  fevals are accounted for as the reference does (one per individual per generation).
#ifndef PAGMO_CUDA_CUDA_ALGORITHMS_HPP
#define PAGMO_CUDA_CUDA_ALGORITHMS_HPP

#include <iomanip>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <tuple>
#include <string>
#include <vector>

#include <pagmo/algorithm.hpp>
#include <random>

#include <pagmo/exceptions.hpp>
#include <pagmo/population.hpp>
#include <pagmo/problem.hpp>
#include <pagmo/rng.hpp>
#include <pagmo/s11n.hpp>
#include <pagmo/threading.hpp>
#include <pagmo/types.hpp>
#include <pagmo/utils/generic.hpp>
#include <pagmo/utils/multi_objective.hpp>

#include <pagmo_cuda/cuda_bfe.hpp>
#include <pagmo_cuda/pgc.h>

namespace pagmo_cuda
{

namespace detail
{
// runs `call` with the log capture of pgc.h around it (pgc_log_capture_begin / _end) when verbosity > 0; rows: row_len doubles each
template <typename Call>
inline int with_log_capture(pgc_ctx *ctx, unsigned verbosity, unsigned gens, std::size_t row_len, pagmo::vector_double &rows, Call call)
{
    if (!verbosity) return call();
    const std::size_t max_rows = (gens ? (gens - 1u) / verbosity + 1u : 0u) + 1u;
    if (int rc = pgc_log_capture_begin(ctx, verbosity, max_rows, row_len)) return rc;
    const int rc = call();
    rows.assign(max_rows * row_len, 0.);
    std::size_t n_rows = 0;
    const int rc2 = pgc_log_capture_end(ctx, rows.data(), &n_rows);
    rows.resize(n_rows * row_len);
    return rc != PGC_OK ? rc : rc2;
}
} // namespace detail

// Common part: population <-> device round trip around pgc_algo_evolve_device.
class cuda_algorithm_base
{
public:
    pagmo::population evolve(pagmo::population pop) const
    {
        const auto &prob = pop.get_problem();
        const auto n = pop.size();
        if (m_desc.gens == 0u || n == 0u) return pop; // the reference UDAs return early on gen == 0 (de.cpp:118-120)
        if (prob.get_nc() != 0u) {
            pagmo_throw(std::invalid_argument, "Non linear constraints detected in " + prob.get_name() + " instance. " + get_name()
                                                   + " cannot deal with them");
        }
        if (prob.is_stochastic()) {
            pagmo_throw(std::invalid_argument, "The problem appears to be stochastic " + get_name() + " cannot deal with it");
        }
        const auto h = m_cache->find(prob, m_device);
        if (!h) {
            pagmo_throw(std::invalid_argument, get_name() + " cannot evolve a population of '" + prob.get_name()
                                                   + "': no CUDA evaluator exists for this UDP type; there is no CPU fallback");
        }
        const auto nx = prob.get_nx(), nf = prob.get_nf();
        pagmo::vector_double x(n * nx), f(n * nf);
        for (decltype(pop.size()) i = 0; i < n; ++i) {
            std::copy(pop.get_x()[i].begin(), pop.get_x()[i].end(), x.begin() + static_cast<std::ptrdiff_t>(i * nx));
            std::copy(pop.get_f()[i].begin(), pop.get_f()[i].end(), f.begin() + static_cast<std::ptrdiff_t>(i * nf));
        }
        unsigned done;
        if (m_verbosity) {
            // nspso logs the problem's absolute feval counter (nspso.cpp:191), the others the evaluations of this call
            const double fevals0 = m_desc.algo == PGC_ALGO_NSPSO ? static_cast<double>(prob.get_fevals()) : 0.;
            m_log_rows.clear();
            done = h->evolve_full(m_desc, x, f, m_generation, m_desc.memory ? &m_state : nullptr, m_verbosity, &m_log_rows, &m_log_row_len);
            for (std::size_t r = 0; m_log_row_len && r < m_log_rows.size() / m_log_row_len; ++r) {
                double *row = m_log_rows.data() + r * m_log_row_len;
                row[1] += fevals0;
                if (r % 50u == 0u) std::cout << "\n" << get_name() << ": log columns as in pgc.h (pgc_algo_evolve_logged_device)\n";
                for (std::size_t c = 0; c < m_log_row_len && c < 7u; ++c) std::cout << std::setw(c ? 15 : 7) << row[c];
                std::cout << '\n';
            }
        } else {
            done = m_desc.memory ? h->evolve_memory(m_desc, x, f, m_generation, m_state) : h->evolve(m_desc, x, f, m_generation);
        }
        m_generation += m_desc.gens;
        for (decltype(pop.size()) i = 0; i < n; ++i) {
            pop.set_xf(i, pagmo::vector_double(x.begin() + static_cast<std::ptrdiff_t>(i * nx), x.begin() + static_cast<std::ptrdiff_t>((i + 1) * nx)),
                       pagmo::vector_double(f.begin() + static_cast<std::ptrdiff_t>(i * nf), f.begin() + static_cast<std::ptrdiff_t>((i + 1) * nf)));
        }
        prob.increment_fevals(static_cast<unsigned long long>(done) * n); // what `done` generations of prob.fitness() would have counted
        return pop;
    }
    void set_seed(unsigned seed)
    {
        m_desc.seed = seed;
    }
    // algorithm::set_verbosity (de.hpp:160-178): level > 0 records one log line every `level` generations (and prints it); the
    // typed get_log() of each UDA below returns them in the reference's tuple layout
    void set_verbosity(unsigned level)
    {
        m_verbosity = level;
    }
    unsigned get_verbosity() const
    {
        return m_verbosity;
    }
    // the log of the last evolve() as rows of get_log_row_len() doubles
    const pagmo::vector_double &get_log_rows() const
    {
        return m_log_rows;
    }
    std::size_t get_log_row_len() const
    {
        return m_log_row_len;
    }
    unsigned get_seed() const
    {
        return static_cast<unsigned>(m_desc.seed);
    }
    unsigned get_gen() const
    {
        return m_desc.gens;
    }
    // what a device-resident island (cuda_island.hpp) needs to run this algorithm on ITS copy of the population
    const pgc_algo_desc &descriptor() const
    {
        return m_desc;
    }
    unsigned generation() const
    {
        return m_generation;
    }
    void advance_generation(unsigned gens) const
    {
        m_generation += gens;
    }
    int device() const
    {
        return m_device;
    }
    std::string get_name() const
    {
        return m_name + " [CUDA sm_100a]";
    }
    pagmo::thread_safety get_thread_safety() const
    {
        return pagmo::thread_safety::basic;
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        // every constructor argument lives in m_desc: pagmo default-constructs a UDA before loading it (island / archipelago
        // save + load, pygmo pickling), so anything left out here would silently come back as a default
        std::vector<unsigned> allowed(m_desc.allowed_variants, m_desc.allowed_variants + 18);
        pagmo::detail::archive(ar, m_device, m_name, m_generation, m_desc.algo, m_desc.gens, m_desc.variant, m_desc.variant_adptv,
                               m_desc.neighb_type, m_desc.neighb_param, m_desc.n_allowed, allowed, m_desc.F, m_desc.CR, m_desc.ftol,
                               m_desc.xtol, m_desc.omega, m_desc.eta1, m_desc.eta2, m_desc.max_vel, m_desc.cr, m_desc.eta_c, m_desc.m,
                               m_desc.eta_m, m_desc.seed, m_desc.param_m, m_desc.param_s, m_desc.crossover, m_desc.mutation,
                               m_desc.selection, m_desc.cma_cc, m_desc.cma_cs, m_desc.cma_c1, m_desc.cma_cmu, m_desc.sigma0,
                               m_desc.force_bounds, m_desc.nspso_c1, m_desc.nspso_c2, m_desc.nspso_chi, m_desc.nspso_v_coeff,
                               m_desc.leader_selection_range, m_desc.diversity, m_desc.memory, m_state.a, m_state.b, m_state.c, m_state.u,
                               m_state.initialized, m_verbosity, m_log_rows, m_log_row_len, m_state.es);
        for (std::size_t i = 0; i < 18u && i < allowed.size(); ++i) m_desc.allowed_variants[i] = allowed[i]; // no-op when saving
    }

protected:
    cuda_algorithm_base(int algo, const char *name, unsigned gen, unsigned seed, int device)
        : m_device(device), m_name(name), m_cache(std::make_shared<detail::twin_cache>())
    {
        detail::check(pgc_algo_defaults(algo, gen, seed, &m_desc), "pgc_algo_defaults");
    }
    static void no_memory(bool memory, const char *who)
    {
        if (memory) pagmo_throw(std::invalid_argument, std::string(who) + ": memory = true is not supported on the device path");
    }
    // memory = true (sade.hpp:138, de1220.hpp:158, pso_gen.hpp:127, nspso.hpp:59): F / CR / variant, the velocities, nspso's archive
    // survive between evolve() calls.  The state is a value member (host arrays that travel with the population), so a copy of the
    // algorithm - pagmo copies the UDA in and out of an island around every evolve - carries it like the reference's mutable members.
    void keep_memory(bool memory)
    {
        m_desc.memory = memory ? 1u : 0u;
    }
    pgc_algo_desc m_desc{};
    int m_device = 0;
    std::string m_name;
    mutable unsigned m_generation = 1; // Philox generation counter: successive evolve() calls continue the stream
    mutable detail::problem_handle::algo_state m_state;
    unsigned m_verbosity = 0;
    mutable pagmo::vector_double m_log_rows;
    mutable std::size_t m_log_row_len = 0;
    // rows -> the reference's tuple types
    template <typename Line, typename Make>
    std::vector<Line> typed_log(Make make) const
    {
        std::vector<Line> out;
        for (std::size_t r = 0; m_log_row_len && r < m_log_rows.size() / m_log_row_len; ++r) out.push_back(make(m_log_rows.data() + r * m_log_row_len));
        return out;
    }
    std::shared_ptr<detail::twin_cache> m_cache;
};

class cuda_de : public cuda_algorithm_base
{
public:
    cuda_de(unsigned gen = 1u, double F = 0.8, double CR = 0.9, unsigned variant = 2u, double ftol = 1e-6, double xtol = 1e-6,
            unsigned seed = pagmo::random_device::next(), int device = 0)
        : cuda_algorithm_base(PGC_ALGO_DE, "DE: Differential Evolution", gen, seed, device)
    {
        if (variant < 1u || variant > 10u) { // de.cpp:58-61
            pagmo_throw(std::invalid_argument, "The Differential Evolution variant must be in [1, .., 10], while a value of "
                                                   + std::to_string(variant) + " was detected.");
        }
        if (CR < 0. || F < 0. || CR > 1. || F > 1.) { // de.cpp:62-65
            pagmo_throw(std::invalid_argument, "The F and CR parameters must be in the [0,1] range");
        }
        m_desc.F = F, m_desc.CR = CR, m_desc.variant = variant, m_desc.ftol = ftol, m_desc.xtol = xtol;
    }
    using log_line_type = std::tuple<unsigned, unsigned long long, double, double, double>; // Gen, Fevals, Best, dx, df (de.hpp:104)
    using log_type = std::vector<log_line_type>;
    log_type get_log() const
    {
        return typed_log<log_line_type>([](const double *r) {
            return log_line_type(static_cast<unsigned>(r[0]), static_cast<unsigned long long>(r[1]), r[2], r[3], r[4]);
        });
    }
};

class cuda_sade : public cuda_algorithm_base
{
public:
    cuda_sade(unsigned gen = 1u, unsigned variant = 2u, unsigned variant_adptv = 1u, double ftol = 1e-6, double xtol = 1e-6, bool memory = false,
              unsigned seed = pagmo::random_device::next(), int device = 0)
        : cuda_algorithm_base(PGC_ALGO_SADE, "saDE: Self-adaptive Differential Evolution", gen, seed, device)
    {
        if (variant < 1u || variant > 18u) { // sade.cpp:62-65
            pagmo_throw(std::invalid_argument, "The Differential Evolution mutation variant must be in [1, .., 18], while a value of "
                                                   + std::to_string(variant) + " was detected.");
        }
        if (variant_adptv < 1u || variant_adptv > 2u) { // sade.cpp:66-69
            pagmo_throw(std::invalid_argument, "The variant for self-adaptation must be in [1,2], while a value of "
                                                   + std::to_string(variant_adptv) + " was detected.");
        }
        keep_memory(memory);
        m_desc.variant = variant, m_desc.variant_adptv = variant_adptv, m_desc.ftol = ftol, m_desc.xtol = xtol;
    }
    using log_line_type = std::tuple<unsigned, unsigned long long, double, double, double, double, double>; // Gen, Fevals, Best, F, CR, dx, df
    using log_type = std::vector<log_line_type>;
    log_type get_log() const
    {
        return typed_log<log_line_type>([](const double *r) {
            return log_line_type(static_cast<unsigned>(r[0]), static_cast<unsigned long long>(r[1]), r[2], r[3], r[4], r[5], r[6]);
        });
    }
};

class cuda_de1220 : public cuda_algorithm_base
{
public:
    cuda_de1220(unsigned gen = 1u, std::vector<unsigned> allowed_variants = {2u, 3u, 7u, 10u, 13u, 14u, 15u, 16u}, unsigned variant_adptv = 1u,
                double ftol = 1e-6, double xtol = 1e-6, bool memory = false, unsigned seed = pagmo::random_device::next(), int device = 0)
        : cuda_algorithm_base(PGC_ALGO_DE1220, "sa-DE1220: Self-adaptive Differential Evolution 1220", gen, seed, device)
    {
        for (auto v : allowed_variants) { // de1220.cpp:62-68
            if (v < 1u || v > 18u) {
                pagmo_throw(std::invalid_argument,
                            "All mutation variants considered must be in [1, .., 18], while a value of " + std::to_string(v) + " was detected.");
            }
        }
        if (variant_adptv < 1u || variant_adptv > 2u) { // de1220.cpp:70-73
            pagmo_throw(std::invalid_argument, "The variant for self-adaptation must be in [1,2], while a value of "
                                                   + std::to_string(variant_adptv) + " was detected.");
        }
        if (allowed_variants.empty() || allowed_variants.size() > 18u) {
            pagmo_throw(std::invalid_argument, "cuda_de1220: between 1 and 18 allowed variants are required");
        }
        keep_memory(memory);
        m_desc.n_allowed = static_cast<unsigned>(allowed_variants.size());
        for (std::size_t i = 0; i < allowed_variants.size(); ++i) m_desc.allowed_variants[i] = allowed_variants[i];
        m_desc.variant_adptv = variant_adptv, m_desc.ftol = ftol, m_desc.xtol = xtol;
    }
    // Gen, Fevals, Best, F, CR, Variant, dx, df (de1220.hpp:142)
    using log_line_type = std::tuple<unsigned, unsigned long long, double, double, double, unsigned, double, double>;
    using log_type = std::vector<log_line_type>;
    log_type get_log() const
    {
        return typed_log<log_line_type>([](const double *r) {
            return log_line_type(static_cast<unsigned>(r[0]), static_cast<unsigned long long>(r[1]), r[2], r[3], r[4], static_cast<unsigned>(r[5]),
                                 r[6], r[7]);
        });
    }
};

class cuda_pso_gen : public cuda_algorithm_base
{
public:
    cuda_pso_gen(unsigned gen = 1u, double omega = 0.7298, double eta1 = 2.05, double eta2 = 2.05, double max_vel = 0.5, unsigned variant = 5u,
                 unsigned neighb_type = 2u, unsigned neighb_param = 4u, bool memory = false, unsigned seed = pagmo::random_device::next(),
                 int device = 0)
        : cuda_algorithm_base(PGC_ALGO_PSO_GEN, "GPSO: Generational Particle Swarm Optimization", gen, seed, device)
    {
        keep_memory(memory);
        m_desc.omega = omega, m_desc.eta1 = eta1, m_desc.eta2 = eta2, m_desc.max_vel = max_vel, m_desc.variant = variant;
        m_desc.neighb_type = neighb_type, m_desc.neighb_param = neighb_param; // range checks: pso.cu (pso_gen.cpp:69-107)
    }
    // Gen, Fevals, gbest, Mean Vel., Mean lbest, Avg. Dist. (pso_gen.hpp:111)
    using log_line_type = std::tuple<unsigned, unsigned long long, double, double, double, double>;
    using log_type = std::vector<log_line_type>;
    log_type get_log() const
    {
        return typed_log<log_line_type>([](const double *r) {
            return log_line_type(static_cast<unsigned>(r[0]), static_cast<unsigned long long>(r[1]), r[2], r[3], r[4], r[5]);
        });
    }
};

// pagmo::pso (pso.hpp:110, same constructor arguments as pso_gen): the reference's pso updates the swarm particle by particle
// (pso.cpp:115-443); on the device every generation is batched, i.e. it runs as pso_gen does.
class cuda_pso : public cuda_pso_gen
{
public:
    using cuda_pso_gen::cuda_pso_gen;
    std::string get_name() const
    {
        return "PSO: Particle Swarm Optimization [CUDA sm_100a, generational]";
    }
};

class cuda_nsga2 : public cuda_algorithm_base
{
public:
    cuda_nsga2(unsigned gen = 1u, double cr = 0.95, double eta_c = 10., double m = 0.01, double eta_m = 50.,
               unsigned seed = pagmo::random_device::next(), int device = 0)
        : cuda_algorithm_base(PGC_ALGO_NSGA2, "NSGA-II:", gen, seed, device)
    {
        m_desc.cr = cr, m_desc.eta_c = eta_c, m_desc.m = m, m_desc.eta_m = eta_m; // range checks: nsga2.cu (nsga2.cpp:71-86)
    }
    using log_line_type = std::tuple<unsigned, unsigned long long, pagmo::vector_double>; // Gen, Fevals, ideal point
    using log_type = std::vector<log_line_type>;
    log_type get_log() const
    {
        const std::size_t len = m_log_row_len;
        return typed_log<log_line_type>([len](const double *r) {
            return log_line_type(static_cast<unsigned>(r[0]), static_cast<unsigned long long>(r[1]), pagmo::vector_double(r + 2, r + len));
        });
    }
};

class cuda_sga : public cuda_algorithm_base
{
public:
    cuda_sga(unsigned gen = 1u, double cr = .90, double eta_c = 1., double m = 0.02, double param_m = 1., unsigned param_s = 2u,
             std::string crossover = "exponential", std::string mutation = "polynomial", std::string selection = "tournament",
             unsigned seed = pagmo::random_device::next(), int device = 0)
        : cuda_algorithm_base(PGC_ALGO_SGA, "SGA: Genetic Algorithm", gen, seed, device)
    {
        // the string -> strategy maps of sga.cpp:78-108; the numeric range checks (sga.cpp:116-160) are made by pgc_sga_evolve_device
        const auto pick = [](const std::string &what, const std::string &v, std::initializer_list<const char *> names) -> unsigned {
            unsigned k = 0;
            for (const char *n : names) {
                if (v == n) return k;
                ++k;
            }
            pagmo_throw(std::invalid_argument, "The " + what + " type is unknown: " + v); // sga.cpp:138-155
        };
        m_desc.crossover = pick("crossover", crossover, {"exponential", "binomial", "single", "sbx"});
        m_desc.mutation = pick("mutation", mutation, {"gaussian", "uniform", "polynomial"});
        m_desc.selection = pick("selection", selection, {"tournament", "truncated"});
        m_desc.cr = cr, m_desc.eta_c = eta_c, m_desc.m = m, m_desc.param_m = param_m, m_desc.param_s = param_s;
    }
    using log_line_type = std::tuple<unsigned, unsigned long long, double, double>; // Gen, Fevals, Best, Improvement (sga.hpp:151)
    using log_type = std::vector<log_line_type>;
    log_type get_log() const
    {
        return typed_log<log_line_type>([](const double *r) {
            return log_line_type(static_cast<unsigned>(r[0]), static_cast<unsigned long long>(r[1]), r[2], r[3]);
        });
    }
};

// pagmo::cmaes (cmaes.hpp:110).  Sampling, evaluation, recombination and the rank-mu matrix run on the device; the evolution paths and
// the eigendecomposition of C (a Jacobi solver where the reference uses Eigen) on the host: pgc_cmaes_evolve_device.
class cuda_cmaes : public cuda_algorithm_base
{
public:
    cuda_cmaes(unsigned gen = 1u, double cc = -1, double cs = -1, double c1 = -1, double cmu = -1, double sigma0 = 0.5, double ftol = 1e-6,
               double xtol = 1e-6, bool memory = false, bool force_bounds = false, unsigned seed = pagmo::random_device::next(), int device = 0)
        : cuda_algorithm_base(PGC_ALGO_CMAES, "CMA-ES: Covariance Matrix Adaptation Evolutionary Strategy", gen, seed, device)
    {
        if (((cc < 0.) || (cc > 1.)) && !(cc == -1)) { // cmaes.cpp:64-67
            pagmo_throw(std::invalid_argument, "cc must be in [0,1] or -1 if its value has to be initialized automatically, a value of "
                                                   + std::to_string(cc) + " was detected");
        }
        if (((cs < 0.) || (cs > 1.)) && !(cs == -1)) {
            pagmo_throw(std::invalid_argument, "cs needs to be in [0,1] or -1 if its value has to be initialized automatically, a value of "
                                                   + std::to_string(cs) + " was detected");
        }
        if (((c1 < 0.) || (c1 > 1.)) && !(c1 == -1)) {
            pagmo_throw(std::invalid_argument, "c1 needs to be in [0,1] or -1 if its value has to be initialized automatically, a value of "
                                                   + std::to_string(c1) + " was detected");
        }
        if (((cmu < 0.) || (cmu > 1.)) && !(cmu == -1)) {
            pagmo_throw(std::invalid_argument, "cmu needs to be in [0,1] or -1 if its value has to be initialized automatically, a value of "
                                                   + std::to_string(cmu) + " was detected");
        }
        keep_memory(memory);
        m_desc.cma_cc = cc, m_desc.cma_cs = cs, m_desc.cma_c1 = c1, m_desc.cma_cmu = cmu, m_desc.sigma0 = sigma0;
        m_desc.ftol = ftol, m_desc.xtol = xtol, m_desc.force_bounds = force_bounds ? 1u : 0u;
    }
    using log_line_type = std::tuple<unsigned, unsigned long long, double, double, double, double>; // Gen, Fevals, Best, dx, df, sigma
    using log_type = std::vector<log_line_type>;
    log_type get_log() const
    {
        return typed_log<log_line_type>([](const double *r) {
            return log_line_type(static_cast<unsigned>(r[0]), static_cast<unsigned long long>(r[1]), r[2], r[3], r[4], r[5]);
        });
    }
};

// pagmo::xnes (xnes.hpp:107-108).  Sampling, evaluation and the natural-gradient contractions run on the device; the mean / A / sigma
// updates (exp of the symmetric d_A through a Jacobi solver where the reference uses Eigen's matrix exponential) on the host:
// pgc_xnes_evolve_device.  eta_mu, eta_sigma, eta_b travel in the descriptor's cma_cc, cma_cs, cma_c1.
class cuda_xnes : public cuda_algorithm_base
{
public:
    cuda_xnes(unsigned gen = 1u, double eta_mu = -1, double eta_sigma = -1, double eta_b = -1, double sigma0 = -1, double ftol = 1e-6,
              double xtol = 1e-6, bool memory = false, bool force_bounds = false, unsigned seed = pagmo::random_device::next(), int device = 0)
        : cuda_algorithm_base(PGC_ALGO_XNES, "xNES: Exponential Natural Evolution Strategies", gen, seed, device)
    {
        const auto check_eta = [](double v, const char *name, const char *verb) { // xnes.cpp:55-78
            if (((v <= 0.) || (v > 1.)) && !(v == -1)) {
                pagmo_throw(std::invalid_argument, std::string(name) + verb + " in ]0,1] or -1 if its value has to be initialized automatically, a value of "
                                                       + std::to_string(v) + " was detected");
            }
        };
        check_eta(eta_mu, "eta_mu", " must be");
        check_eta(eta_sigma, "eta_sigma", " needs to be");
        check_eta(eta_b, "eta_b", " needs to be");
        check_eta(sigma0, "sigma0", " needs to be");
        keep_memory(memory);
        m_desc.cma_cc = eta_mu, m_desc.cma_cs = eta_sigma, m_desc.cma_c1 = eta_b, m_desc.sigma0 = sigma0;
        m_desc.ftol = ftol, m_desc.xtol = xtol, m_desc.force_bounds = force_bounds ? 1u : 0u;
    }
    using log_line_type = std::tuple<unsigned, unsigned long long, double, double, double, double>; // Gen, Fevals, Best, dx, df, sigma
    using log_type = std::vector<log_line_type>;
    log_type get_log() const
    {
        return typed_log<log_line_type>([](const double *r) {
            return log_line_type(static_cast<unsigned>(r[0]), static_cast<unsigned long long>(r[1]), r[2], r[3], r[4], r[5]);
        });
    }
};

// pagmo::nspso (nspso.hpp:59-62): non-dominated sorting PSO; generational in the reference itself, so this is the same algorithm on
// Philox draws (pgc_nspso_evolve_device, nspso.cu).  memory = true keeps the velocities and the archive between evolve() calls.
class cuda_nspso : public cuda_algorithm_base
{
public:
    cuda_nspso(unsigned gen = 1u, double omega = 0.6, double c1 = 2.0, double c2 = 2.0, double chi = 1.0, double v_coeff = 0.5,
               unsigned leader_selection_range = 60u, std::string diversity_mechanism = "crowding distance", bool memory = false,
               unsigned seed = pagmo::random_device::next(), int device = 0)
        : cuda_algorithm_base(PGC_ALGO_NSPSO, "NSPSO", gen, seed, device)
    {
        if (omega < 0. || omega > 1.) { // nspso.cpp:58-62
            pagmo_throw(std::invalid_argument, "The particles' inertia weight must be in the [0,1] range, while a value of "
                                                   + std::to_string(omega) + " was detected");
        }
        if (c1 <= 0 || c2 <= 0 || chi <= 0) { // :63-66
            pagmo_throw(std::invalid_argument, "first and second magnitude of the force "
                                               "coefficients and velocity scaling factor should be greater than 0");
        }
        if (v_coeff <= 0 || v_coeff > 1) { // :67-70
            pagmo_throw(std::invalid_argument, "velocity scaling factor should be in ]0,1] range, while a value of" + std::to_string(v_coeff)
                                                   + " was detected");
        }
        if (leader_selection_range > 100) { // :71-75
            pagmo_throw(std::invalid_argument, "leader selection range coefficient should be in the ]0,100] range, while a value of"
                                                   + std::to_string(leader_selection_range) + " was detected");
        }
        unsigned div = 0;
        if (diversity_mechanism == "crowding distance") div = 0;
        else if (diversity_mechanism == "niche count") div = 1;
        else if (diversity_mechanism == "max min") div = 2;
        else pagmo_throw(std::invalid_argument, "Non existing diversity mechanism method."); // :76-80
        keep_memory(memory);
        m_desc.omega = omega, m_desc.nspso_c1 = c1, m_desc.nspso_c2 = c2, m_desc.nspso_chi = chi, m_desc.nspso_v_coeff = v_coeff;
        m_desc.leader_selection_range = leader_selection_range, m_desc.diversity = div;
    }
    using log_line_type = std::tuple<unsigned, unsigned long long, pagmo::vector_double>; // Gen, Fevals, ideal point
    using log_type = std::vector<log_line_type>;
    log_type get_log() const
    {
        const std::size_t len = m_log_row_len;
        return typed_log<log_line_type>([len](const double *r) {
            return log_line_type(static_cast<unsigned>(r[0]), static_cast<unsigned long long>(r[1]), pagmo::vector_double(r + 2, r + len));
        });
    }
};

// pagmo::moead_gen (moead_gen.hpp), the reference's generational MOEA/D: the weight vectors and their neighbourhoods are computed
// with pagmo's own decomposition_weights / kNN (on an engine seeded like the reference's, so "random" weights are the reference's),
// candidate construction, batch evaluation and the sequential insertion run on the device (pgc_moead_gen_evolve_device, moead.cu).
class cuda_moead_gen
{
public:
    cuda_moead_gen(unsigned gen = 1u, std::string weight_generation = "grid", std::string decomposition = "tchebycheff",
                   pagmo::population::size_type neighbours = 20u, double CR = 1.0, double F = 0.5, double eta_m = 20., double realb = 0.9,
                   unsigned limit = 2u, bool preserve_diversity = true, unsigned seed = pagmo::random_device::next(), int device = 0)
        : m_gen(gen), m_weight_generation(std::move(weight_generation)), m_decomposition(std::move(decomposition)), m_neighbours(neighbours),
          m_CR(CR), m_F(F), m_eta_m(eta_m), m_realb(realb), m_limit(limit), m_preserve_diversity(preserve_diversity), m_seed(seed),
          m_device(device), m_cache(std::make_shared<detail::twin_cache>())
    { // moead_gen.cpp:60-104
        if (m_weight_generation != "random" && m_weight_generation != "grid" && m_weight_generation != "low discrepancy") {
            pagmo_throw(std::invalid_argument, "Weight generation method requested is '" + m_weight_generation
                                                   + "', but only one of 'random', 'low discrepancy', 'grid' is allowed");
        }
        if (m_decomposition != "tchebycheff" && m_decomposition != "weighted" && m_decomposition != "bi") {
            pagmo_throw(std::invalid_argument, "Weight generation method requested is '" + m_decomposition
                                                   + "', but only one of 'tchebycheff', 'weighted', 'bi' is allowed");
        }
        if (CR > 1.0 || CR < 0.) {
            pagmo_throw(std::invalid_argument,
                        "The parameter CR (used by the differential evolution operator) needs to be in [0,1], while a value of "
                            + std::to_string(CR) + " was detected");
        }
        if (F > 1.0 || F < 0.) {
            pagmo_throw(std::invalid_argument,
                        "The parameter F (used by the differential evolution operator) needs to be in [0,1], while a value of "
                            + std::to_string(F) + " was detected");
        }
        if (eta_m < 0.) {
            pagmo_throw(std::invalid_argument, "The distribution index for the polynomial mutation (eta_m) needs to be positive, while a value of "
                                                   + std::to_string(eta_m) + " was detected");
        }
        if (realb > 1.0 || realb < 0.) {
            pagmo_throw(std::invalid_argument, "The chance of considering a neighbourhood (realb) needs to be in [0,1], while a value of "
                                                   + std::to_string(realb) + " was detected");
        }
        if (neighbours < 2) {
            pagmo_throw(std::invalid_argument,
                        "The size of the weight's neighborhood needs to be >= 2, while a size of " + std::to_string(neighbours) + " was detected");
        }
    }
    pagmo::population evolve(pagmo::population pop) const
    {
        const auto &prob = pop.get_problem();
        const auto NP = pop.size();
        if (!NP) pagmo_throw(std::invalid_argument, get_name() + " cannot work on an empty population"); // :146-148
        if (prob.get_nf() < 2u) {
            pagmo_throw(std::invalid_argument, "This is a multiobjective algorithm, while number of objectives detected in " + prob.get_name()
                                                   + " is " + std::to_string(prob.get_nf()));
        }
        if (prob.get_nc() != 0u) {
            pagmo_throw(std::invalid_argument, "Non linear constraints detected in " + prob.get_name() + " instance. " + get_name()
                                                   + " cannot deal with them");
        }
        if (prob.is_stochastic()) {
            pagmo_throw(std::invalid_argument, "The problem appears to be stochastic " + get_name() + " cannot deal with it");
        }
        if (m_neighbours > NP - 1u) {
            pagmo_throw(std::invalid_argument, "The neighbourhood size specified (T) is " + std::to_string(m_neighbours)
                                                   + ": too large for the input population having size " + std::to_string(NP));
        }
        if (m_gen == 0u) return pop;
        const auto h = m_cache->find(prob, m_device);
        if (!h) {
            pagmo_throw(std::invalid_argument, get_name() + " cannot evolve a population of '" + prob.get_name()
                                                   + "': no CUDA evaluator exists for this UDP type; there is no CPU fallback");
        }
        // weights and neighbourhoods exactly as the reference computes them at the top of evolve() (:155, :168)
        std::mt19937 engine(m_seed);
        const auto weights = pagmo::decomposition_weights(prob.get_nf(), NP, m_weight_generation, engine);
        const auto neigh = pagmo::kNN(weights, m_neighbours);
        const auto nx = prob.get_nx(), nf = prob.get_nf();
        pagmo::vector_double w(NP * nf), x(NP * nx), f(NP * nf);
        std::vector<uint32_t> nb(NP * m_neighbours);
        for (decltype(pop.size()) i = 0; i < NP; ++i) {
            std::copy(weights[i].begin(), weights[i].end(), w.begin() + static_cast<std::ptrdiff_t>(i * nf));
            for (decltype(m_neighbours) j = 0; j < m_neighbours; ++j) nb[i * m_neighbours + j] = static_cast<uint32_t>(neigh[i][j]);
            std::copy(pop.get_x()[i].begin(), pop.get_x()[i].end(), x.begin() + static_cast<std::ptrdiff_t>(i * nx));
            std::copy(pop.get_f()[i].begin(), pop.get_f()[i].end(), f.begin() + static_cast<std::ptrdiff_t>(i * nf));
        }
        const int method = m_decomposition == "weighted" ? 0 : (m_decomposition == "tchebycheff" ? 1 : 2);
        h->on_device(x, f, "pgc_moead_gen_evolve_device", [&](double *dx, double *df, std::size_t n) {
            return detail::with_log_capture(h->context(), m_verbosity, m_gen, 3u + nf, m_log_rows, [&] {
                return pgc_moead_gen_evolve_device(h->raw(), dx, df, n, m_gen, w.data(), nb.data(), static_cast<unsigned>(m_neighbours), method, m_CR,
                                                   m_F, m_eta_m, m_realb, m_limit, m_preserve_diversity ? 1 : 0, m_seed, m_generation, nullptr);
            });
        });
        m_log_row_len = 3u + nf;
        m_generation += m_gen;
        for (decltype(pop.size()) i = 0; i < NP; ++i) {
            pop.set_xf(i, pagmo::vector_double(x.begin() + static_cast<std::ptrdiff_t>(i * nx), x.begin() + static_cast<std::ptrdiff_t>((i + 1) * nx)),
                       pagmo::vector_double(f.begin() + static_cast<std::ptrdiff_t>(i * nf), f.begin() + static_cast<std::ptrdiff_t>((i + 1) * nf)));
        }
        prob.increment_fevals(static_cast<unsigned long long>(m_gen) * NP);
        return pop;
    }
    void set_seed(unsigned seed) { m_seed = seed; }
    unsigned get_seed() const { return m_seed; }
    unsigned get_gen() const { return m_gen; }
    std::string get_name() const { return "MOEAD-GEN: MOEA/D - DE [CUDA sm_100a]"; }
    void set_verbosity(unsigned level) { m_verbosity = level; }
    unsigned get_verbosity() const { return m_verbosity; }
    using log_line_type = std::tuple<unsigned, unsigned long long, double, pagmo::vector_double>; // Gen, Fevals, ADF, ideal point (moead_gen.hpp:76)
    using log_type = std::vector<log_line_type>;
    log_type get_log() const
    {
        log_type out;
        for (std::size_t r = 0; m_log_row_len && r < m_log_rows.size() / m_log_row_len; ++r) {
            const double *v = m_log_rows.data() + r * m_log_row_len;
            out.emplace_back(static_cast<unsigned>(v[0]), static_cast<unsigned long long>(v[1]), v[2], pagmo::vector_double(v + 3, v + m_log_row_len));
        }
        return out;
    }
    pagmo::thread_safety get_thread_safety() const { return pagmo::thread_safety::basic; }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_gen, m_weight_generation, m_decomposition, m_neighbours, m_CR, m_F, m_eta_m, m_realb, m_limit,
                               m_preserve_diversity, m_seed, m_device, m_generation);
    }

private:
    unsigned m_gen;
    std::string m_weight_generation, m_decomposition;
    pagmo::population::size_type m_neighbours;
    double m_CR, m_F, m_eta_m, m_realb;
    unsigned m_limit;
    bool m_preserve_diversity;
    unsigned m_seed;
    int m_device;
    mutable unsigned m_generation = 1;
    unsigned m_verbosity = 0;
    mutable pagmo::vector_double m_log_rows;
    mutable std::size_t m_log_row_len = 0;
    std::shared_ptr<detail::twin_cache> m_cache;
};

// pagmo::gaco (gaco.hpp:104-107), extended ant colony optimisation: penalties, the solution archive, the pheromone values and the ants
// of every generation are computed on the device (pgc_gaco_evolve_device, gaco.cu) - the same generational algorithm as the
// reference's, on Philox draws.  Unconstrained single-objective problems (the ones with a device evaluator).  Like the reference, the object keeps its oracle parameter, kernel width and stopping counters between evolve() calls.
class cuda_gaco
{
public:
    cuda_gaco(unsigned gen = 1u, unsigned ker = 63u, double q = 1.0, double oracle = 0., double acc = 0.01, unsigned threshold = 1u,
              unsigned n_gen_mark = 7u, unsigned impstop = 100000u, unsigned evalstop = 100000u, double focus = 0., bool memory = false,
              unsigned seed = pagmo::random_device::next(), int device = 0)
        : m_gen(gen), m_ker(ker), m_q(q), m_oracle(oracle), m_acc(acc), m_threshold(threshold), m_n_gen_mark(n_gen_mark), m_impstop(impstop),
          m_evalstop(evalstop), m_focus(focus), m_memory(memory), m_seed(seed), m_device(device), m_cache(std::make_shared<detail::twin_cache>())
    { // gaco.cpp:62-94
        if (acc < 0.) {
            pagmo_throw(std::invalid_argument, "The accuracy parameter must be >=0, while a value of " + std::to_string(acc) + " was detected");
        }
        if (focus < 0.) {
            pagmo_throw(std::invalid_argument, "The focus parameter must be >=0  while a value of " + std::to_string(focus) + " was detected");
        }
        if ((threshold < 1 || threshold > gen) && gen != 0 && memory == false) {
            pagmo_throw(std::invalid_argument, "If memory is inactive, the threshold parameter must be either in [1,m_gen] while a value of "
                                                   + std::to_string(threshold) + " was detected");
        }
        if (threshold < 1 && gen != 0 && memory == true) {
            pagmo_throw(std::invalid_argument,
                        "If memory is active, the threshold parameter must be >=1 while a value of " + std::to_string(threshold) + " was detected");
        }
        if (q < 0.) {
            pagmo_throw(std::invalid_argument,
                        "The convergence speed parameter must be >=0  while a value of " + std::to_string(q) + " was detected");
        }
        if (ker < 2u) {
            pagmo_throw(std::invalid_argument, "The ker size parameter must be >=2  while a value of " + std::to_string(ker) + " was detected");
        }
    }
    pagmo::population evolve(pagmo::population pop) const
    {
        const auto &prob = pop.get_problem();
        const auto NP = pop.size();
        if (!NP) return pop; // gaco.cpp:147-149
        if (prob.is_stochastic()) {
            pagmo_throw(std::invalid_argument, "The problem appears to be stochastic " + get_name() + " cannot deal with it");
        }
        if (m_gen == 0u) return pop;
        if (NP < 2u) {
            pagmo_throw(std::invalid_argument,
                        get_name() + " needs at least 2 individuals in the population, " + std::to_string(NP) + " detected");
        }
        if (m_ker > NP) {
            pagmo_throw(std::invalid_argument, get_name() + " cannot work with a solution archive bigger than the population size");
        }
        if (prob.get_nobj() != 1u) {
            pagmo_throw(std::invalid_argument,
                        "Multiple objectives detected in " + prob.get_name() + " instance. " + get_name() + " cannot deal with them");
        }
        if (prob.get_nc() != 0u) {
            pagmo_throw(std::invalid_argument, get_name() + ": constrained problems have no device evaluator; there is no CPU fallback");
        }
        const auto h = m_cache->find(prob, m_device);
        if (!h) {
            pagmo_throw(std::invalid_argument, get_name() + " cannot evolve a population of '" + prob.get_name()
                                                   + "': no CUDA evaluator exists for this UDP type; there is no CPU fallback");
        }
        const auto nx = prob.get_nx();
        pagmo::vector_double x(NP * nx), f(NP);
        for (decltype(pop.size()) i = 0; i < NP; ++i) {
            std::copy(pop.get_x()[i].begin(), pop.get_x()[i].end(), x.begin() + static_cast<std::ptrdiff_t>(i * nx));
            f[i] = pop.get_f()[i][0];
        }
        unsigned done = 0;
        // the algorithm object's members travel in m_state; with memory = true also its archive (a value member: copies of the algorithm
        // carry it).  The evalstop counter watches the POPULATION's champion (gaco.cpp:338-347), which the population brings along.
        m_state.memory = m_memory ? 1u : 0u;
        if (m_memory) {
            if (m_archive.size() != static_cast<std::size_t>(m_ker) * (nx + 2u)) {
                m_archive.assign(static_cast<std::size_t>(m_ker) * (nx + 2u), 0.);
                m_state.counter = 0u; // another problem dimension: the archive starts again
            }
            m_state.h_archive = m_archive.data();
            m_state.h_archive_len = m_archive.size();
        }
        m_state.has_champion = 1u;
        m_state.champion_f = pop.champion_f()[0];
        h->on_device(x, f, "pgc_gaco_evolve_device", [&](double *dx, double *df, std::size_t n) {
            return detail::with_log_capture(h->context(), m_verbosity, m_gen, 7u, m_log_rows, [&] {
                return pgc_gaco_evolve_device(h->raw(), dx, df, n, m_gen, m_ker, m_q, m_oracle, m_acc, m_threshold, m_n_gen_mark, m_impstop,
                                              m_evalstop, m_focus, m_seed, m_generation, &m_state, &done, nullptr);
            });
        });
        m_generation += m_gen;
        for (decltype(pop.size()) i = 0; i < NP; ++i) {
            pop.set_xf(i, pagmo::vector_double(x.begin() + static_cast<std::ptrdiff_t>(i * nx), x.begin() + static_cast<std::ptrdiff_t>((i + 1) * nx)),
                       pagmo::vector_double{f[i]});
        }
        prob.increment_fevals(static_cast<unsigned long long>(done) * NP);
        return pop;
    }
    void set_seed(unsigned seed) { m_seed = seed; }
    unsigned get_seed() const { return m_seed; }
    unsigned get_gen() const { return m_gen; }
    // the oracle parameter as the runs so far have left it (gaco.cpp:349-402)
    double get_oracle() const { return m_state.initialized ? m_state.oracle : m_oracle; }
    void set_verbosity(unsigned level) { m_verbosity = level; }
    unsigned get_verbosity() const { return m_verbosity; }
    // Gen, Fevals, Best, Kernel, Oracle, dx, dp (gaco.hpp:91)
    using log_line_type = std::tuple<unsigned, unsigned long long, double, unsigned, double, double, double>;
    using log_type = std::vector<log_line_type>;
    log_type get_log() const
    {
        log_type out;
        for (std::size_t r = 0; r < m_log_rows.size() / 7u; ++r) {
            const double *v = m_log_rows.data() + r * 7u;
            out.emplace_back(static_cast<unsigned>(v[0]), static_cast<unsigned long long>(v[1]), v[2], static_cast<unsigned>(v[3]), v[4], v[5], v[6]);
        }
        return out;
    }
    std::string get_name() const { return "GACO: Ant Colony Optimization [CUDA sm_100a]"; }
    pagmo::thread_safety get_thread_safety() const { return pagmo::thread_safety::basic; }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_gen, m_ker, m_q, m_oracle, m_acc, m_threshold, m_n_gen_mark, m_impstop, m_evalstop, m_focus, m_memory, m_seed,
                               m_device, m_generation, m_state.oracle, m_state.q, m_state.n_evalstop, m_state.n_impstop, m_state.gen_mark,
                               m_state.initialized, m_state.fevals, m_state.counter, m_archive);
    }

private:
    unsigned m_gen, m_ker;
    double m_q, m_oracle, m_acc;
    unsigned m_threshold, m_n_gen_mark, m_impstop, m_evalstop;
    double m_focus;
    bool m_memory;
    unsigned m_seed;
    int m_device;
    mutable unsigned m_generation = 1;
    unsigned m_verbosity = 0;
    mutable pagmo::vector_double m_log_rows, m_archive;
    mutable pgc_gaco_state m_state{};
    std::shared_ptr<detail::twin_cache> m_cache;
};

// pagmo::maco (maco.hpp:107-109), multi-objective hypervolume-based ant colony optimisation: non-dominated sorting, the hypervolume
// contributions of the fronts, the ants and their evaluation run on the device; the host walks the fronts to rebuild the archive
// (pgc_maco_evolve_device, gaco.cu).  memory = true is not built.
class cuda_maco
{
public:
    cuda_maco(unsigned gen = 1u, unsigned ker = 63u, double q = 1.0, unsigned threshold = 1u, unsigned n_gen_mark = 7u,
              unsigned evalstop = 100000u, double focus = 0., bool memory = false, unsigned seed = pagmo::random_device::next(), int device = 0)
        : m_gen(gen), m_ker(ker), m_q(q), m_threshold(threshold), m_n_gen_mark(n_gen_mark), m_evalstop(evalstop), m_focus(focus), m_seed(seed),
          m_device(device), m_cache(std::make_shared<detail::twin_cache>())
    { // maco.cpp:64-83
        if (focus < 0.) {
            pagmo_throw(std::invalid_argument, "The focus parameter must be >=0  while a value of " + std::to_string(focus) + " was detected");
        }
        if (memory) pagmo_throw(std::invalid_argument, "cuda_maco: memory = true is not supported on the device path");
        if ((threshold < 1 || threshold > gen) && gen != 0) {
            pagmo_throw(std::invalid_argument, "If memory is inactive, the threshold parameter must be either in [1,m_gen] while a value of "
                                                   + std::to_string(threshold) + " was detected");
        }
    }
    pagmo::population evolve(pagmo::population pop) const
    {
        const auto &prob = pop.get_problem();
        const auto NP = pop.size();
        if (!NP) pagmo_throw(std::invalid_argument, get_name() + " cannot work on an empty population"); // maco.cpp:126-152
        if (prob.is_stochastic()) {
            pagmo_throw(std::invalid_argument, "The problem appears to be stochastic " + get_name() + " cannot deal with it");
        }
        if (m_gen == 0u) return pop;
        if (m_ker > NP) {
            pagmo_throw(std::invalid_argument, get_name() + " cannot work with a solution archive bigger than the population size");
        }
        if (prob.get_nc() != 0u) {
            pagmo_throw(std::invalid_argument,
                        "Non linear constraints detected in " + prob.get_name() + " instance. " + get_name() + " cannot deal with them.");
        }
        if (prob.get_nf() < 2u) {
            pagmo_throw(std::invalid_argument, "This is a multiobjective algorithm, while number of objectives detected in " + prob.get_name()
                                                   + " is " + std::to_string(prob.get_nf()));
        }
        const auto h = m_cache->find(prob, m_device);
        if (!h) {
            pagmo_throw(std::invalid_argument, get_name() + " cannot evolve a population of '" + prob.get_name()
                                                   + "': no CUDA evaluator exists for this UDP type; there is no CPU fallback");
        }
        const auto nx = prob.get_nx(), nf = prob.get_nf();
        pagmo::vector_double x(NP * nx), f(NP * nf);
        for (decltype(pop.size()) i = 0; i < NP; ++i) {
            std::copy(pop.get_x()[i].begin(), pop.get_x()[i].end(), x.begin() + static_cast<std::ptrdiff_t>(i * nx));
            std::copy(pop.get_f()[i].begin(), pop.get_f()[i].end(), f.begin() + static_cast<std::ptrdiff_t>(i * nf));
        }
        unsigned done = 0;
        h->on_device(x, f, "pgc_maco_evolve_device", [&](double *dx, double *df, std::size_t n) {
            return detail::with_log_capture(h->context(), m_verbosity, m_gen, 2u + nf, m_log_rows, [&] {
                return pgc_maco_evolve_device(h->raw(), dx, df, n, m_gen, m_ker, m_q, m_threshold, m_n_gen_mark, m_evalstop, m_focus, m_seed,
                                              m_generation, &m_state, &done, nullptr);
            });
        });
        m_log_row_len = 2u + nf;
        m_generation += m_gen;
        for (decltype(pop.size()) i = 0; i < NP; ++i) {
            pop.set_xf(i, pagmo::vector_double(x.begin() + static_cast<std::ptrdiff_t>(i * nx), x.begin() + static_cast<std::ptrdiff_t>((i + 1) * nx)),
                       pagmo::vector_double(f.begin() + static_cast<std::ptrdiff_t>(i * nf), f.begin() + static_cast<std::ptrdiff_t>((i + 1) * nf)));
        }
        prob.increment_fevals(static_cast<unsigned long long>(done) * NP);
        return pop;
    }
    void set_seed(unsigned seed) { m_seed = seed; }
    unsigned get_seed() const { return m_seed; }
    unsigned get_gen() const { return m_gen; }
    std::string get_name() const { return "MHACO: Multi-objective Hypervolume-based Ant Colony Optimization [CUDA sm_100a]"; }
    void set_verbosity(unsigned level) { m_verbosity = level; }
    unsigned get_verbosity() const { return m_verbosity; }
    using log_line_type = std::tuple<unsigned, unsigned long long, pagmo::vector_double>; // Gen, Fevals, ideal point (maco.hpp:93)
    using log_type = std::vector<log_line_type>;
    log_type get_log() const
    {
        log_type out;
        for (std::size_t r = 0; m_log_row_len && r < m_log_rows.size() / m_log_row_len; ++r) {
            const double *v = m_log_rows.data() + r * m_log_row_len;
            out.emplace_back(static_cast<unsigned>(v[0]), static_cast<unsigned long long>(v[1]), pagmo::vector_double(v + 2, v + m_log_row_len));
        }
        return out;
    }
    pagmo::thread_safety get_thread_safety() const { return pagmo::thread_safety::basic; }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_gen, m_ker, m_q, m_threshold, m_n_gen_mark, m_evalstop, m_focus, m_seed, m_device, m_generation, m_state.q,
                               m_state.n_evalstop, m_state.gen_mark, m_state.initialized);
    }

private:
    unsigned m_gen, m_ker;
    double m_q;
    unsigned m_threshold, m_n_gen_mark, m_evalstop;
    double m_focus;
    unsigned m_seed;
    int m_device;
    mutable unsigned m_generation = 1;
    unsigned m_verbosity = 0;
    mutable pagmo::vector_double m_log_rows;
    mutable std::size_t m_log_row_len = 0;
    mutable pgc_maco_state m_state{};
    std::shared_ptr<detail::twin_cache> m_cache;
};

} // namespace pagmo_cuda

PAGMO_S11N_ALGORITHM_EXPORT_KEY(pagmo_cuda::cuda_gaco)
PAGMO_S11N_ALGORITHM_EXPORT_KEY(pagmo_cuda::cuda_maco)
PAGMO_S11N_ALGORITHM_EXPORT_KEY(pagmo_cuda::cuda_moead_gen)
PAGMO_S11N_ALGORITHM_EXPORT_KEY(pagmo_cuda::cuda_nspso)
PAGMO_S11N_ALGORITHM_EXPORT_KEY(pagmo_cuda::cuda_cmaes)
PAGMO_S11N_ALGORITHM_EXPORT_KEY(pagmo_cuda::cuda_xnes)
PAGMO_S11N_ALGORITHM_EXPORT_KEY(pagmo_cuda::cuda_sga)
PAGMO_S11N_ALGORITHM_EXPORT_KEY(pagmo_cuda::cuda_de)
PAGMO_S11N_ALGORITHM_EXPORT_KEY(pagmo_cuda::cuda_sade)
PAGMO_S11N_ALGORITHM_EXPORT_KEY(pagmo_cuda::cuda_de1220)
PAGMO_S11N_ALGORITHM_EXPORT_KEY(pagmo_cuda::cuda_pso_gen)
PAGMO_S11N_ALGORITHM_EXPORT_KEY(pagmo_cuda::cuda_pso)
PAGMO_S11N_ALGORITHM_EXPORT_KEY(pagmo_cuda::cuda_nsga2)

#endif
