// tests/cpp/test_adapters.cpp - the drop-in boundary exercised through pagmo's OWN type-erased classes.
// Compiled against the unmodified reference headers (+ oracle/shim for the absent Boost/TBB) and linked with
// oracle/_ref/libpagmo_ref.so (pagmo::problem, pagmo::bfe, pagmo::population, thread_bfe, the stock UDPs) and
// pagmo2_b200/libpgc.so.  Mirrors the reference's bfe-equivalence tests: tests/thread_bfe.cpp:66-97,
// tests/default_bfe.cpp, tests/bfe.cpp.  Run on the GPU box by tests/test_gpu_adapters.py.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <pagmo/batch_evaluators/default_bfe.hpp>
#include <pagmo/batch_evaluators/member_bfe.hpp>
#include <pagmo/batch_evaluators/thread_bfe.hpp>
#include <pagmo/bfe.hpp>
#include <pagmo/population.hpp>
#include <pagmo/problem.hpp>
#include <pagmo/algorithms/nsga2.hpp>
#include <pagmo/problems/cec2014.hpp>
#include <pagmo/problems/decompose.hpp>
#include <pagmo/problems/translate.hpp>
#include <pagmo/problems/unconstrain.hpp>
#include <pagmo/problems/dtlz.hpp>
#include <pagmo/problems/lennard_jones.hpp>
#include <pagmo/problems/rastrigin.hpp>
#include <pagmo/problems/zdt.hpp>

#include <pagmo/algorithm.hpp>
#include <pagmo/algorithms/sade.hpp>
#include <pagmo/problems/wfg.hpp>

#include <pagmo_cuda/cuda_algorithms.hpp>
#include <pagmo_cuda/cuda_bfe.hpp>
#include <pagmo_cuda/cuda_hypervolume.hpp>
#include <pagmo/utils/hv_algos/hv_hvwfg.hpp>
#include <pagmo/utils/hypervolume.hpp>
#include <pagmo/utils/hv_algos/hv_hv2d.hpp>
#include <pagmo/utils/hv_algos/hv_hv3d.hpp>

#include "cec_synth.h"

namespace oracle_ref { void ensure_cec2014_tables(unsigned, unsigned); }

static int g_fail = 0;
#define CHECK(cond)                                                                                                    \
    do {                                                                                                               \
        if (!(cond)) {                                                                                                 \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);                                                \
            ++g_fail;                                                                                                  \
        }                                                                                                              \
    } while (0)

static double max_rel(const pagmo::vector_double &a, const pagmo::vector_double &b)
{
    double m = 0;
    if (a.size() != b.size()) return 1e300;
    for (std::size_t i = 0; i < a.size(); ++i) m = std::max(m, std::abs(a[i] - b[i]) / std::max(std::abs(b[i]), 1e-300));
    return m;
}

static pagmo::vector_double random_batch(const pagmo::problem &p, std::size_t n, unsigned seed)
{
    std::mt19937 e(seed);
    const auto lb = p.get_lb(), ub = p.get_ub();
    pagmo::vector_double dvs(n * p.get_nx());
    for (std::size_t i = 0; i < n; ++i)
        for (std::size_t j = 0; j < p.get_nx(); ++j)
            dvs[i * p.get_nx() + j] = std::uniform_real_distribution<double>(lb[j], ub[j])(e);
    return dvs;
}

// --bench N: the headline batch (N decision vectors, D = 100) through the COMPILED plugin call a pagmo user makes -
// pagmo::bfe{cuda_bfe{}}(problem, dvs) with a pageable std::vector<double> in and a fresh std::vector<double> out - for a
// sample of the 30 functions; prints one JSON line (bench.py: e2e.adapter).
static int bench_adapter(std::size_t n)
{
    using namespace pagmo_cuda;
    const unsigned dim = 100u;
    const unsigned funcs[] = {1u, 8u, 17u, 23u, 28u};
    std::vector<double> dvs(n * dim);
    {
        std::mt19937_64 e(20141);
        std::uniform_real_distribution<double> u(-100., 100.);
        for (auto &v : dvs) v = u(e);
    }
    pagmo::bfe b{cuda_bfe{}};
    double secs = 0., checksum = 0.;
    std::size_t evals = 0;
    for (unsigned func : funcs) {
        std::vector<double> mr(10u * dim * dim), lines(1000), shift;
        std::vector<int> shuf(10u * dim);
        cec2014_synth_rotation(func, dim, mr.data());
        cec2014_synth_shift(func, lines.data());
        cec2014_synth_shuffle(func, dim, shuf.data());
        for (unsigned i = 0; i < 1000u; ++i)
            if (i % 100u < dim) shift.push_back(lines[i]);
        pagmo::problem p{cuda_cec2014{func, dim, mr, shift, shuf}};
        {
            const std::vector<double> warm(dvs.begin(), dvs.begin() + 4096 * dim);
            (void)b(p, warm);
        }
        const auto t0 = std::chrono::steady_clock::now();
        const auto fvs = b(p, dvs);
        secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        evals += fvs.size();
        checksum += fvs[0] + fvs[fvs.size() - 1];
    }
    std::printf("{\"value\": %.6g, \"unit\": \"evals/s\", \"path\": \"pagmo::bfe{cuda_bfe{}}(problem{cuda_cec2014}, pageable std::vector<double>) -> "
                "fresh std::vector<double>, incl. pagmo's two size-check passes\", \"functions\": [1, 8, 17, 23, 28], \"individuals\": %zu, "
                "\"seconds\": %.4f, \"checksum\": %.17g}\n",
                static_cast<double>(evals) / secs, n, secs, checksum);
    return 0;
}

int main(int argc, char **argv)
{
    using namespace pagmo_cuda;
    const double tol = 1e-12;
    if (argc >= 3 && std::string(argv[1]) == "--bench") return bench_adapter(static_cast<std::size_t>(std::atoll(argv[2])));

    // ---- 1. cuda_bfe as a UDBFE on a STOCK pagmo UDP (pagmo::rastrigin) -------------------------------------
    {
        pagmo::problem p{pagmo::rastrigin{10u}};
        pagmo::bfe gpu{cuda_bfe{}}, cpu{pagmo::thread_bfe{}};
        CHECK(gpu.get_name().find("CUDA") != std::string::npos);
        CHECK(gpu.get_thread_safety() == pagmo::thread_safety::basic);
        const auto dvs = random_batch(p, 1024, 23);
        const auto f0 = p.get_fevals();
        const auto fg = gpu(p, dvs);
        CHECK(p.get_fevals() == f0 + 1024u); // bumped once, by pagmo::bfe (bfe.cpp:107)
        const auto fc = cpu(p, dvs);
        CHECK(max_rel(fg, fc) <= tol);
        // the wrapper's own input validation still applies (bfe_impl.cpp:72-77)
        bool threw = false;
        try {
            gpu(p, pagmo::vector_double(15));
        } catch (const std::invalid_argument &) {
            threw = true;
        }
        CHECK(threw);
        // empty batch
        CHECK(gpu(p, pagmo::vector_double{}).empty());
        // copies share the device state and stay usable
        pagmo::bfe gpu2(gpu);
        CHECK(max_rel(gpu2(p, dvs), fc) <= tol);
    }

    // ---- 2. a UDP without a device evaluator: throw, never fall back to the CPU -------------------------------
    {
        oracle_ref::ensure_cec2014_tables(3u, 10u);
        pagmo::problem p{pagmo::cec2014{3u, 10u}}; // the stock UDP keeps its data tables private: no device twin
        pagmo::bfe gpu{cuda_bfe{}};
        bool threw = false;
        try {
            gpu(p, random_batch(p, 4, 1));
        } catch (const std::invalid_argument &e) {
            threw = std::string(e.what()).find("no CPU fallback") != std::string::npos;
        }
        CHECK(threw);
    }

    // ---- 3. CUDA-backed UDP: fitness / batch_fitness / default_bfe / member_bfe / population ------------------
    for (unsigned func : {5u, 17u, 23u, 30u}) {
        const unsigned dim = 30u;
        std::vector<double> mr(10u * dim * dim), lines(1000), shift;
        std::vector<int> shuf(10u * dim);
        cec2014_synth_rotation(func, dim, mr.data());
        cec2014_synth_shift(func, lines.data());
        cec2014_synth_shuffle(func, dim, shuf.data());
        for (unsigned i = 0; i < 1000u; ++i)
            if (i % 100u < dim) shift.push_back(lines[i]); // cec2014.cpp:76-86
        oracle_ref::ensure_cec2014_tables(func, dim);
        pagmo::problem ref{pagmo::cec2014{func, dim}};
        pagmo::problem gpu{cuda_cec2014{func, dim, mr, shift, shuf}};
        CHECK(gpu.has_batch_fitness());
        CHECK(gpu.get_nx() == dim && gpu.get_nobj() == 1u);
        CHECK(gpu.get_bounds() == ref.get_bounds());
        CHECK(gpu.get_name().find("CEC2014 - f" + std::to_string(func)) != std::string::npos);
        CHECK(gpu.extract<cuda_cec2014>()->get_origin_shift() == ref.extract<pagmo::cec2014>()->get_origin_shift());
        const auto dvs = random_batch(ref, 777, 100 + func);
        const auto want = pagmo::bfe{pagmo::thread_bfe{}}(ref, dvs);
        // single fitness
        const pagmo::vector_double x0(dvs.begin(), dvs.begin() + dim);
        CHECK(std::abs(gpu.fitness(x0)[0] - want[0]) <= tol * std::abs(want[0]));
        // problem::batch_fitness (problem.cpp:413-427) counts fevals itself
        const auto f0 = gpu.get_fevals();
        CHECK(max_rel(gpu.batch_fitness(dvs), want) <= tol);
        CHECK(gpu.get_fevals() == f0 + 777u);
        // default_bfe prefers the member (default_bfe.cpp:56-57); member_bfe; cuda_bfe: all the same numbers
        CHECK(max_rel(pagmo::bfe{}(gpu, dvs), want) <= tol);
        CHECK(max_rel(pagmo::bfe{pagmo::member_bfe{}}(gpu, dvs), want) <= tol);
        CHECK(max_rel(pagmo::bfe{cuda_bfe{}}(gpu, dvs), want) <= tol);
        CHECK(gpu.get_fevals() == f0 + 4u * 777u);
        // population constructed through the bfe (population.cpp:82-103): same seed => same decision vectors as the
        // reference population, fitness within tolerance
        pagmo::population pg{gpu, pagmo::bfe{cuda_bfe{}}, 64u, 42u}, pr{ref, pagmo::bfe{pagmo::thread_bfe{}}, 64u, 42u};
        CHECK(pg.get_x() == pr.get_x());
        double worst = 0;
        for (std::size_t i = 0; i < 64u; ++i) worst = std::max(worst, max_rel(pg.get_f()[i], pr.get_f()[i]));
        CHECK(worst <= tol);
        CHECK(pg.best_idx() == pr.best_idx());
        std::printf("cec2014 f%u D=%u via pagmo::problem/bfe/population: ok (worst rel %.2e)\n", func, dim, worst);
    }

    // ---- 3b. multi-objective: stock zdt/dtlz through cuda_bfe, CUDA UDPs, and nsga2 with set_bfe ----------------
    {
        pagmo::bfe gpu{cuda_bfe{}}, cpu{pagmo::thread_bfe{}};
        for (unsigned id = 1; id <= 6u; ++id) {
            pagmo::problem ref{pagmo::zdt{id, 13u}}, twin{cuda_zdt{id, 13u}};
            const auto dvs = random_batch(ref, 300, 7 + id);
            const auto want = cpu(ref, dvs);
            CHECK(max_rel(gpu(ref, dvs), want) <= tol);       // stock UDP recognised by name
            CHECK(max_rel(pagmo::bfe{}(twin, dvs), want) <= tol); // CUDA UDP via default_bfe
            CHECK(twin.get_nobj() == 2u && twin.get_nix() == ref.get_nix() && twin.get_bounds() == ref.get_bounds());
        }
        for (unsigned id = 1; id <= 7u; ++id) {
            pagmo::problem ref{pagmo::dtlz{id, 12u, 3u, 100u}}, twin{cuda_dtlz{id, 12u, 3u, 100u}};
            const auto dvs = random_batch(ref, 300, 70 + id);
            const auto want = cpu(ref, dvs);
            if (id != 4u) CHECK(max_rel(gpu(ref, dvs), want) <= tol);
            CHECK(max_rel(gpu(twin, dvs), want) <= tol);
            CHECK(twin.get_nobj() == 3u);
        }
        // the reference's own bfe-equivalence pattern (tests/nsga2.cpp:204-232): same seed, with and without a bfe
        pagmo::problem prob{pagmo::zdt{1u, 30u}};
        pagmo::population pop1{prob, 40u, 23u}, pop2{prob, 40u, 23u};
        pagmo::nsga2 a1{10u, 0.95, 10., 0.01, 50., 32u}, a2{10u, 0.95, 10., 0.01, 50., 32u};
        a2.set_bfe(gpu);
        pop1 = a1.evolve(pop1);
        pop2 = a2.evolve(pop2);
        double worst = 0;
        for (std::size_t i = 0; i < 40u; ++i) worst = std::max(worst, max_rel(pop2.get_f()[i], pop1.get_f()[i]));
        CHECK(pop1.get_x() == pop2.get_x()); // identical trajectories: fitness differences stay below every comparison
        CHECK(worst <= tol);
        std::printf("nsga2 on zdt1 with set_bfe(cuda_bfe): same final population as the sequential run (rel %.2e)\n", worst);
    }

    // ---- 3b. wfg / lennard_jones UDPs and the stock lennard_jones twin ----
    {
        pagmo::bfe gpu{cuda_bfe{}}, cpu{pagmo::thread_bfe{}};
        for (unsigned id = 1; id <= 9u; ++id) {
            pagmo::problem ref{pagmo::wfg{id, 12u, 3u, 4u}}, twin{cuda_wfg{id, 12u, 3u, 4u}};
            const auto dvs = random_batch(ref, 200, 170 + id);
            CHECK(max_rel(gpu(twin, dvs), cpu(ref, dvs)) <= 1e-11);
            CHECK(twin.get_nobj() == 3u && twin.get_bounds() == ref.get_bounds());
        }
        pagmo::problem lj{pagmo::lennard_jones{7u}}, ljc{cuda_lennard_jones{7u}};
        const auto dvs = random_batch(lj, 100, 5);
        CHECK(max_rel(gpu(lj, dvs), cpu(lj, dvs)) <= 1e-10);
        CHECK(max_rel(gpu(ljc, dvs), cpu(lj, dvs)) <= 1e-10);
    }

    // ---- 3b'. meta-problems: cuda_translate / cuda_decompose against the stock pagmo::translate / pagmo::decompose ------------
    {
        pagmo::bfe gpu{cuda_bfe{}}, cpu{pagmo::thread_bfe{}};
        const pagmo::vector_double t{0.1, -0.2, 0.3, 0.4, -0.5};
        pagmo::problem ref{pagmo::translate{pagmo::rastrigin{5u}, t}}, twin{cuda_translate{cuda_simple<PGC_RASTRIGIN>{5u}, t}};
        CHECK(twin.get_bounds() == ref.get_bounds());
        CHECK(twin.get_name().find("[translated]") != std::string::npos);
        const auto dvs = random_batch(ref, 300, 77);
        CHECK(max_rel(gpu(twin, dvs), cpu(ref, dvs)) <= tol);
        CHECK(max_rel(twin.batch_fitness(dvs), cpu(ref, dvs)) <= tol);
        // the stock meta-problem over a CUDA UDP also runs (host de-shift + device batch_fitness, translate.cpp:118-153)
        pagmo::problem mixed{pagmo::translate{cuda_simple<PGC_RASTRIGIN>{5u}, t}};
        CHECK(max_rel(mixed.batch_fitness(dvs), cpu(ref, dvs)) <= tol);
        bool threw = false;
        try {
            cuda_translate{cuda_simple<PGC_RASTRIGIN>{5u}, {1., 2.}};
        } catch (const std::invalid_argument &) {
            threw = true; // translate.cpp:83-87
        }
        CHECK(threw);
        for (const char *method : {"weighted", "tchebycheff", "bi"}) {
            const pagmo::vector_double w{0.3, 0.7}, z{0.1, -0.2};
            pagmo::problem dref{pagmo::decompose{pagmo::zdt{1u, 30u}, w, z, method, false}};
            pagmo::problem dtwin{cuda_decompose{cuda_zdt{1u, 30u}, w, z, method, false}};
            CHECK(dtwin.get_nobj() == 1u && dtwin.get_bounds() == dref.get_bounds());
            const auto xs = random_batch(dref, 200, 78);
            CHECK(max_rel(gpu(dtwin, xs), cpu(dref, xs)) <= tol);
        }
        threw = false;
        try {
            cuda_decompose{cuda_zdt{1u, 30u}, {0.5, 0.6}, {0., 0.}};
        } catch (const std::invalid_argument &) {
            threw = true; // decompose.cpp:109-115
        }
        CHECK(threw);
        threw = false;
        try {
            cuda_decompose{cuda_simple<PGC_RASTRIGIN>{5u}, {0.5, 0.5}, {0., 0.}};
        } catch (const std::invalid_argument &) {
            threw = true; // decompose.cpp:70-72
        }
        CHECK(threw);
    }

    // ---- 3b''. constrained UDPs and cuda_unconstrain against the stock hock_schittkowski_71 / luksan_vlcek1 / pagmo::unconstrain -------
    {
        pagmo::bfe gpu{cuda_bfe{}}, cpu{pagmo::thread_bfe{}};
        pagmo::problem hs{pagmo::hock_schittkowski_71{}}, hsc{cuda_hock_schittkowski_71{}};
        CHECK(hsc.get_nobj() == 1u && hsc.get_nec() == 1u && hsc.get_nic() == 1u && hsc.get_nf() == 3u);
        CHECK(hsc.get_bounds() == hs.get_bounds());
        const auto dvs = random_batch(hs, 500, 91);
        CHECK(gpu(hsc, dvs) == cpu(hs, dvs)); // products and sums only: bit for bit
        CHECK(gpu(hs, dvs) == cpu(hs, dvs));  // the stock UDP finds its device twin
        CHECK(hsc.feasibility_x(pagmo::vector_double{1., 4.74299963, 3.82114998, 1.37940829}) == false); // tolerances are zero
        hsc.set_c_tol(1e-6);
        CHECK(hsc.feasibility_x(pagmo::vector_double{1., 4.74299963, 3.82114998, 1.37940829}) == true);
        pagmo::problem lv{pagmo::luksan_vlcek1{12u}}, lvc{cuda_luksan_vlcek1{12u}};
        CHECK(lvc.get_nec() == 10u && lvc.get_nic() == 0u && lvc.get_bounds() == lv.get_bounds());
        const auto xs = random_batch(lv, 300, 92);
        const auto a = gpu(lvc, xs), b = cpu(lv, xs), c = gpu(lv, xs);
        double worst = 0.;
        for (std::size_t i = 0; i < a.size(); ++i) worst = std::max({worst, std::abs(a[i] - b[i]), std::abs(c[i] - b[i])});
        CHECK(worst <= 1e-12 * 2e5); // terms up to 5 e^10 of mixed sign: 1e-12 of their magnitude
        const pagmo::vector_double w{3., 0.5}, ctol{2., 40.};
        for (const char *method : {"death penalty", "kuri", "weighted", "ignore_c", "ignore_o"}) {
            const bool wt = std::string(method) == "weighted";
            pagmo::problem inner{pagmo::hock_schittkowski_71{}};
            inner.set_c_tol(ctol);
            pagmo::problem uref{pagmo::unconstrain{inner, method, wt ? w : pagmo::vector_double{}}};
            pagmo::problem utwin{cuda_unconstrain{cuda_hock_schittkowski_71{}, method, wt ? w : pagmo::vector_double{}, ctol}};
            CHECK(utwin.get_nobj() == 1u && utwin.get_nc() == 0u && utwin.get_bounds() == uref.get_bounds());
            CHECK(utwin.get_name().find("[unconstrained]") != std::string::npos);
            CHECK(gpu(utwin, dvs) == cpu(uref, dvs));
            CHECK(utwin.batch_fitness(dvs) == cpu(uref, dvs));
            // the stock meta-problem over the CUDA UDP: device batch_fitness, host penalty (unconstrain.cpp:244-263)
            pagmo::problem hsc2{cuda_hock_schittkowski_71{}};
            hsc2.set_c_tol(ctol);
            pagmo::problem mixed{pagmo::unconstrain{hsc2, method, wt ? w : pagmo::vector_double{}}};
            CHECK(mixed.batch_fitness(dvs) == cpu(uref, dvs));
        }
        const auto throws = [](auto make) {
            try {
                make();
            } catch (const std::invalid_argument &) {
                return true;
            }
            return false;
        };
        CHECK(throws([] { cuda_unconstrain{cuda_simple<PGC_RASTRIGIN>{5u}}; }));                             // unconstrain.cpp:73-77
        CHECK(throws([] { cuda_unconstrain{cuda_hock_schittkowski_71{}, "weighted", {1.}}; }));               // :79-82
        CHECK(throws([] { cuda_unconstrain{cuda_hock_schittkowski_71{}, "mispelled"}; }));                    // :84-88
        CHECK(throws([] { cuda_unconstrain{cuda_hock_schittkowski_71{}, "kuri", {1., 1.}}; }));               // :90-93
        CHECK(throws([] { cuda_luksan_vlcek1{2u}; }));                                                        // luksan_vlcek1.cpp:46-49
        { // the device algorithms refuse constraints like the reference's do, and run on the unconstrained problem
            using namespace pagmo_cuda;
            pagmo::population pc{hsc, 32u, 5u};
            CHECK(throws([&] { pagmo::algorithm{cuda_de{5u}}.evolve(pc); }));
            pagmo::problem up{cuda_unconstrain{cuda_hock_schittkowski_71{}, "weighted", {10., 10.}, {1e-3, 1e-3}}};
            pagmo::population pu{up, 64u, 5u};
            const double before = pu.champion_f()[0];
            pu = pagmo::algorithm{cuda_de{60u, 0.8, 0.9, 2u, 0., 0., 3u}}.evolve(pu);
            CHECK(pu.champion_f()[0] < before);
            std::printf("cuda_de on unconstrain{hs71, weighted}: %.6g -> %.6g (optimum 17.014)\n", before, pu.champion_f()[0]);
        }
    }

    // ---- 3c. CUDA UDAs behind pagmo::algorithm: evolve() keeps the population consistent and counts fevals like the reference ----
    {
        using namespace pagmo_cuda;
        pagmo::problem prob{pagmo::rastrigin{10u}};
        pagmo::population pop{prob, 64u, 23u};
        const double before = pop.champion_f()[0];
        const auto fe0 = pop.get_problem().get_fevals();
        pagmo::algorithm algo{cuda_sade{30u, 2u, 1u, 0., 0., false, 41u}};
        pop = algo.evolve(pop);
        CHECK(pop.get_problem().get_fevals() - fe0 == 30u * 64u);
        CHECK(pop.champion_f()[0] < before);
        for (std::size_t i = 0; i < pop.size(); ++i) CHECK(max_rel(prob.fitness(pop.get_x()[i]), pop.get_f()[i]) <= tol);
        pagmo::population pop2{prob, 64u, 23u};
        pop2 = pagmo::algorithm{cuda_sade{30u, 2u, 1u, 0., 0., false, 41u}}.evolve(pop2);
        CHECK(pop2.get_x() == pop.get_x()); // deterministic given the seeds
        for (const auto &a : {pagmo::algorithm{cuda_de{20u, 0.8, 0.9, 2u, 0., 0., 3u}}, pagmo::algorithm{cuda_de1220{20u, {2u, 3u, 7u}, 1u, 0., 0., false, 3u}},
                              pagmo::algorithm{cuda_pso_gen{20u, 0.7298, 2.05, 2.05, 0.5, 5u, 2u, 4u, false, 3u}},
                              pagmo::algorithm{cuda_sga{20u, .9, 1., 0.02, 1., 2u, "sbx", "gaussian", "truncated", 3u}},
                              pagmo::algorithm{cuda_pso{20u, 0.7298, 2.05, 2.05, 0.5, 5u, 2u, 4u, false, 3u}},
                              pagmo::algorithm{cuda_xnes{20u, -1, -1, -1, -1, 0., 0., false, true, 3u}}}) {
            pagmo::population q{prob, 32u, 9u};
            const double b = q.champion_f()[0];
            q = a.evolve(q);
            CHECK(q.champion_f()[0] <= b);
            std::printf("%s: %.4g -> %.4g\n", a.get_name().c_str(), b, q.champion_f()[0]);
        }
        { // memory = true (sade.cpp:137-156, de1220.cpp:147-165, pso_gen.cpp:193-201): the adaptation state / velocities survive
          // between evolve() calls, so 3 + 3 generations of jDE are the uninterrupted 6; without memory the second call re-draws them
            const auto split_equals_whole = [&](auto make) {
                pagmo::population whole{prob, 32u, 9u}, split{prob, 32u, 9u};
                whole = pagmo::algorithm{make(6u)}.evolve(whole);
                pagmo::algorithm a3{make(3u)};
                split = a3.evolve(split); // pagmo::algorithm::evolve is const: the state lives in the UDA's mutable members
                split = a3.evolve(split);
                return whole.get_x() == split.get_x() && whole.get_f() == split.get_f();
            };
            CHECK(split_equals_whole([](unsigned g) { return cuda_sade{g, 2u, 1u, 0., 0., true, 41u}; }));
            CHECK(!split_equals_whole([](unsigned g) { return cuda_sade{g, 2u, 1u, 0., 0., false, 41u}; }));
            CHECK(split_equals_whole([](unsigned g) { return cuda_de1220{g, {2u, 3u, 7u, 10u}, 1u, 0., 0., true, 5u}; }));
            CHECK(split_equals_whole([](unsigned g) { return cuda_cmaes{g, -1, -1, -1, -1, 0.5, 0., 0., true, false, 5u}; }));
            CHECK(!split_equals_whole([](unsigned g) { return cuda_cmaes{g, -1, -1, -1, -1, 0.5, 0., 0., false, false, 5u}; }));
            CHECK(split_equals_whole([](unsigned g) { return cuda_xnes{g, -1, -1, -1, -1, 0., 0., true, false, 5u}; }));
            CHECK(!split_equals_whole([](unsigned g) { return cuda_de1220{g, {2u, 3u, 7u, 10u}, 1u, 0., 0., false, 5u}; }));
            { // pso_gen with memory restarts from the particles' best positions with the KEPT velocities (pso_gen.cpp:193-201)
                pagmo::population a{prob, 32u, 9u}, b{prob, 32u, 9u};
                pagmo::algorithm keep{cuda_pso_gen{3u, 0.7298, 2.05, 2.05, 0.5, 5u, 2u, 4u, true, 3u}};
                pagmo::algorithm redraw{cuda_pso_gen{3u, 0.7298, 2.05, 2.05, 0.5, 5u, 2u, 4u, false, 3u}};
                a = keep.evolve(a), b = redraw.evolve(b);
                CHECK(a.get_x() == b.get_x()); // the first call draws the same velocities either way
                a = keep.evolve(a), b = redraw.evolve(b);
                CHECK(a.get_x() != b.get_x());
            }
            // a copy of the algorithm carries the state (pagmo copies the UDA in and out of an island around every evolve)
            pagmo::population whole{prob, 32u, 9u}, split{prob, 32u, 9u};
            whole = pagmo::algorithm{cuda_sade{6u, 2u, 1u, 0., 0., true, 41u}}.evolve(whole);
            pagmo::algorithm first{cuda_sade{3u, 2u, 1u, 0., 0., true, 41u}};
            split = first.evolve(split);
            pagmo::algorithm second = first;
            split = second.evolve(split);
            CHECK(whole.get_x() == split.get_x());
            // a population of another size restarts the adaptation instead of reading a stale state
            pagmo::population other{prob, 48u, 2u};
            other = second.evolve(other);
            CHECK(other.size() == 48u);
        }
        { // algorithm::set_verbosity + get_log (de.hpp:160-199): one line every `level` generations, in the reference's tuple layout
            pagmo::algorithm a{cuda_de{10u, 0.8, 0.9, 2u, 0., 0., 3u}};
            a.set_verbosity(3u);
            pagmo::population q{prob, 32u, 9u}, plain{prob, 32u, 9u};
            q = a.evolve(q);
            plain = pagmo::algorithm{cuda_de{10u, 0.8, 0.9, 2u, 0., 0., 3u}}.evolve(plain);
            CHECK(q.get_x() == plain.get_x()); // logging does not change the run
            const auto log = a.extract<cuda_de>()->get_log();
            CHECK(log.size() == 4u && std::get<0>(log[0]) == 1u && std::get<0>(log[3]) == 10u && std::get<1>(log[3]) == 10u * 32u);
            CHECK(std::get<2>(log[3]) == q.champion_f()[0]); // the last line is the final generation here: Best = the champion
            CHECK(std::get<2>(log[0]) >= std::get<2>(log[3]) && std::get<3>(log[3]) > 0. && std::get<4>(log[3]) > 0.);
            pagmo::algorithm s{cuda_sade{6u, 2u, 1u, 0., 0., false, 41u}};
            s.set_verbosity(1u);
            pagmo::population qs{prob, 32u, 9u};
            qs = s.evolve(qs);
            const auto slog = s.extract<cuda_sade>()->get_log();
            CHECK(slog.size() == 6u && std::get<3>(slog[5]) >= 0.1 && std::get<3>(slog[5]) <= 1. && std::get<4>(slog[5]) >= 0. && std::get<4>(slog[5]) <= 1.);
            pagmo::algorithm p{cuda_pso_gen{5u, 0.7298, 2.05, 2.05, 0.5, 5u, 2u, 4u, false, 3u}};
            p.set_verbosity(2u);
            pagmo::population qp{prob, 32u, 9u};
            qp = p.evolve(qp);
            const auto plog = p.extract<cuda_pso_gen>()->get_log();
            CHECK(plog.size() == 3u && std::get<2>(plog[2]) == qp.champion_f()[0] && std::get<5>(plog[2]) > 0.);
            pagmo::algorithm m{cuda_nsga2{5u, 0.95, 10., 0.01, 50., 32u}};
            m.set_verbosity(2u);
            pagmo::population qm{pagmo::problem{pagmo::zdt{1u, 30u}}, 40u, 5u};
            const auto ideal0 = pagmo::ideal(qm.get_f());
            qm = m.evolve(qm);
            const auto mlog = m.extract<cuda_nsga2>()->get_log();
            CHECK(mlog.size() == 3u && std::get<1>(mlog[0]) == 0u && std::get<2>(mlog[0]) == ideal0 && std::get<1>(mlog[2]) == 4u * 40u);
            pagmo::algorithm g{cuda_sga{9u, .9, 1., 0.02, 1., 2u, "sbx", "gaussian", "truncated", 3u}};
            g.set_verbosity(4u);
            pagmo::population qg{prob, 32u, 9u};
            const double parents_best = qg.champion_f()[0];
            qg = g.evolve(qg);
            const auto glog = g.extract<cuda_sga>()->get_log(); // generations 1, 5, 9
            CHECK(glog.size() == 3u && std::get<0>(glog[1]) == 5u && std::get<2>(glog[0]) == parents_best);
            pagmo::algorithm c{cuda_cmaes{6u, -1, -1, -1, -1, 0.5, 0., 0., false, false, 3u}};
            c.set_verbosity(5u);
            pagmo::population qc{prob, 16u, 9u};
            qc = c.evolve(qc);
            const auto clog = c.extract<cuda_cmaes>()->get_log(); // generations 1, 6
            CHECK(clog.size() == 2u && std::get<5>(clog[0]) == 0.5 && std::get<1>(clog[1]) == 5u * 16u);
        }
        { // cuda_gaco: the population comes back consistent, fevals as the reference counts them, the oracle parameter is a member
            pagmo::algorithm g{cuda_gaco{15u, 13u, 1.0, 1e9, 0.01, 1u, 7u, 100000u, 100000u, 0., false, 5u}};
            pagmo::population q{prob, 40u, 9u};
            const double b = q.champion_f()[0];
            const auto fe = q.get_problem().get_fevals();
            q = g.evolve(q);
            CHECK(q.get_problem().get_fevals() - fe == 15u * 40u);
            CHECK(q.champion_f()[0] <= b);
            for (std::size_t i = 0; i < q.size(); ++i) CHECK(max_rel(prob.fitness(q.get_x()[i]), q.get_f()[i]) <= tol);
            CHECK(g.extract<cuda_gaco>()->get_oracle() < 1e9);
            {
                pagmo::algorithm gl{cuda_gaco{7u, 13u, 1.0, 1e9, 0.01, 1u, 7u, 100000u, 100000u, 0., false, 5u}};
                gl.set_verbosity(3u);
                pagmo::population ql{prob, 40u, 9u};
                const double b0 = ql.champion_f()[0];
                ql = gl.evolve(ql);
                const auto log = gl.extract<cuda_gaco>()->get_log(); // generations 1, 4 inside the loop, 7 after it (gaco.cpp:254-287, :405-445)
                CHECK(log.size() == 3u && std::get<0>(log[2]) == 7u && std::get<1>(log[2]) == 7u * 40u && std::get<2>(log[0]) == b0);
                CHECK(std::get<3>(log[1]) == 13u && std::get<2>(log[2]) == ql.champion_f()[0]);
            }
            std::printf("%s: %.4g -> %.4g (oracle %.4g)\n", g.get_name().c_str(), b, q.champion_f()[0], g.extract<cuda_gaco>()->get_oracle());
            { // memory = true: the archive lives in the algorithm object, copies carry it, and the population keeps the last ants
                pagmo::algorithm gm{cuda_gaco{1u, 8u, 1.0, 1e9, 0.01, 3u, 7u, 100000u, 100000u, 0., true, 5u}};
                pagmo::population qa{prob, 30u, 9u}, qb{prob, 30u, 9u};
                for (int c = 0; c < 5; ++c) qa = gm.evolve(qa);
                pagmo::algorithm g1{cuda_gaco{1u, 8u, 1.0, 1e9, 0.01, 3u, 7u, 100000u, 100000u, 0., true, 5u}};
                for (int c = 0; c < 3; ++c) qb = g1.evolve(qb);
                pagmo::algorithm g2 = g1; // a copy continues where the original stood
                for (int c = 0; c < 2; ++c) qb = g2.evolve(qb);
                CHECK(qa.get_x() == qb.get_x() && qa.get_f() == qb.get_f());
                CHECK(gm.extract<cuda_gaco>()->get_oracle() < 1e9);
            }
            bool threw = false;
            try {
                cuda_gaco{5u, 1u};
            } catch (const std::invalid_argument &) {
                threw = true; // gaco.cpp:90-93
            }
            CHECK(threw);
        }
        { // cuda_maco: fitness consistent with the decision vectors, fevals as the reference counts them, deterministic given the seed
            pagmo::problem z1{pagmo::zdt{1u, 30u}};
            pagmo::population a{z1, 40u, 5u}, b{z1, 40u, 5u};
            const auto fe = a.get_problem().get_fevals();
            a = pagmo::algorithm{cuda_maco{8u, 12u, 1.0, 1u, 7u, 100000u, 0., false, 5u}}.evolve(a);
            b = pagmo::algorithm{cuda_maco{8u, 12u, 1.0, 1u, 7u, 100000u, 0., false, 5u}}.evolve(b);
            CHECK(a.get_problem().get_fevals() - fe == 8u * 40u);
            CHECK(a.get_x() == b.get_x() && a.get_f() == b.get_f());
            for (std::size_t i = 0; i < a.size(); ++i) CHECK(max_rel(z1.fitness(a.get_x()[i]), a.get_f()[i]) <= tol);
            pagmo::algorithm ml{cuda_maco{5u, 12u, 1.0, 1u, 7u, 100000u, 0., false, 5u}};
            ml.set_verbosity(2u);
            pagmo::population c{z1, 40u, 5u};
            const auto ideal0 = pagmo::ideal(c.get_f());
            c = ml.evolve(c);
            const auto mlog = ml.extract<cuda_maco>()->get_log();
            CHECK(mlog.size() == 3u && std::get<1>(mlog[1]) == 2u * 40u && std::get<2>(mlog[0]) == ideal0);
            pagmo::algorithm dl{cuda_moead_gen{4u, "grid", "tchebycheff", 5u, 1.0, 0.5, 20., 0.9, 2u, true, 5u}};
            dl.set_verbosity(2u);
            pagmo::population d{z1, 40u, 5u};
            d = dl.evolve(d);
            const auto dlog = dl.extract<cuda_moead_gen>()->get_log();
            CHECK(dlog.size() == 2u && std::get<0>(dlog[1]) == 3u && std::get<3>(dlog[0]) == ideal0 && std::get<2>(dlog[0]) > 0.);
        }
        pagmo::problem zp{pagmo::zdt{1u, 30u}};
        pagmo::population mo{zp, 40u, 5u};
        mo = pagmo::algorithm{cuda_nsga2{10u, 0.95, 10., 0.01, 50., 32u}}.evolve(mo);
        CHECK(mo.size() == 40u && mo.get_f()[0].size() == 2u);
        { // ZDT5 is all-integer: cuda_nsga2 applies the reference's integer operators, the decision vectors stay integral
            pagmo::problem z5{pagmo::zdt{5u, 11u}};
            pagmo::population ip{z5, 40u, 3u};
            ip = pagmo::algorithm{cuda_nsga2{8u, 0.95, 10., 0.05, 50., 11u}}.evolve(ip);
            bool integral = true;
            for (const auto &xv : ip.get_x())
                for (double g : xv) integral = integral && g == std::floor(g) && g >= 0. && g <= 1.;
            CHECK(integral && ip.get_problem().get_nix() == z5.get_nx());
            bool refused = false;
            try {
                pagmo::algorithm{cuda_de{2u}}.evolve(pagmo::population{pagmo::problem{pagmo::decompose{pagmo::zdt{5u, 11u}, {0.5, 0.5}, {0., 0.}}}, 16u, 1u});
            } catch (const std::invalid_argument &) {
                refused = true; // the other device UDAs refuse integer genes (or find no twin): never a silent continuous run
            }
            CHECK(refused);
        }
        // cuda_nspso: the three diversity mechanisms run, fitness stays consistent with the decision vectors, fevals counted like the
        // reference (NP per generation), the constructor rejects what nspso.cpp:58-80 rejects
        for (const char *div : {"crowding distance", "niche count", "max min"}) {
            pagmo::population sw{zp, 40u, 9u};
            const auto fe0 = sw.get_problem().get_fevals();
            sw = pagmo::algorithm{cuda_nspso{6u, 0.6, 2.0, 2.0, 1.0, 0.5, 60u, div, false, 17u}}.evolve(sw);
            CHECK(sw.get_problem().get_fevals() - fe0 == 6u * 40u);
            for (decltype(sw.size()) i = 0; i < sw.size(); ++i) {
                const auto fi = zp.fitness(sw.get_x()[i]);
                CHECK(std::abs(fi[0] - sw.get_f()[i][0]) <= 1e-12 * (1 + std::abs(fi[0])) && std::abs(fi[1] - sw.get_f()[i][1]) <= 1e-12 * (1 + std::abs(fi[1])));
            }
        }
        // cuda_moead_gen: pagmo's own weights / neighbourhoods, evolve on the device, and the same checks as the reference's
        for (const char *dec : {"tchebycheff", "weighted", "bi"}) {
            pagmo::population mp{zp, 40u, 9u};
            const auto fe0 = mp.get_problem().get_fevals();
            mp = pagmo::algorithm{cuda_moead_gen{8u, "grid", dec, 10u, 1.0, 0.5, 20., 0.9, 2u, true, 23u}}.evolve(mp);
            CHECK(mp.get_problem().get_fevals() - fe0 == 8u * 40u);
            for (decltype(mp.size()) i = 0; i < mp.size(); ++i) {
                const auto fi = zp.fitness(mp.get_x()[i]);
                CHECK(std::abs(fi[0] - mp.get_f()[i][0]) <= 1e-12 * (1 + std::abs(fi[0])) && std::abs(fi[1] - mp.get_f()[i][1]) <= 1e-12 * (1 + std::abs(fi[1])));
            }
        }
        {
            bool threw_T = false;
            try {
                pagmo::population small{zp, 8u, 1u};
                pagmo::algorithm{cuda_moead_gen{1u, "grid", "tchebycheff", 20u}}.evolve(small); // T > NP - 1, moead_gen.cpp:161-166
            } catch (const std::invalid_argument &) {
                threw_T = true;
            }
            CHECK(threw_T);
        }
        {
            bool bad_arg = false;
            try {
                cuda_nspso{1u, 0.6, 2.0, 2.0, 1.0, 0.5, 60u, "no such mechanism"};
            } catch (const std::invalid_argument &) {
                bad_arg = true;
            }
            CHECK(bad_arg);
        }
        bool threw = false;
        try {
            oracle_ref::ensure_cec2014_tables(1u, 10u);
            pagmo::population bad{pagmo::problem{pagmo::cec2014{1u, 10u}}, 16u, 1u};
            pagmo::algorithm{cuda_sade{2u}}.evolve(bad); // stock cec2014 keeps its tables private: no device twin
        } catch (const std::invalid_argument &) {
            threw = true;
        }
        CHECK(threw);
    }

    // ---- 3d. cuda_hv behind pagmo::hypervolume (reference tests/hypervolume.cpp pattern: same answers as hv2d / hv3d) ----
    {
        std::mt19937 e(77);
        std::uniform_real_distribution<double> u(0., 1.);
        for (unsigned m : {2u, 3u}) {
            std::vector<pagmo::vector_double> pts(200, pagmo::vector_double(m));
            for (auto &p : pts) {
                double nrm = 0;
                for (auto &v : p) {
                    v = u(e);
                    nrm += v * v;
                }
                for (auto &v : p) v /= std::sqrt(nrm);
            }
            const pagmo::vector_double r(m, 1.2);
            pagmo::hypervolume hv{pts, true};
            pagmo_cuda::cuda_hv gpu_algo;
            pagmo::hv2d a2;
            pagmo::hv3d a3;
            pagmo::hv_algorithm &cpu_algo = (m == 2u) ? static_cast<pagmo::hv_algorithm &>(a2) : static_cast<pagmo::hv_algorithm &>(a3);
            const double want = hv.compute(r, cpu_algo), got = hv.compute(r, gpu_algo);
            CHECK(std::abs(got - want) <= 1e-12 * want);
            const auto cw = hv.contributions(r, cpu_algo), cg = hv.contributions(r, gpu_algo);
            CHECK(max_rel(cg, cw) <= 1e-9);
            CHECK(hv.least_contributor(r, gpu_algo) == hv.least_contributor(r, cpu_algo));
            CHECK(hv.greatest_contributor(r, gpu_algo) == hv.greatest_contributor(r, cpu_algo));
            CHECK(std::abs(hv.exclusive(5u, r, gpu_algo) - cw[5]) <= 1e-9 * std::max(cw[5], 1e-300));
            std::printf("cuda_hv m=%u: hv %.15g (cpu %.15g)\n", m, got, want);
        }
    }

    // ---- 3d'. four objectives (hvwfg's territory) and the approximation algorithms behind pagmo::hypervolume ----
    {
        std::mt19937 e(78);
        std::uniform_real_distribution<double> u(0.05, 1.);
        std::vector<pagmo::vector_double> pts(60, pagmo::vector_double(4));
        for (auto &p : pts) {
            double nrm = 0;
            for (auto &v : p) {
                v = u(e);
                nrm += v * v;
            }
            for (auto &v : p) v /= std::sqrt(nrm);
        }
        const pagmo::vector_double r(4, 1.2);
        pagmo::hypervolume hv{pts, true};
        pagmo_cuda::cuda_hv gpu_algo;
        pagmo::hvwfg wfg;
        const double want = hv.compute(r, wfg), got = hv.compute(r, gpu_algo);
        CHECK(std::abs(got - want) <= 1e-12 * want);
        const auto cw = hv.contributions(r, wfg), cg = hv.contributions(r, gpu_algo);
        CHECK(max_rel(cg, cw) <= 1e-9);
        pagmo_cuda::cuda_bf_fpras fpras{0.02, 0.01, 5u};
        CHECK(std::abs(hv.compute(r, fpras) - want) <= 0.02 * want);
        pagmo_cuda::cuda_bf_approx approx{true, 1u, 0.05, 1e-4, 0.775, 0.2, 0.1, 0.25, 5u};
        const auto lo = hv.least_contributor(r, approx), hi = hv.greatest_contributor(r, approx);
        CHECK(cw[lo] <= 1.05 * *std::min_element(cw.begin(), cw.end()) && 1.05 * cw[hi] >= *std::max_element(cw.begin(), cw.end()));
        bool threw = false;
        try {
            hv.compute(r, approx);
        } catch (const std::invalid_argument &) {
            threw = true; // hv_bf_approx.cpp:69-73
        }
        CHECK(threw);
        std::printf("cuda_hv m=4: hv %.15g (hvwfg %.15g), fpras %.6g, least / greatest contributor %llu / %llu\n", got, want, hv.compute(r, fpras), lo, hi);
    }

    // ---- 3e. one batch sharded over several devices (here: every visible device, or device 0 twice): same values as one device ----
    {
        int ndev = 0;
        pgc_device_count(&ndev);
        std::vector<int> devs = ndev >= 2 ? std::vector<int>{0, 1} : std::vector<int>{0, 0};
        pagmo::bfe multi{cuda_bfe{devs}}, single{cuda_bfe{0}};
        const unsigned dim = 30u;
        std::vector<double> Mr(10u * dim * dim), Os(10u * 100u);
        std::vector<int> S(10u * dim);
        cec2014_synth_rotation(27u, dim, Mr.data());
        cec2014_synth_shift(27u, Os.data());
        cec2014_synth_shuffle(27u, dim, S.data());
        std::vector<double> Osc; // compaction of the ctor, cec2014.cpp:76-86
        for (unsigned k = 0; k < 10u; ++k) Osc.insert(Osc.end(), Os.begin() + 100u * k, Os.begin() + 100u * k + dim);
        pagmo::problem p{cuda_cec2014{27u, dim, Mr, Osc, S}};
        for (std::size_t n : {std::size_t(0), std::size_t(1), std::size_t(1001)}) {
            const auto dvs = random_batch(p, n, 11);
            CHECK(multi(p, dvs) == single(p, dvs));
        }
        pagmo::problem r{pagmo::rastrigin{9u}};
        const auto dvs = random_batch(r, 4097, 12);
        CHECK(multi(r, dvs) == single(r, dvs));
        std::printf("cuda_bfe over devices {%d,%d}: identical to one device\n", devs[0], devs[1]);
    }

    // ---- 3f. island threads: different problems on the same device, evaluated concurrently from their own threads ----
    {
        const unsigned dims[4] = {5u, 17u, 30u, 64u};
        std::vector<pagmo::problem> probs;
        std::vector<pagmo::vector_double> batches, wants;
        pagmo::bfe gpu{cuda_bfe{}};
        for (unsigned d : dims) {
            probs.emplace_back(cuda_rastrigin{d});
            batches.push_back(random_batch(probs.back(), 3000, d));
            wants.push_back(gpu(probs.back(), batches.back()));
        }
        std::vector<int> bad(4, 0);
        std::vector<std::thread> th;
        for (int t = 0; t < 4; ++t)
            th.emplace_back([&, t]() {
                pagmo::bfe mine{cuda_bfe{}};
                for (int rep = 0; rep < 25; ++rep)
                    if (mine(probs[t], batches[t]) != wants[t]) ++bad[t];
            });
        for (auto &x : th) x.join();
        CHECK(bad[0] + bad[1] + bad[2] + bad[3] == 0);
    }

    // ---- 4. constructor errors surface as std::invalid_argument, like the reference UDP (cec2014.cpp:51-64) ----
    {
        bool threw = false;
        try {
            cuda_cec2014 bad{29u, 2u, std::vector<double>(40), std::vector<double>(20), std::vector<int>(20, 1)};
        } catch (const std::invalid_argument &) {
            threw = true;
        }
        CHECK(threw);
        threw = false;
        try {
            cuda_rosenbrock bad{1u};
        } catch (const std::invalid_argument &) {
            threw = true;
        }
        CHECK(threw);
    }

    std::printf(g_fail ? "ADAPTERS FAILED (%d)\n" : "ADAPTERS OK\n", g_fail);
    return g_fail ? 1 : 0;
}
