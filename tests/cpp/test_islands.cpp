// tests/cpp/test_islands.cpp - GPU islands behind pagmo's OWN island / archipelago classes, and the device-resident archipelago.
// Compiled against the unmodified reference headers (+ oracle/shim) and linked with oracle/_ref/libpagmo_ref.so - which contains
// the reference's island.cpp, archipelago.cpp, topology.cpp, ring.cpp, thread_island.cpp, task_queue.cpp unmodified - and
// pagmo2_b200/libpgc.so.  Mirrors reference tests/archipelago.cpp (evolve / wait_check / champions / migration log) and
// tests/thread_island.cpp.  Run on the GPU box by tests/test_adapters.py; `--devices N` also exercises the NCCL path.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <random>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include <pagmo/algorithm.hpp>
#include <pagmo/algorithms/de.hpp>
#include <pagmo/archipelago.hpp>
#include <pagmo/island.hpp>
#include <pagmo/islands/thread_island.hpp>
#include <pagmo/population.hpp>
#include <pagmo/problem.hpp>
#include <pagmo/problems/rastrigin.hpp>
#include <pagmo/r_policies/fair_replace.hpp>
#include <pagmo/s_policies/select_best.hpp>
#include <pagmo/topologies/fully_connected.hpp>
#include <pagmo/topologies/ring.hpp>
#include <pagmo/topologies/unconnected.hpp>

#include <pagmo_cuda/cuda_algorithms.hpp>
#include <pagmo_cuda/cuda_bfe.hpp>
#include <pagmo_cuda/cuda_island.hpp>

static int g_fail = 0;
#define CHECK(cond)                                                                                                    \
    do {                                                                                                               \
        if (!(cond)) {                                                                                                 \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);                                                \
            ++g_fail;                                                                                                  \
        }                                                                                                              \
    } while (0)

// ---- a minimal in-memory archive with Boost.Serialization's `ar & x` interface, to round-trip the adapters' serialize() ----
struct mem_oarchive {
    using is_loading = std::false_type; // as Boost's archives declare themselves
    using is_saving = std::true_type;
    std::vector<unsigned char> buf;
    template <typename T, typename std::enable_if<std::is_arithmetic<T>::value || std::is_enum<T>::value, int>::type = 0>
    mem_oarchive &operator&(const T &v)
    {
        const auto *p = reinterpret_cast<const unsigned char *>(&v);
        buf.insert(buf.end(), p, p + sizeof(T));
        return *this;
    }
    mem_oarchive &operator&(const std::string &s)
    {
        *this & s.size();
        buf.insert(buf.end(), s.begin(), s.end());
        return *this;
    }
    template <typename T>
    mem_oarchive &operator&(const std::vector<T> &v)
    {
        *this & v.size();
        for (const auto &e : v) *this & e;
        return *this;
    }
};
struct mem_iarchive {
    using is_loading = std::true_type;
    using is_saving = std::false_type;
    const std::vector<unsigned char> &buf;
    std::size_t pos = 0;
    template <typename T, typename std::enable_if<std::is_arithmetic<T>::value || std::is_enum<T>::value, int>::type = 0>
    mem_iarchive &operator&(T &v)
    {
        std::memcpy(&v, buf.data() + pos, sizeof(T));
        pos += sizeof(T);
        return *this;
    }
    mem_iarchive &operator&(std::string &s)
    {
        std::size_t n = 0;
        *this & n;
        s.assign(reinterpret_cast<const char *>(buf.data() + pos), n);
        pos += n;
        return *this;
    }
    template <typename T>
    mem_iarchive &operator&(std::vector<T> &v)
    {
        std::size_t n = 0;
        *this & n;
        v.resize(n);
        for (auto &e : v) *this & e;
        return *this;
    }
};

template <typename Uda>
static Uda round_trip(const Uda &a)
{
    mem_oarchive oa;
    const_cast<Uda &>(a).serialize(oa, 0u);
    Uda b; // pagmo default-constructs a UDA before loading into it
    mem_iarchive ia{oa.buf};
    b.serialize(ia, 0u);
    return b;
}

static bool same_population(const pagmo::population &a, const pagmo::population &b)
{
    return a.get_x() == b.get_x() && a.get_f() == b.get_f() && a.get_ID() == b.get_ID();
}

int main(int argc, char **argv)
{
    using namespace pagmo_cuda;
    int ndev = 1;
    for (int i = 1; i + 1 < argc; ++i)
        if (std::string(argv[i]) == "--devices") ndev = std::atoi(argv[i + 1]);

    // ---- 0. serialisation of the CUDA UDPs: a load into a default-constructed UDP re-creates the same device problem ----
    {
        const auto same_problem = [](const auto &a, const auto &b, unsigned seed) {
            pagmo::problem pa{a}, pb{b};
            if (pa.get_nx() != pb.get_nx() || pa.get_nf() != pb.get_nf() || pa.get_nec() != pb.get_nec() || pa.get_bounds() != pb.get_bounds()
                || pa.get_name() != pb.get_name())
                return false;
            const auto lb = pa.get_lb(), ub = pa.get_ub();
            std::mt19937 e(seed);
            pagmo::vector_double xs(16u * pa.get_nx());
            for (std::size_t i = 0; i < xs.size(); ++i) xs[i] = std::uniform_real_distribution<double>(lb[i % lb.size()], ub[i % lb.size()])(e);
            return pa.batch_fitness(xs) == pb.batch_fitness(xs);
        };
        cuda_rastrigin r(7u);
        CHECK(same_problem(r, round_trip(r), 1u));
        cuda_zdt z(3u, 17u);
        CHECK(same_problem(z, round_trip(z), 2u));
        cuda_dtlz dz(2u, 9u, 3u, 100u);
        CHECK(same_problem(dz, round_trip(dz), 3u));
        cuda_lennard_jones lj(9u);
        CHECK(same_problem(lj, round_trip(lj), 4u));
        cuda_luksan_vlcek1 lv(8u);
        CHECK(same_problem(lv, round_trip(lv), 5u));
        cuda_hock_schittkowski_71 hs;
        CHECK(same_problem(hs, round_trip(hs), 6u));
        { // cec2014 with its tables, and the meta-problems with their chains: translate{cec2014}, unconstrain{translate{hs71}}, decompose{zdt}
            const unsigned D = 10u;
            std::mt19937 e(11u);
            std::vector<double> mr(static_cast<std::size_t>(D) * D, 0.), os(D);
            for (unsigned i = 0; i < D; ++i) mr[static_cast<std::size_t>(i) * D + (i + 3u) % D] = (i % 2u) ? -1. : 1.; // a signed permutation: orthogonal
            for (auto &v : os) v = std::uniform_real_distribution<double>(-80., 80.)(e);
            cuda_cec2014 c(1u, D, mr, os);
            CHECK(same_problem(c, round_trip(c), 7u));
            pagmo::vector_double t(D);
            for (auto &v : t) v = std::uniform_real_distribution<double>(-1., 1.)(e);
            cuda_translate tc(c, t);
            const auto tc2 = round_trip(tc);
            CHECK(same_problem(tc, tc2, 8u) && tc2.get_translation() == t);
            cuda_unconstrain u(cuda_translate(hs, {0.5, -0.25, 0.125, 1.}), "weighted", {2., 3.}, {1e-3, 0.5});
            CHECK(same_problem(u, round_trip(u), 9u));
            cuda_decompose dc(z, {0.25, 0.75}, {0.1, -0.1}, "bi");
            const auto dc2 = round_trip(dc);
            CHECK(same_problem(dc, dc2, 10u) && dc2.get_z() == dc.get_z());
        }
    }

    // ---- 1. serialisation: every constructor argument survives save + load into a default-constructed UDA ----
    {
        cuda_sade a(37u, 7u, 2u, 1e-9, 1e-8, false, 4242u, 0);
        const auto b = round_trip(a);
        CHECK(b.get_gen() == 37u && b.get_seed() == 4242u);
        CHECK(b.descriptor().variant == 7u && b.descriptor().variant_adptv == 2u && b.descriptor().ftol == 1e-9 && b.descriptor().xtol == 1e-8);
        CHECK(b.descriptor().algo == a.descriptor().algo && b.get_name() == a.get_name());
        cuda_de1220 c(5u, {3u, 9u, 17u}, 1u, 1e-6, 1e-6, false, 7u, 0);
        const auto d = round_trip(c);
        CHECK(d.descriptor().n_allowed == 3u && d.descriptor().allowed_variants[0] == 3u && d.descriptor().allowed_variants[2] == 17u);
        cuda_sga e(4u, .7, 2., 0.05, .5, 3u, "sbx", "gaussian", "truncated", 99u, 0);
        const auto f = round_trip(e);
        CHECK(f.descriptor().crossover == e.descriptor().crossover && f.descriptor().mutation == e.descriptor().mutation
              && f.descriptor().selection == e.descriptor().selection && f.descriptor().param_s == 3u && f.descriptor().cr == .7);
        cuda_pso_gen g(3u, 0.6, 1.9, 2.1, 0.4, 3u, 1u, 6u, false, 5u, 0);
        const auto h = round_trip(g);
        CHECK(h.descriptor().omega == 0.6 && h.descriptor().eta2 == 2.1 && h.descriptor().neighb_type == 1u && h.descriptor().neighb_param == 6u);
        cuda_cmaes cm(9u, 0.3, 0.4, -1, 0.2, 0.25, 1e-7, 1e-9, false, true, 77u, 0);
        const auto cm2 = round_trip(cm);
        CHECK(cm2.descriptor().cma_cc == 0.3 && cm2.descriptor().cma_cs == 0.4 && cm2.descriptor().cma_c1 == -1 && cm2.descriptor().cma_cmu == 0.2
              && cm2.descriptor().sigma0 == 0.25 && cm2.descriptor().force_bounds == 1u && cm2.descriptor().xtol == 1e-9);
        // same results before and after the round trip
        pagmo::population pop{pagmo::problem{pagmo::rastrigin{10u}}, 32u, 5u};
        CHECK(same_population(pagmo::algorithm{a}.evolve(pop), pagmo::algorithm{b}.evolve(pop)));
        std::printf("serialisation round trips ok\n");
    }

    // ---- 2. cuda_island behind pagmo::island == thread_island running the same cuda UDA (thread_island.cpp:79-132) ----
    {
        pagmo::problem prob{pagmo::rastrigin{10u}};
        pagmo::population pop{prob, 64u, 23u};
        pagmo::island a{cuda_island{0}, pagmo::algorithm{cuda_de1220{20u, {2u, 3u, 7u, 10u, 13u, 14u, 15u, 16u}, 1u, 1e-12, 1e-12, false, 41u}}, pop};
        pagmo::island b{pagmo::thread_island{}, pagmo::algorithm{cuda_de1220{20u, {2u, 3u, 7u, 10u, 13u, 14u, 15u, 16u}, 1u, 1e-12, 1e-12, false, 41u}}, pop};
        for (int k = 0; k < 3; ++k) {
            a.evolve();
            b.evolve();
            a.wait_check();
            b.wait_check();
            CHECK(same_population(a.get_population(), b.get_population()));
        }
        const auto *udi = a.extract<cuda_island>();
        CHECK(udi != nullptr);
        if (udi) {
            const auto c = udi->get_counters(); // 3 runs, the population was uploaded once: it stayed on the device
            CHECK(c.first == 3u && c.second == 1u);
            std::printf("cuda_island: %llu run_evolve calls, %llu uploads\n", c.first, c.second);
        }
        CHECK(a.get_population().get_problem().get_fevals() == b.get_population().get_problem().get_fevals());
        CHECK(a.get_name().find("CUDA island") != std::string::npos);
        // cuda_cmaes through a plain pagmo::algorithm and through a cuda_island: the same population, and it gets better
        {
            pagmo::population p2{pagmo::problem{pagmo::rastrigin{8u}}, 24u, 3u};
            const double before = p2.get_f()[p2.best_idx()][0];
            const auto direct = pagmo::algorithm{cuda_cmaes{60u, -1, -1, -1, -1, 0.5, 0., 0., false, false, 9u}}.evolve(p2);
            pagmo::island ci{cuda_island{0}, pagmo::algorithm{cuda_cmaes{60u, -1, -1, -1, -1, 0.5, 0., 0., false, false, 9u}}, p2};
            ci.evolve();
            ci.wait_check();
            CHECK(same_population(ci.get_population(), direct));
            CHECK(direct.get_f()[direct.best_idx()][0] < before);
            std::printf("cuda_cmaes: best %g -> %g\n", before, direct.get_f()[direct.best_idx()][0]);
        }
        // the other descriptor-based UDAs (xnes, nspso), sade with memory (state resident in the island), and the UDAs with their own
        // entry point (gaco, maco, moead_gen run through their own evolve()): an island gives what the algorithm alone gives
        {
            pagmo::population so{pagmo::problem{pagmo::rastrigin{8u}}, 24u, 3u}, mo{pagmo::problem{pagmo::zdt{1u, 10u}}, 24u, 3u};
            const auto through_island = [](const pagmo::algorithm &a, const pagmo::population &p, unsigned times) {
                pagmo::island isl{cuda_island{0}, a, p};
                for (unsigned t = 0; t < times; ++t) {
                    isl.evolve();
                    isl.wait_check();
                }
                return isl.get_population();
            };
            const auto alone = [](pagmo::algorithm a, pagmo::population p, unsigned times) {
                for (unsigned t = 0; t < times; ++t) p = a.evolve(p);
                return p;
            };
            for (const auto &a : {pagmo::algorithm{cuda_xnes{10u, -1, -1, -1, -1, 0., 0., false, false, 9u}},
                                  pagmo::algorithm{cuda_sade{6u, 2u, 1u, 0., 0., true, 9u}},
                                  pagmo::algorithm{cuda_gaco{6u, 8u, 1.0, 1e9, 0.01, 1u, 7u, 100000u, 100000u, 0., false, 9u}}})
                CHECK(same_population(through_island(a, so, 2u), alone(a, so, 2u)));
            for (const auto &a : {pagmo::algorithm{cuda_nspso{5u, 0.6, 2., 2., 1., 0.5, 60u, "crowding distance", false, 9u}},
                                  pagmo::algorithm{cuda_maco{5u, 8u, 1.0, 1u, 7u, 100000u, 0., false, 9u}},
                                  pagmo::algorithm{cuda_moead_gen{4u, "grid", "tchebycheff", 5u, 1.0, 0.5, 20., 0.9, 2u, true, 9u}}})
                CHECK(same_population(through_island(a, mo, 2u), alone(a, mo, 2u)));
        }
        // a stock CPU algorithm is refused (no CPU fallback)
        pagmo::island c{cuda_island{0}, pagmo::algorithm{pagmo::de{5u}}, pop};
        c.evolve();
        bool threw = false;
        try {
            c.wait_check();
        } catch (const std::invalid_argument &) {
            threw = true;
        }
        CHECK(threw);
    }

    // ---- 3. a STOCK pagmo::archipelago{ring} of 8 cuda_islands: pagmo's own island threads, migration database and policies ----
    {
        pagmo::archipelago archi{pagmo::ring{}};
        for (unsigned i = 0; i < 8u; ++i) {
            archi.push_back(cuda_island{static_cast<int>(i % static_cast<unsigned>(ndev))}, pagmo::algorithm{cuda_sade{25u, 2u, 1u, 1e-30, 1e-30, false, 100u + i}},
                            pagmo::problem{cuda_rastrigin{20u}}, 64u, pagmo::fair_replace{}, pagmo::select_best{}, 7u + i);
        }
        std::vector<double> before;
        for (const auto &f : archi.get_champions_f()) before.push_back(f[0]);
        archi.evolve(6u);
        archi.wait_check();
        const auto after = archi.get_champions_f();
        CHECK(after.size() == 8u);
        for (std::size_t i = 0; i < after.size(); ++i) CHECK(after[i][0] <= before[i]);
        const auto log = archi.get_migration_log();
        CHECK(!log.empty());
        for (const auto &e : log) { // every migration runs along a ring edge (ring.cpp:74-116)
            const auto s = std::get<4>(e), d = std::get<5>(e);
            CHECK((s + 1u) % 8u == d || (d + 1u) % 8u == s);
        }
        std::printf("stock archipelago{ring}: %zu islands, %zu migrations, best %g -> %g\n", after.size(), log.size(),
                    *std::min_element(before.begin(), before.end()), std::min_element(after.begin(), after.end(), [](auto &a, auto &b) { return a[0] < b[0]; })->at(0));
    }

    // ---- 4. the device-resident archipelago: reproducible, migrates along topology edges, champions never get worse ----
    auto run_cuda_archi = [&](int devices, pagmo::migration_type mt, pagmo::migrant_handling mh, const pagmo::topology &topo, unsigned rounds) {
        cuda_archipelago archi{topo, 99u};
        archi.set_migration_type(mt);
        archi.set_migrant_handling(mh);
        for (unsigned i = 0; i < 8u; ++i) {
            archi.push_back(static_cast<int>(i % static_cast<unsigned>(devices)), pagmo::algorithm{cuda_sade{10u, 2u, 1u, 1e-30, 1e-30, false, 500u + i}},
                            pagmo::problem{cuda_rastrigin{12u}}, 48u, pagmo::fair_replace{2}, pagmo::select_best{2}, 300u + i);
        }
        std::vector<std::vector<pagmo::vector_double>> champs;
        champs.push_back(archi.get_champions_f());
        for (unsigned r = 0; r < rounds; ++r) {
            archi.evolve();
            archi.wait_check();
            champs.push_back(archi.get_champions_f());
        }
        std::vector<pagmo::individuals_group_t> pops;
        for (std::size_t i = 0; i < archi.size(); ++i) pops.push_back(archi.get_individuals(i));
        return std::make_tuple(champs, pops, archi.get_migration_log());
    };
    {
        const auto ring = pagmo::topology{pagmo::ring{}};
        const auto r1 = run_cuda_archi(1, pagmo::migration_type::p2p, pagmo::migrant_handling::preserve, ring, 6u);
        const auto r2 = run_cuda_archi(1, pagmo::migration_type::p2p, pagmo::migrant_handling::preserve, ring, 6u);
        CHECK(std::get<1>(r1) == std::get<1>(r2)); // same seeds, same result: bit for bit
        const auto &champs = std::get<0>(r1);
        for (std::size_t r = 1; r < champs.size(); ++r)
            for (std::size_t i = 0; i < 8u; ++i) CHECK(champs[r][i][0] <= champs[r - 1][i][0]);
        const auto &log = std::get<2>(r1);
        CHECK(!log.empty());
        std::set<unsigned long long> seen;
        for (const auto &e : log) {
            const auto s = std::get<4>(e), d = std::get<5>(e);
            CHECK((s + 1u) % 8u == d || (d + 1u) % 8u == s);
            // the logged individual is in the destination's final or some earlier population: its fitness is what was logged
            CHECK(std::get<3>(e).size() == 1u && std::isfinite(std::get<3>(e)[0]));
            seen.insert(std::get<1>(e));
        }
        std::printf("cuda_archipelago{ring, p2p, preserve}: %zu migrations (%zu distinct migrants)\n", log.size(), seen.size());
        // broadcast + evict on a fully connected topology: more traffic, an extracted database entry serves one puller only
        const auto full = pagmo::topology{pagmo::fully_connected{}};
        const auto b1 = run_cuda_archi(1, pagmo::migration_type::broadcast, pagmo::migrant_handling::evict, full, 4u);
        const auto b2 = run_cuda_archi(1, pagmo::migration_type::broadcast, pagmo::migrant_handling::preserve, full, 4u);
        CHECK(std::get<2>(b2).size() > std::get<2>(b1).size());
        CHECK(std::get<2>(b1).size() > 0u);
        // unconnected: no migration, every island is an independent run
        const auto u1 = run_cuda_archi(1, pagmo::migration_type::p2p, pagmo::migrant_handling::preserve, pagmo::topology{pagmo::unconnected{}}, 3u);
        CHECK(std::get<2>(u1).empty());
        if (ndev >= 2) { // islands spread over GPUs: the migrants go over NCCL and the run is the SAME run
            const auto n1 = run_cuda_archi(ndev, pagmo::migration_type::p2p, pagmo::migrant_handling::preserve, ring, 6u);
            CHECK(std::get<1>(n1) == std::get<1>(r1));
            CHECK(std::get<2>(n1).size() == log.size());
            const auto nb = run_cuda_archi(ndev, pagmo::migration_type::broadcast, pagmo::migrant_handling::evict, full, 4u);
            CHECK(std::get<1>(nb) == std::get<1>(b1));
            std::printf("cuda_archipelago over %d GPUs (NCCL send/recv) == 1 GPU, bit for bit\n", ndev);
        }
    }

    if (g_fail) {
        std::printf("%d checks failed\n", g_fail);
        return 1;
    }
    std::printf("ISLANDS OK\n");
    return 0;
}
