// cuda_island.hpp - GPU islands behind pagmo::island / pagmo::archipelago (header-only, over the C ABI in pgc.h).
//
//   pagmo_cuda::cuda_island       a user-defined island (UDI, reference include/pagmo/island.hpp:92-141): run_evolve(island&)
//                                 as thread_island does it (src/islands/thread_island.cpp:79-132), but the population lives in HBM
//                                 between calls.  It drops into a STOCK pagmo::archipelago: pagmo's own island thread, migration
//                                 database, policies and topology keep working on the host (src/island.cpp:428-652); the host
//                                 copy is refreshed after every run_evolve (island::set_population is the UDI contract) and the
//                                 device copy is re-uploaded only when migration changed the host population.
//   pagmo_cuda::cuda_archipelago  an archipelago whose islands NEVER leave the device: populations, select_best, fair_replace and
//                                 the migrants themselves stay in HBM; migrants travel along the edges of a pagmo::topology as
//                                 ncclSend / ncclRecv between GPUs (pgc_migrate), one island per GPU or several per GPU, one process
//                                 for all GPUs or one process per GPU.  The decisions of island.cpp:461-620 (which in-edge,
//                                 Bernoulli(weight), p2p / broadcast, preserve / evict) are taken on the host as the reference takes
//                                 them, from Philox draws (seed, tag 9, round, island, slot) so that every process replays them
//                                 identically.  Islands advance in lock step: round r pulls what round r-1 published (the
//                                 reference's island threads race, island.cpp:461); with evict the islands pull in index order.
//
// Only the pagmo_cuda:: UDAs (cuda_algorithms.hpp) can run on these islands and only problems with a device evaluator
// (cuda_bfe.hpp); anything else throws std::invalid_argument - there is no CPU fallback.
#ifndef PAGMO_CUDA_CUDA_ISLAND_HPP
#define PAGMO_CUDA_CUDA_ISLAND_HPP

#include <cstring>
#include <exception>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

#include <pagmo/algorithm.hpp>
#include <pagmo/archipelago.hpp>
#include <pagmo/exceptions.hpp>
#include <pagmo/island.hpp>
#include <pagmo/population.hpp>
#include <pagmo/problem.hpp>
#include <pagmo/r_policies/fair_replace.hpp>
#include <pagmo/s11n.hpp>
#include <pagmo/s_policies/select_best.hpp>
#include <pagmo/threading.hpp>
#include <pagmo/topologies/unconnected.hpp>
#include <pagmo/topology.hpp>
#include <pagmo/types.hpp>

#include <pagmo_cuda/cuda_algorithms.hpp>
#include <pagmo_cuda/cuda_bfe.hpp>
#include <pagmo_cuda/pgc.h>

namespace pagmo_cuda
{

namespace detail
{

// the pagmo_cuda UDA inside a type-erased pagmo::algorithm, or nullptr
inline const cuda_algorithm_base *find_cuda_uda(const pagmo::algorithm &a)
{
#define PGC_TRY_UDA(T)                                                                                                 \
    if (const auto *p = a.extract<T>()) return p;
    PGC_TRY_UDA(cuda_de)
    PGC_TRY_UDA(cuda_sade)
    PGC_TRY_UDA(cuda_de1220)
    PGC_TRY_UDA(cuda_pso_gen)
    PGC_TRY_UDA(cuda_pso)
    PGC_TRY_UDA(cuda_nsga2)
    PGC_TRY_UDA(cuda_sga)
    PGC_TRY_UDA(cuda_cmaes)
    PGC_TRY_UDA(cuda_xnes)
    PGC_TRY_UDA(cuda_nspso)
#undef PGC_TRY_UDA
    return nullptr;
}

// the pagmo_cuda UDAs with their own entry point (no pgc_algo_desc): an island runs them through their own evolve()
inline bool is_standalone_cuda_uda(const pagmo::algorithm &a)
{
    return a.is<cuda_gaco>() || a.is<cuda_maco>() || a.is<cuda_moead_gen>();
}

// RAII pgc_island
class island_handle
{
public:
    island_handle(std::shared_ptr<problem_handle> prob, std::size_t n, std::size_t max_migrants, std::size_t max_in_edges)
        : m_prob(std::move(prob)), m_n(n)
    {
        check(pgc_island_create(m_prob->raw(), n, max_migrants, max_in_edges, &m_isl), "pgc_island_create");
    }
    ~island_handle()
    {
        pgc_island_destroy(m_isl);
    }
    island_handle(const island_handle &) = delete;
    island_handle &operator=(const island_handle &) = delete;
    pgc_island *raw() const { return m_isl; }
    const std::shared_ptr<problem_handle> &problem() const { return m_prob; }
    std::size_t size() const { return m_n; }

private:
    std::shared_ptr<problem_handle> m_prob; // declared first: destroyed after the island that borrows it
    std::size_t m_n;
    pgc_island *m_isl = nullptr;
};

// the migration rate of a fair_replace / select_best policy as the C ABI takes it
inline std::pair<int, double> policy_rate(const boost::variant<pagmo::pop_size_t, double> &r)
{
    if (r.which() == 0) return {0, static_cast<double>(boost::get<pagmo::pop_size_t>(r))};
    return {1, boost::get<double>(r)};
}

} // namespace detail

// ---------------------------------------------------------------------------------------------------------------------------
// The UDI.
class cuda_island
{
public:
    explicit cuda_island(int device = 0) : m_device(device), m_state(std::make_shared<state>()) {}
    // a copy is another island: it starts without a resident population
    cuda_island(const cuda_island &o) : m_device(o.m_device), m_state(std::make_shared<state>()) {}
    cuda_island(cuda_island &&) noexcept = default;
    cuda_island &operator=(const cuda_island &o)
    {
        if (this != &o) {
            m_device = o.m_device;
            m_state = std::make_shared<state>();
        }
        return *this;
    }
    cuda_island &operator=(cuda_island &&) noexcept = default;

    void run_evolve(pagmo::island &isl) const
    {
        // copies of the island's algorithm and population, as thread_island.cpp:97-123
        auto algo = isl.get_algorithm();
        auto pop = isl.get_population();
        if (algo.get_thread_safety() < pagmo::thread_safety::basic) {
            pagmo_throw(std::invalid_argument, "the 'cuda_island' UDI requires an algorithm providing at least the 'basic' thread safety "
                                               "guarantee, but an algorithm of type '"
                                                   + algo.get_name() + "' does not");
        }
        const auto *uda = detail::find_cuda_uda(algo);
        if (!uda && detail::is_standalone_cuda_uda(algo)) {
            // cuda_gaco / cuda_maco / cuda_moead_gen keep their own state and entry point: the device work is theirs, the island only
            // does what thread_island does around it (thread_island.cpp:118-131)
            isl.set_population(algo.evolve(pop));
            isl.set_algorithm(algo);
            return;
        }
        if (!uda) {
            pagmo_throw(std::invalid_argument, "the 'cuda_island' UDI runs the pagmo_cuda:: algorithms only (cuda_de, cuda_sade, cuda_de1220, "
                                               "cuda_pso_gen, cuda_nsga2, cuda_sga, cuda_cmaes, cuda_xnes, cuda_nspso, cuda_gaco, cuda_maco, "
                                               "cuda_moead_gen); an algorithm of type '"
                                                   + algo.get_name() + "' was given and there is no CPU fallback");
        }
        const auto &prob = pop.get_problem();
        const auto n = pop.size();
        const auto &desc = uda->descriptor();
        if (desc.gens == 0u || n == 0u) { // the UDAs return the population unchanged (de.cpp:118-120)
            isl.set_population(pop);
            isl.set_algorithm(algo);
            return;
        }
        if (prob.get_nc() != 0u) {
            pagmo_throw(std::invalid_argument, "Non linear constraints detected in " + prob.get_name() + " instance. " + algo.get_name()
                                                   + " cannot deal with them");
        }
        if (prob.is_stochastic()) {
            pagmo_throw(std::invalid_argument, "The problem appears to be stochastic " + algo.get_name() + " cannot deal with it");
        }
        state &s = *m_state;
        std::lock_guard<std::mutex> lk(s.mtx);
        const auto h = s.cache.find(prob, m_device);
        if (!h) {
            pagmo_throw(std::invalid_argument, "cuda_island cannot evolve a population of '" + prob.get_name()
                                                   + "': no CUDA evaluator exists for this UDP type; there is no CPU fallback");
        }
        const auto nx = prob.get_nx(), nf = prob.get_nf();
        // flat image of the host population
        std::vector<unsigned long long> ids(pop.get_ID());
        pagmo::vector_double x(n * nx), f(n * nf);
        for (decltype(pop.size()) i = 0; i < n; ++i) {
            std::copy(pop.get_x()[i].begin(), pop.get_x()[i].end(), x.begin() + static_cast<std::ptrdiff_t>(i * nx));
            std::copy(pop.get_f()[i].begin(), pop.get_f()[i].end(), f.begin() + static_cast<std::ptrdiff_t>(i * nf));
        }
        std::lock_guard<std::mutex> dev_lk(detail::device_mutex(h->device())); // the shared context of the device: one stream
        if (!s.isl || s.isl->problem() != h || s.isl->size() != n) {
            s.isl = std::make_unique<detail::island_handle>(h, n, 1u, 1u);
            s.ids.clear();
        }
        // the device copy is stale only if something (migration, the user) touched the host population since the last run
        const bool same = s.ids == ids && s.x == x && s.f == f;
        if (!same) {
            detail::check(pgc_island_upload(s.isl->raw(), reinterpret_cast<const uint64_t *>(ids.data()), x.data(), f.data()), "pgc_island_upload");
            ++s.uploads;
        }
        detail::check(pgc_island_set_generation(s.isl->raw(), uda->generation()), "pgc_island_set_generation");
        unsigned done = 0;
        detail::check(pgc_island_evolve(s.isl->raw(), &desc, &done), "pgc_island_evolve");
        uda->advance_generation(desc.gens);
        detail::check(pgc_island_download(s.isl->raw(), nullptr, x.data(), f.data()), "pgc_island_download");
        ++s.runs;
        s.ids = std::move(ids);
        s.x = x;
        s.f = f;
        for (decltype(pop.size()) i = 0; i < n; ++i) {
            pop.set_xf(i, pagmo::vector_double(x.begin() + static_cast<std::ptrdiff_t>(i * nx), x.begin() + static_cast<std::ptrdiff_t>((i + 1) * nx)),
                       pagmo::vector_double(f.begin() + static_cast<std::ptrdiff_t>(i * nf), f.begin() + static_cast<std::ptrdiff_t>((i + 1) * nf)));
        }
        prob.increment_fevals(static_cast<unsigned long long>(done) * n);
        isl.set_population(pop);
        isl.set_algorithm(algo);
    }
    std::string get_name() const
    {
        return "CUDA island (sm_100a, device " + std::to_string(m_device) + ")";
    }
    std::string get_extra_info() const
    {
        std::lock_guard<std::mutex> lk(m_state->mtx);
        return "\tDevice-resident population: " + std::string(m_state->isl ? "yes" : "not yet") + "\n\trun_evolve calls: "
               + std::to_string(m_state->runs) + "\n\tHost -> device population uploads: " + std::to_string(m_state->uploads);
    }
    // run_evolve calls so far / how many of them had to upload the host population (tests, diagnostics)
    std::pair<unsigned long long, unsigned long long> get_counters() const
    {
        std::lock_guard<std::mutex> lk(m_state->mtx);
        return {m_state->runs, m_state->uploads};
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_device); // the resident population is a cache of the island's own population
    }

private:
    struct state {
        std::mutex mtx;
        detail::twin_cache cache;
        std::unique_ptr<detail::island_handle> isl;
        std::vector<unsigned long long> ids; // host mirror of what the device holds
        pagmo::vector_double x, f;
        unsigned long long runs = 0, uploads = 0;
    };
    int m_device;
    std::shared_ptr<state> m_state;
};

// ---------------------------------------------------------------------------------------------------------------------------
// The device-resident archipelago.
class cuda_archipelago
{
public:
    using size_type = std::size_t;
    // (round, ID, decision vector, fitness vector, source island, destination island): archipelago::migration_entry_t
    // (archipelago.hpp:187) with the round index where the reference stores a wall-clock timestamp
    using migration_entry_t = std::tuple<double, unsigned long long, pagmo::vector_double, pagmo::vector_double, size_type, size_type>;
    using migration_log_t = std::vector<migration_entry_t>;

    // one process drives every GPU it names in push_back (NCCL communicator created on first use, pgc_comm_init)
    explicit cuda_archipelago(pagmo::topology t = pagmo::topology{pagmo::unconnected{}}, unsigned long long seed = 0u)
        : m_topology(std::move(t)), m_seed(seed)
    {
    }
    // one process per GPU: `unique_id` = the 128 bytes rank 0 got from pgc_comm_unique_id and the application distributed.
    // Every process pushes back ALL islands in the same order, naming the rank that owns each; it materialises its own only.
    cuda_archipelago(pagmo::topology t, unsigned long long seed, int nranks, int rank, int device, const std::vector<unsigned char> &unique_id)
        : m_topology(std::move(t)), m_seed(seed), m_nranks(nranks), m_rank(rank)
    {
        pgc_comm *c = nullptr;
        detail::check(pgc_comm_init_rank(device, nranks, rank, unique_id.data(), unique_id.size(), &c), "pgc_comm_init_rank");
        m_comm.reset(c, [](pgc_comm *p) { pgc_comm_destroy(p); });
    }

    // archipelago::push_back(algo, prob, pop_size, r_pol, s_pol, seed) (archipelago.hpp:430-470) for an island on `device`.
    // Multi-process archipelagos name the owning rank as well; single-process ones leave owner_rank = -1.
    void push_back(int device, const pagmo::algorithm &algo, const pagmo::problem &prob, size_type pop_size,
                   const pagmo::fair_replace &r_pol = pagmo::fair_replace{}, const pagmo::select_best &s_pol = pagmo::select_best{},
                   unsigned long long seed = 0u, int owner_rank = -1)
    {
        const auto *uda = detail::find_cuda_uda(algo);
        if (!uda) {
            pagmo_throw(std::invalid_argument, "cuda_archipelago runs the pagmo_cuda:: algorithms only; an algorithm of type '" + algo.get_name()
                                                   + "' was given and there is no CPU fallback");
        }
        if (m_nranks > 0 && (owner_rank < 0 || owner_rank >= m_nranks)) {
            pagmo_throw(std::invalid_argument, "cuda_archipelago::push_back: a multi-process archipelago needs the owning rank of every island");
        }
        entry e;
        e.device = device;
        e.owner = owner_rank;
        e.desc = uda->descriptor();
        e.n = pop_size;
        e.r_rate = detail::policy_rate(r_pol.get_migr_rate());
        e.s_rate = detail::policy_rate(s_pol.get_migr_rate());
        e.nx = prob.get_nx();
        e.nf = prob.get_nf();
        // the number of migrants this island publishes (select_best.cpp:80-100), known to every process
        e.k_out = e.s_rate.first ? static_cast<size_type>(e.s_rate.second * static_cast<double>(pop_size))
                                 : static_cast<size_type>(e.s_rate.second);
        if (e.k_out > pop_size) e.k_out = pop_size;
        e.local = m_nranks == 0 || owner_rank == m_rank;
        if (e.local) {
            const auto shared = m_cache.find(prob, device);
            if (!shared) {
                pagmo_throw(std::invalid_argument, "cuda_archipelago cannot host a population of '" + prob.get_name()
                                                       + "': no CUDA evaluator exists for this UDP type; there is no CPU fallback");
            }
            e.prob = shared->private_on(device); // own context = own stream: islands sharing a GPU overlap on it
            e.seed = seed;
        }
        m_islands.push_back(std::move(e));
        m_topology.push_back(); // archipelago.cpp:283-301: the topology grows with the archipelago
        m_built = false;
    }

    size_type size() const { return m_islands.size(); }
    pagmo::migration_type get_migration_type() const { return m_migr_type; }
    void set_migration_type(pagmo::migration_type mt) { m_migr_type = mt; }
    pagmo::migrant_handling get_migrant_handling() const { return m_migr_handling; }
    void set_migrant_handling(pagmo::migrant_handling mh) { m_migr_handling = mh; }
    const pagmo::topology &get_topology() const { return m_topology; }
    const migration_log_t &get_migration_log() const { return m_log; }
    void set_migration_log_enabled(bool on) { m_log_on = on; }
    unsigned long long get_round() const { return m_round; }

    // archipelago::evolve(n) followed by wait_check(): n rounds of (pull migrants, replace, evolve, select + publish) per island
    void evolve(unsigned n = 1u)
    {
        build();
        for (unsigned r = 0; r < n; ++r) round();
    }
    void wait_check() const {} // evolve() is synchronous

    // the individuals of LOCAL island i as an individuals_group_t (ids, dvs, fvs) - island::get_population's content
    pagmo::individuals_group_t get_individuals(size_type i)
    {
        build();
        entry &e = local_entry(i);
        std::vector<unsigned long long> ids(e.n);
        pagmo::vector_double x(e.n * e.nx), f(e.n * e.nf);
        detail::check(pgc_island_download(e.isl->raw(), reinterpret_cast<uint64_t *>(ids.data()), x.data(), f.data()), "pgc_island_download");
        pagmo::individuals_group_t g;
        std::get<0>(g) = std::move(ids);
        for (size_type j = 0; j < e.n; ++j) {
            std::get<1>(g).emplace_back(x.begin() + static_cast<std::ptrdiff_t>(j * e.nx), x.begin() + static_cast<std::ptrdiff_t>((j + 1) * e.nx));
            std::get<2>(g).emplace_back(f.begin() + static_cast<std::ptrdiff_t>(j * e.nf), f.begin() + static_cast<std::ptrdiff_t>((j + 1) * e.nf));
        }
        return g;
    }
    // archipelago::get_champions_f / get_champions_x (archipelago.cpp:520-560) over the LOCAL single-objective islands
    std::vector<pagmo::vector_double> get_champions_f()
    {
        return champions(false);
    }
    std::vector<pagmo::vector_double> get_champions_x()
    {
        return champions(true);
    }
    bool is_local(size_type i) const { return m_islands.at(i).local; }

private:
    struct entry {
        int device = 0, owner = -1;
        bool local = true;
        pgc_algo_desc desc{};
        size_type n = 0, nx = 0, nf = 0, k_out = 0;
        std::pair<int, double> r_rate{0, 1.}, s_rate{0, 1.};
        unsigned long long seed = 0;
        std::shared_ptr<detail::problem_handle> prob;
        std::unique_ptr<detail::island_handle> isl;
        bool published = false; // its database entry holds migrants (false until the first select, and after an eviction)
    };

    entry &local_entry(size_type i)
    {
        entry &e = m_islands.at(i);
        if (!e.local) pagmo_throw(std::invalid_argument, "cuda_archipelago: island " + std::to_string(i) + " lives in another process");
        return e;
    }

    std::vector<pagmo::vector_double> champions(bool want_x)
    {
        build();
        std::vector<pagmo::vector_double> out;
        for (auto &e : m_islands) {
            if (!e.local) continue;
            pagmo::vector_double x(e.nx), f(1);
            detail::check(pgc_island_champion(e.isl->raw(), x.data(), f.data()), "pgc_island_champion");
            out.push_back(want_x ? x : f);
        }
        return out;
    }

    // islands, communicator and in-edge lists; called before anything touches the device
    void build()
    {
        if (m_built) return;
        const size_type G = m_islands.size();
        m_conn.assign(G, {});
        size_type cap = 1, max_in = 1;
        for (size_type i = 0; i < G; ++i) {
            m_conn[i] = m_topology.get_connections(i);
            max_in = std::max(max_in, m_conn[i].first.size());
            cap = std::max(cap, m_islands[i].k_out);
        }
        // communicator ranks: one per GPU.  Single process: rank = position of the device among the devices in use.
        if (m_nranks == 0) {
            std::vector<int> devs;
            for (auto &e : m_islands)
                if (std::find(devs.begin(), devs.end(), e.device) == devs.end()) devs.push_back(e.device);
            for (auto &e : m_islands) e.owner = static_cast<int>(std::find(devs.begin(), devs.end(), e.device) - devs.begin());
            if (devs.size() > 1u && !m_comm) {
                pgc_comm *c = nullptr;
                detail::check(pgc_comm_init(static_cast<int>(devs.size()), devs.data(), &c), "pgc_comm_init");
                m_comm.reset(c, [](pgc_comm *p) { pgc_comm_destroy(p); });
            }
        }
        for (auto &e : m_islands) { // islands that share a GPU fill it together: each keeps fuller tiles (pgc_ctx_set_sharers)
            if (!e.local) continue;
            int sharers = 0;
            for (const auto &o : m_islands) sharers += (o.local && o.device == e.device) ? 1 : 0;
            detail::check(pgc_ctx_set_sharers(e.prob->context(), sharers), "pgc_ctx_set_sharers");
        }
        for (auto &e : m_islands) {
            if (!e.local || e.isl) continue;
            e.isl = std::make_unique<detail::island_handle>(e.prob, e.n, cap, max_in);
            detail::check(pgc_island_init(e.isl->raw(), e.seed), "pgc_island_init"); // population(prob, bfe, n, seed) on the device
        }
        m_raw.assign(G, nullptr);
        m_owner.assign(G, 0);
        for (size_type i = 0; i < G; ++i) {
            m_raw[i] = m_islands[i].local ? m_islands[i].isl->raw() : nullptr;
            m_owner[i] = m_islands[i].owner;
        }
        m_built = true;
    }

    double draw(size_type island, unsigned slot) const
    {
        double u = 0.;
        detail::check(pgc_philox_u01(m_seed, 9u /* migration decisions */, static_cast<uint32_t>(m_round), static_cast<uint32_t>(island), slot, &u),
                      "pgc_philox_u01");
        return u;
    }

    void round()
    {
        const size_type G = m_islands.size();
        // ---- island.cpp:461-575 for every destination, in island order: the edges that carry migrants this round
        struct pull {
            bool replace = false;                              // r_pol.replace runs (possibly with no migrants)
            std::vector<std::pair<size_type, bool>> sources;   // (source island, its database entry still held migrants)
        };
        std::vector<pull> pulls(G);
        std::vector<uint32_t> es, ed, eslot;
        for (size_type g = 0; g < G; ++g) {
            const auto &src = m_conn[g].first;
            const auto &w = m_conn[g].second;
            if (src.empty()) continue;
            auto take = [&](size_type s) {
                const bool had = m_islands[s].published;
                if (m_migr_handling == pagmo::migrant_handling::evict) m_islands[s].published = false; // extract_migrants empties the entry
                pulls[g].sources.emplace_back(s, had);
            };
            if (m_migr_type == pagmo::migration_type::p2p) {
                size_type j = static_cast<size_type>(draw(g, 0u) * static_cast<double>(src.size())); // :497-499
                if (j >= src.size()) j = src.size() - 1u;
                if (!(draw(g, 1u) < w[j])) continue; // :502
                pulls[g].replace = true;
                take(src[j]);
            } else {
                pulls[g].replace = true;
                for (size_type j = 0; j < src.size(); ++j)
                    if (draw(g, static_cast<unsigned>(j)) < w[j]) take(src[j]); // :551-575
            }
            for (size_type q = 0; q < pulls[g].sources.size(); ++q) {
                if (!pulls[g].sources[q].second) continue; // an empty database entry: nothing travels
                es.push_back(static_cast<uint32_t>(pulls[g].sources[q].first));
                ed.push_back(static_cast<uint32_t>(g));
                eslot.push_back(static_cast<uint32_t>(q));
            }
        }
        // slots whose source had nothing to give hold an empty group
        for (size_type g = 0; g < G; ++g) {
            if (!m_islands[g].local) continue;
            for (size_type q = 0; q < pulls[g].sources.size(); ++q)
                if (!pulls[g].sources[q].second)
                    detail::check(pgc_island_inbox_upload(m_raw[g], q, nullptr, nullptr, nullptr, 0u), "pgc_island_inbox_upload");
        }
        // ---- the migrants travel: outbox of the source -> inbox slot of the destination, device to device
        detail::check(pgc_migrate(m_comm.get(), m_raw.data(), m_owner.data(), G, es.data(), ed.data(), eslot.data(), es.size()), "pgc_migrate");
        // ---- replace, evolve, select on every local island, each from its own thread (its own stream)
        std::vector<std::thread> workers;
        std::vector<std::exception_ptr> errors(G);
        std::vector<migration_log_t> logs(G);
        for (size_type g = 0; g < G; ++g) {
            if (!m_islands[g].local) continue;
            workers.emplace_back([&, g]() {
                try {
                    step(g, pulls[g].replace, pulls[g].sources, logs[g]);
                } catch (...) {
                    errors[g] = std::current_exception();
                }
            });
        }
        for (auto &t : workers) t.join();
        for (auto &e : errors)
            if (e) std::rethrow_exception(e);
        for (auto &l : logs) m_log.insert(m_log.end(), l.begin(), l.end());
        for (auto &e : m_islands) e.published = e.k_out > 0u; // set_migrants: every island re-published (replicated knowledge)
        ++m_round;
    }

    void step(size_type g, bool replace, const std::vector<std::pair<size_type, bool>> &sources, migration_log_t &log)
    {
        entry &e = m_islands[g];
        pgc_island *isl = e.isl->raw();
        if (replace) { // r_pol.replace + set_individuals, island.cpp:505-517 / :577-585
            std::vector<uint64_t> acc_ids(m_log_on ? sources.size() * std::max<size_type>(1u, max_cap()) : 0u);
            std::vector<uint32_t> acc_slot(acc_ids.size());
            size_t n_acc = 0;
            detail::check(pgc_island_replace(isl, e.r_rate.first, e.r_rate.second, sources.size(), m_log_on ? acc_ids.data() : nullptr,
                                             m_log_on ? acc_slot.data() : nullptr, m_log_on ? &n_acc : nullptr),
                          "pgc_island_replace");
            if (n_acc) { // the migration log: immigrants that made it into the population, :525-536 / :594-607
                const auto inds = get_individuals_unlocked(e);
                for (size_t a = 0; a < n_acc; ++a) {
                    const auto &ids = std::get<0>(inds);
                    const auto it = std::find(ids.begin(), ids.end(), acc_ids[a]);
                    const auto pos = static_cast<size_type>(it - ids.begin());
                    log.emplace_back(static_cast<double>(m_round), acc_ids[a], std::get<1>(inds)[pos], std::get<2>(inds)[pos],
                                     sources[acc_slot[a]].first, g);
                }
            }
        }
        unsigned done = 0;
        detail::check(pgc_island_evolve(isl, &e.desc, &done), "pgc_island_evolve");                             // run_evolve, :623
        detail::check(pgc_island_select(isl, e.s_rate.first, e.s_rate.second, nullptr), "pgc_island_select");   // s_pol.select, :629-640
    }

    size_type max_cap() const
    {
        size_type c = 1;
        for (const auto &e : m_islands) c = std::max(c, e.k_out);
        return c;
    }

    pagmo::individuals_group_t get_individuals_unlocked(entry &e)
    {
        std::vector<unsigned long long> ids(e.n);
        pagmo::vector_double x(e.n * e.nx), f(e.n * e.nf);
        detail::check(pgc_island_download(e.isl->raw(), reinterpret_cast<uint64_t *>(ids.data()), x.data(), f.data()), "pgc_island_download");
        pagmo::individuals_group_t g;
        std::get<0>(g) = std::move(ids);
        for (size_type j = 0; j < e.n; ++j) {
            std::get<1>(g).emplace_back(x.begin() + static_cast<std::ptrdiff_t>(j * e.nx), x.begin() + static_cast<std::ptrdiff_t>((j + 1) * e.nx));
            std::get<2>(g).emplace_back(f.begin() + static_cast<std::ptrdiff_t>(j * e.nf), f.begin() + static_cast<std::ptrdiff_t>((j + 1) * e.nf));
        }
        return g;
    }

    pagmo::topology m_topology;
    unsigned long long m_seed = 0;
    int m_nranks = 0, m_rank = 0;
    std::shared_ptr<pgc_comm> m_comm;
    detail::twin_cache m_cache;
    std::vector<entry> m_islands;
    std::vector<std::pair<std::vector<std::size_t>, pagmo::vector_double>> m_conn;
    std::vector<pgc_island *> m_raw;
    std::vector<int> m_owner;
    pagmo::migration_type m_migr_type = pagmo::migration_type::p2p;
    pagmo::migrant_handling m_migr_handling = pagmo::migrant_handling::preserve;
    migration_log_t m_log;
    bool m_log_on = true, m_built = false;
    unsigned long long m_round = 0;
};

} // namespace pagmo_cuda

PAGMO_S11N_ISLAND_EXPORT_KEY(pagmo_cuda::cuda_island)

#endif
