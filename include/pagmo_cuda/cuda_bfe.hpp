// cuda_bfe.hpp - header-only C++ adapters that put libpgc.so behind pagmo's own plugin API.
//
//   pagmo_cuda::cuda_bfe        a user-defined batch fitness evaluator (UDBFE, reference include/pagmo/bfe.hpp:68-108):
//                               `pagmo::bfe{cuda_bfe{}}`, `algo.set_bfe(...)`, `population{prob, cuda_bfe{}, n}`.
//   pagmo_cuda::cuda_zdt/_dtlz/_wfg  CUDA-backed multi-objective UDPs (same constructor arguments as pagmo::zdt / dtlz / wfg)
//   pagmo_cuda::cuda_cec2013, cuda_lennard_jones   likewise (cec2013 takes its data tables, like cuda_cec2014)
//   pagmo_cuda::cuda_cec2014    CUDA-backed UDPs (reference include/pagmo/problem.hpp:394-411,532-553): mandatory
//   pagmo_cuda::cuda_simple<F>  fitness()/get_bounds() plus batch_fitness(), so pagmo::default_bfe / member_bfe pick the
//                               device path up automatically (default_bfe.cpp:56-57, member_bfe.cpp:40-45).
//
// Compiles against the real pagmo headers (or against oracle/shim in this repository's tests).  There is no CPU
// fallback: a problem without a device evaluator makes cuda_bfe throw std::invalid_argument, a CUDA failure throws
// std::runtime_error - the same exception types pagmo_throw produces (reference exceptions.hpp:112-126), so
// island::wait_check() semantics are unchanged.
#ifndef PAGMO_CUDA_CUDA_BFE_HPP
#define PAGMO_CUDA_CUDA_BFE_HPP

#include <exception>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

#include <pagmo/bfe.hpp>
#include <pagmo/exceptions.hpp>
#include <pagmo/problem.hpp>
#include <pagmo/problems/ackley.hpp>
#include <pagmo/problems/dtlz.hpp>
#include <pagmo/problems/griewank.hpp>
#include <pagmo/problems/hock_schittkowski_71.hpp>
#include <pagmo/problems/lennard_jones.hpp>
#include <pagmo/problems/luksan_vlcek1.hpp>
#include <pagmo/problems/rastrigin.hpp>
#include <pagmo/problems/rosenbrock.hpp>
#include <pagmo/problems/schwefel.hpp>
#include <pagmo/problems/zdt.hpp>
#include <pagmo/s11n.hpp>
#include <pagmo/threading.hpp>
#include <pagmo/types.hpp>

#include <pagmo_cuda/pgc.h>

namespace pagmo_cuda
{

namespace detail
{

[[noreturn]] inline void throw_status(int rc, const char *where)
{
    const std::string msg = std::string(where) + ": " + pgc_last_error();
    if (rc == PGC_ERR_INVALID_ARGUMENT || rc == PGC_ERR_UNSUPPORTED) {
        pagmo_throw(std::invalid_argument, msg);
    }
    pagmo_throw(std::runtime_error, msg);
}

inline void check(int rc, const char *where)
{
    if (rc != PGC_OK) throw_status(rc, where);
}

// One pgc_ctx per (process, device), shared by every adapter object that names the device; destroyed at exit.
inline std::shared_ptr<pgc_ctx> device_context(int device)
{
    static std::mutex mtx;
    static std::map<int, std::shared_ptr<pgc_ctx>> table;
    std::lock_guard<std::mutex> lk(mtx);
    auto it = table.find(device);
    if (it != table.end()) return it->second;
    pgc_ctx *raw = nullptr;
    check(pgc_ctx_create(device, &raw), "pgc_ctx_create");
    std::shared_ptr<pgc_ctx> sp(raw, [](pgc_ctx *c) { pgc_ctx_destroy(c); });
    table.emplace(device, sp);
    return sp;
}

// One mutex per (process, device): a pgc_ctx owns one stream and one staging ring, so every call that goes through the shared
// context of a device is serialised - also between DIFFERENT problems living on that device (e.g. two islands of an archipelago
// evaluating on device 0 from their own threads).
inline std::mutex &device_mutex(int device)
{
    static std::mutex mtx;
    static std::map<int, std::unique_ptr<std::mutex>> table;
    std::lock_guard<std::mutex> lk(mtx);
    auto &slot = table[device];
    if (!slot) slot.reset(new std::mutex);
    return *slot;
}

// RAII handle on a device-side problem.  Copies of an adapter share it; calls are serialised by the device's mutex because
// one pgc_ctx drives one stream and one staging ring (thread_safety::basic, reference threading.hpp:42).
class problem_handle
{
public:
    // ctx = nullptr: the device's shared context (calls serialised by device_mutex); a private context (own stream, see
    // private_on) belongs to one owner - e.g. one island of a cuda_archipelago - which then calls raw() from its own thread only
    problem_handle(int device, const pgc_problem_desc &desc, std::shared_ptr<pgc_ctx> ctx = nullptr)
        : m_ctx(ctx ? std::move(ctx) : device_context(device)), m_device(device), m_desc(desc)
    {
        // the handle keeps its own copy of the description (tables included), so that it can be re-created on another device
        if (desc.rotation) m_rotation.assign(desc.rotation, desc.rotation + desc.rotation_len);
        if (desc.shift) m_shift.assign(desc.shift, desc.shift + desc.shift_len);
        if (desc.shuffle) m_shuffle.assign(desc.shuffle, desc.shuffle + desc.shuffle_len);
        m_desc.rotation = m_rotation.empty() ? nullptr : m_rotation.data();
        m_desc.shift = m_shift.empty() ? nullptr : m_shift.data();
        m_desc.shuffle = m_shuffle.empty() ? nullptr : m_shuffle.data();
        check(pgc_problem_create(m_ctx.get(), &m_desc, &m_prob), "pgc_problem_create");
        read_sizes();
    }
    // meta-problem over another device problem (pgc_problem_translate / pgc_problem_decompose / pgc_problem_unconstrain): kind =
    // PGC_TRANSLATE with a = translation, PGC_DECOMPOSE with a = weight, b = reference point, or PGC_UNCONSTRAIN with a = weights,
    // b = the constraint tolerances the wrapper is built with (empty: the inner problem's).  The wrapper keeps the inner handle alive.
    problem_handle(std::shared_ptr<problem_handle> inner, int kind, pagmo::vector_double a, pagmo::vector_double b, int method)
        : m_ctx(inner->m_ctx), m_device(inner->m_device), m_inner(std::move(inner)), m_meta_kind(kind), m_meta_a(std::move(a)),
          m_meta_b(std::move(b)), m_meta_method(method)
    {
        std::lock_guard<std::mutex> lk(device_mutex(m_device));
        if (kind == PGC_TRANSLATE) {
            check(pgc_problem_translate(m_inner->m_prob, m_meta_a.data(), m_meta_a.size(), &m_prob), "pgc_problem_translate");
        } else if (kind == PGC_UNCONSTRAIN) {
            // unconstrain copies its inner problem, tolerances included (unconstrain.cpp:155-162 reads m_problem.get_c_tol()): the
            // device wrapper takes them at creation, so the shared inner handle gets them for the call only
            pagmo::vector_double saved(m_inner->m_nec + m_inner->m_nic);
            if (!m_meta_b.empty()) {
                check(pgc_problem_c_tol(m_inner->m_prob, saved.data()), "pgc_problem_c_tol");
                check(pgc_problem_set_c_tol(m_inner->m_prob, m_meta_b.data(), m_meta_b.size()), "pgc_problem_set_c_tol");
            }
            const int rc = pgc_problem_unconstrain(m_inner->m_prob, method, m_meta_a.data(), m_meta_a.size(), &m_prob);
            if (!m_meta_b.empty()) pgc_problem_set_c_tol(m_inner->m_prob, saved.data(), saved.size());
            check(rc, "pgc_problem_unconstrain");
        } else {
            if (m_meta_b.size() != m_meta_a.size()) {
                pagmo_throw(std::invalid_argument,
                            "Reference point size must be equal to the number of objectives. The size of the reference point is "
                                + std::to_string(m_meta_b.size()) + " while the problem has " + std::to_string(m_inner->nf())
                                + " objectives");
            }
            check(pgc_problem_decompose(m_inner->m_prob, m_meta_a.data(), m_meta_b.data(), m_meta_a.size(), method, 0, &m_prob),
                  "pgc_problem_decompose");
        }
        read_sizes();
    }
    ~problem_handle()
    {
        pgc_problem_destroy(m_prob); // a wrapper first, then (member order) its inner handle
    }
    problem_handle(const problem_handle &) = delete;
    problem_handle &operator=(const problem_handle &) = delete;

    std::size_t nx() const { return m_nx; }
    std::size_t nf() const { return m_nf; }     // problem::get_nf: the width of a fitness row
    std::size_t nobj() const { return m_nf - m_nec - m_nic; }
    std::size_t nec() const { return m_nec; }
    std::size_t nic() const { return m_nic; }
    int device() const { return m_device; }
    pgc_problem *raw() const { return m_prob; }
    pgc_ctx *context() const { return m_ctx.get(); }

    // the same problem on `device` behind a NEW private context (own stream): islands that share a GPU overlap on it like the
    // reference's island threads share a CPU
    std::shared_ptr<problem_handle> private_on(int device) const
    {
        pgc_ctx *raw_ctx = nullptr;
        check(pgc_ctx_create(device, &raw_ctx), "pgc_ctx_create");
        std::shared_ptr<pgc_ctx> ctx(raw_ctx, [](pgc_ctx *c) { pgc_ctx_destroy(c); });
        return private_with(device, ctx);
    }

    // the same problem on another device (created on first use, then shared)
    std::shared_ptr<problem_handle> twin_on(int device) const
    {
        std::lock_guard<std::mutex> lk(m_twin_mtx);
        auto &slot = m_twins[device];
        if (!slot) {
            slot = m_inner ? std::make_shared<problem_handle>(m_inner->twin_on(device), m_meta_kind, m_meta_a, m_meta_b, m_meta_method)
                           : std::make_shared<problem_handle>(device, m_desc);
        }
        return slot;
    }

    // Serialisation of the description this handle was built from - the UDP's constructor arguments, tables included, and for a
    // meta-problem its own arguments followed by the wrapped problem's - so that a CUDA UDP loaded into a default-constructed
    // object (pagmo::problem / island / archipelago save + load, pygmo pickling) re-creates the same device problem.
    template <typename Archive>
    static void save_chain(Archive &ar, const problem_handle *h)
    {
        int kind = !h ? -1 : (h->m_inner ? h->m_meta_kind : 0);
        ar &kind;
        if (kind < 0) return;
        int device = h->m_device;
        ar &device;
        if (kind != 0) {
            pagmo::vector_double a = h->m_meta_a, b = h->m_meta_b;
            int method = h->m_meta_method;
            ar &a;
            ar &b;
            ar &method;
            save_chain(ar, h->m_inner.get());
            return;
        }
        int family = h->m_desc.family;
        unsigned prob_id = h->m_desc.prob_id, dim = h->m_desc.dim, nobj = h->m_desc.nobj, param = h->m_desc.param;
        std::vector<double> rotation = h->m_rotation, shift = h->m_shift;
        std::vector<int32_t> shuffle = h->m_shuffle;
        ar &family;
        ar &prob_id;
        ar &dim;
        ar &nobj;
        ar &param;
        ar &rotation;
        ar &shift;
        ar &shuffle;
    }
    template <typename Archive>
    static std::shared_ptr<problem_handle> load_chain(Archive &ar)
    {
        int kind = -1, device = 0;
        ar &kind;
        if (kind < 0) return nullptr;
        ar &device;
        if (kind != 0) {
            pagmo::vector_double a, b;
            int method = 0;
            ar &a;
            ar &b;
            ar &method;
            auto inner = load_chain(ar);
            if (!inner) pagmo_throw(std::invalid_argument, "cuda meta-problem archive without an inner problem");
            return std::make_shared<problem_handle>(std::move(inner), kind, std::move(a), std::move(b), method);
        }
        pgc_problem_desc d{};
        std::vector<double> rotation, shift;
        std::vector<int32_t> shuffle;
        ar &d.family;
        ar &d.prob_id;
        ar &d.dim;
        ar &d.nobj;
        ar &d.param;
        ar &rotation;
        ar &shift;
        ar &shuffle;
        d.rotation = rotation.empty() ? nullptr : rotation.data();
        d.rotation_len = rotation.size();
        d.shift = shift.empty() ? nullptr : shift.data();
        d.shift_len = shift.size();
        d.shuffle = shuffle.empty() ? nullptr : shuffle.data();
        d.shuffle_len = shuffle.size();
        return std::make_shared<problem_handle>(device, d); // the handle copies the tables
    }

    // n decision vectors at dvs -> n fitness vectors at fvs (plain pointers: used for the shards of a multi-device batch)
    void evaluate_raw(const double *dvs, std::size_t n, double *fvs) const
    {
        std::lock_guard<std::mutex> lk(device_mutex(m_device));
        check(pgc_eval_host(m_prob, dvs, n, fvs), "pgc_eval_host");
    }

    pagmo::vector_double evaluate(const pagmo::vector_double &dvs) const
    {
        if (dvs.size() % m_nx != 0u) {
            pagmo_throw(std::invalid_argument, "cuda evaluator: a batch of " + std::to_string(dvs.size())
                                                   + " values is not a multiple of the problem dimension "
                                                   + std::to_string(m_nx));
        }
        const std::size_t n = dvs.size() / m_nx;
        pagmo::vector_double fvs(n * m_nf);
        std::lock_guard<std::mutex> lk(device_mutex(m_device));
        check(pgc_eval_host(m_prob, dvs.data(), n, fvs.data()), "pgc_eval_host");
        return fvs;
    }

    std::pair<pagmo::vector_double, pagmo::vector_double> bounds() const
    {
        pagmo::vector_double lb(m_nx), ub(m_nx);
        check(pgc_problem_bounds(m_prob, lb.data(), ub.data()), "pgc_problem_bounds");
        return {std::move(lb), std::move(ub)};
    }

    std::string name() const
    {
        char buf[256];
        check(pgc_problem_name(m_prob, buf, sizeof(buf)), "pgc_problem_name");
        return buf;
    }

    // algorithm::evolve on a host population: upload x [n x nx] / f [n x nf], run on the device, download in place.
    // Returns the number of generations run.
    unsigned evolve(const pgc_algo_desc &algo, pagmo::vector_double &x, pagmo::vector_double &f, unsigned first_generation) const
    {
        unsigned done = 0;
        on_device(x, f, "pgc_algo_evolve_device", [&](double *dx, double *df, std::size_t n) {
            return pgc_algo_evolve_device(m_prob, &algo, dx, df, n, first_generation, &done, nullptr);
        });
        return done;
    }

    // the same for a UDA built with memory = true: `state` (four host arrays, pgc_algo_memory's a / b / c / u) travels with the
    // population; initialized = false on the algorithm's first call with this population size
    struct algo_state {
        pagmo::vector_double a, b, c;
        std::vector<uint32_t> u;
        pagmo::vector_double es; // cmaes / xnes: their host-side state (pgc_es_state_len doubles)
        bool initialized = false;
    };
    unsigned evolve_memory(const pgc_algo_desc &algo, pagmo::vector_double &x, pagmo::vector_double &f, unsigned first_generation,
                           algo_state &st) const
    {
        return evolve_full(algo, x, f, first_generation, &st, 0u, nullptr, nullptr);
    }
    // evolve() in full: optional state (memory = true), optional log (verbosity > 0: *log = the reference's log lines as rows of
    // *row_len doubles, pgc_algo_evolve_logged_device)
    unsigned evolve_full(const pgc_algo_desc &algo, pagmo::vector_double &x, pagmo::vector_double &f, unsigned first_generation, algo_state *st,
                         unsigned verbosity, pagmo::vector_double *log, std::size_t *row_len) const
    {
        const std::size_t n = x.size() / m_nx;
        if (st && (!st->initialized || st->u.size() != n)) { // sade.cpp:137: a population of another size restarts the adaptation
            st->a.assign(n * m_nx, 0.), st->b.assign(n * m_nx, 0.), st->c.assign(n * m_nf, 0.), st->u.assign(n, 0u);
            // (xnes keeps its distribution whatever the population size, xnes.cpp:163; cmaes' state checks the size itself)
            if (algo.algo != PGC_ALGO_XNES && algo.algo != PGC_ALGO_CMAES) st->initialized = false;
        }
        std::size_t max_rows = 0, rl = 0;
        if (verbosity) {
            check(pgc_algo_log_row_len(m_prob, algo.algo, &rl), "pgc_algo_log_row_len");
            max_rows = algo.gens ? (algo.gens - 1u) / verbosity + 1u : 0u;
            log->assign(max_rows * rl, 0.);
            *row_len = rl;
        }
        unsigned done = 0;
        on_device(x, f, "pgc_algo_evolve_logged_device", [&](double *dx, double *df, std::size_t) {
            // (device_mutex is held by on_device)
            void *d[4] = {nullptr, nullptr, nullptr, nullptr};
            void *h[4] = {nullptr, nullptr, nullptr, nullptr};
            std::size_t bytes[4] = {0, 0, 0, 0};
            if (st) {
                h[0] = st->a.data(), h[1] = st->b.data(), h[2] = st->c.data(), h[3] = st->u.data();
                bytes[0] = st->a.size() * sizeof(double), bytes[1] = st->b.size() * sizeof(double), bytes[2] = st->c.size() * sizeof(double);
                bytes[3] = st->u.size() * sizeof(uint32_t);
            }
            int rc = PGC_OK;
            for (int k = 0; st && k < 4 && rc == PGC_OK; ++k) {
                rc = pgc_malloc_device(m_ctx.get(), bytes[k] ? bytes[k] : 8u, &d[k]);
                if (rc == PGC_OK && st->initialized && bytes[k]) rc = pgc_memcpy_h2d(m_ctx.get(), d[k], h[k], bytes[k]);
            }
            pgc_algo_memory mem{static_cast<double *>(d[0]), static_cast<double *>(d[1]), static_cast<double *>(d[2]),
                                static_cast<uint32_t *>(d[3]), st && st->initialized ? 1 : 0, 0, nullptr, 0};
            if (st && (algo.algo == PGC_ALGO_CMAES || algo.algo == PGC_ALGO_XNES)) {
                std::size_t len = 0;
                if (rc == PGC_OK) rc = pgc_es_state_len(algo.algo, m_nx, &len);
                if (st->es.size() != len) st->es.assign(len, 0.);
                mem.h_state = st->es.data();
                mem.h_state_len = len;
            }
            std::size_t n_rows = 0;
            if (rc == PGC_OK) {
                rc = pgc_algo_evolve_logged_device(m_prob, &algo, dx, df, n, first_generation, &done, st ? &mem : nullptr, verbosity,
                                                   verbosity ? log->data() : nullptr, max_rows, &n_rows, nullptr);
            }
            for (int k = 0; st && k < 4 && rc == PGC_OK; ++k)
                if (bytes[k]) rc = pgc_memcpy_d2h(m_ctx.get(), h[k], d[k], bytes[k]);
            for (void *p : d)
                if (p) pgc_free_device(m_ctx.get(), p);
            if (rc == PGC_OK && st) st->initialized = true;
            if (rc == PGC_OK && verbosity) log->resize(n_rows * rl);
            return rc;
        });
        return done;
    }

    // upload x [n x nx] / f [n x nf], run `call(d_x, d_f, n)` (a pgc_*_evolve_device entry point on this handle's problem: see raw()),
    // download both in place; throws what the status says
    template <typename Call>
    void on_device(pagmo::vector_double &x, pagmo::vector_double &f, const char *what, Call call) const
    {
        const std::size_t n = x.size() / m_nx;
        std::lock_guard<std::mutex> lk(device_mutex(m_device));
        void *dx = nullptr, *df = nullptr;
        check(pgc_malloc_device(m_ctx.get(), x.size() * sizeof(double), &dx), "pgc_malloc_device");
        if (int rc = pgc_malloc_device(m_ctx.get(), f.size() * sizeof(double), &df)) {
            pgc_free_device(m_ctx.get(), dx);
            throw_status(rc, "pgc_malloc_device");
        }
        int rc = pgc_memcpy_h2d(m_ctx.get(), dx, x.data(), x.size() * sizeof(double));
        if (rc == PGC_OK) rc = pgc_memcpy_h2d(m_ctx.get(), df, f.data(), f.size() * sizeof(double));
        const char *where = "pgc_memcpy_h2d";
        if (rc == PGC_OK) {
            rc = call(static_cast<double *>(dx), static_cast<double *>(df), n);
            where = what;
        }
        if (rc == PGC_OK) {
            rc = pgc_memcpy_d2h(m_ctx.get(), x.data(), dx, x.size() * sizeof(double));
            if (rc == PGC_OK) rc = pgc_memcpy_d2h(m_ctx.get(), f.data(), df, f.size() * sizeof(double));
            where = "pgc_memcpy_d2h";
        }
        pgc_free_device(m_ctx.get(), dx);
        pgc_free_device(m_ctx.get(), df);
        if (rc != PGC_OK) throw_status(rc, where);
    }

private:
    void read_sizes()
    {
        check(pgc_problem_nx(m_prob, &m_nx), "pgc_problem_nx");
        check(pgc_problem_nf(m_prob, &m_nf), "pgc_problem_nf");
        check(pgc_problem_nec(m_prob, &m_nec), "pgc_problem_nec");
        check(pgc_problem_nic(m_prob, &m_nic), "pgc_problem_nic");
    }
    std::shared_ptr<problem_handle> private_with(int device, const std::shared_ptr<pgc_ctx> &ctx) const
    {
        if (!m_inner) return std::make_shared<problem_handle>(device, m_desc, ctx);
        return std::make_shared<problem_handle>(m_inner->private_with(device, ctx), m_meta_kind, m_meta_a, m_meta_b, m_meta_method);
    }
    std::shared_ptr<pgc_ctx> m_ctx;
    int m_device = 0;
    pgc_problem_desc m_desc{};
    std::vector<double> m_rotation, m_shift;
    std::vector<int32_t> m_shuffle;
    std::shared_ptr<problem_handle> m_inner; // meta-problems only
    int m_meta_kind = 0;
    pagmo::vector_double m_meta_a, m_meta_b;
    int m_meta_method = 0;
    pgc_problem *m_prob = nullptr;
    std::size_t m_nx = 0, m_nf = 0, m_nec = 0, m_nic = 0;
    mutable std::mutex m_twin_mtx;
    mutable std::map<int, std::shared_ptr<problem_handle>> m_twins;
};

inline pgc_problem_desc make_desc(int family, unsigned prob_id, unsigned dim, unsigned nobj = 0, unsigned param = 0)
{
    pgc_problem_desc d{};
    d.family = family;
    d.prob_id = prob_id;
    d.dim = dim;
    d.nobj = nobj;
    d.param = param;
    return d;
}

} // namespace detail

// Common part of the CUDA-backed UDPs: the five members pagmo looks for on a UDP.
class cuda_udp_base
{
public:
    pagmo::vector_double fitness(const pagmo::vector_double &x) const // a batch of one
    {
        return handle().evaluate(x);
    }
    pagmo::vector_double batch_fitness(const pagmo::vector_double &dvs) const
    {
        return handle().evaluate(dvs);
    }
    std::pair<pagmo::vector_double, pagmo::vector_double> get_bounds() const
    {
        return handle().bounds();
    }
    pagmo::vector_double::size_type get_nobj() const
    {
        return handle().nobj();
    }
    pagmo::vector_double::size_type get_nec() const // 0 but for the constrained UDPs (hock_schittkowski_71, luksan_vlcek1)
    {
        return handle().nec();
    }
    pagmo::vector_double::size_type get_nic() const
    {
        return handle().nic();
    }
    std::string get_name() const
    {
        return handle().name() + " [CUDA sm_100a]";
    }
    pagmo::thread_safety get_thread_safety() const
    {
        return pagmo::thread_safety::basic;
    }
    int device() const { return m_device; }
    std::shared_ptr<detail::problem_handle> shared_handle() const { return m_handle; }

protected:
    const detail::problem_handle &handle() const
    {
        if (!m_handle) pagmo_throw(std::runtime_error, "cuda UDP used before its device problem was created");
        return *m_handle;
    }
    // the device problem's own description goes into the archive; a load re-creates the device problem from it (Boost archives and
    // the test archive say which way they go through `is_loading`)
    template <typename Archive>
    void serialize_handle(Archive &ar)
    {
        ar &m_device;
        if (Archive::is_loading::value) {
            m_handle = detail::problem_handle::load_chain(ar);
        } else {
            detail::problem_handle::save_chain(ar, m_handle.get());
        }
    }
    int m_device = 0;
    std::shared_ptr<detail::problem_handle> m_handle;
};

// CEC2014 on the device.  The reference UDP keeps its tables private (cec2014.hpp:227-236), so the CUDA UDP takes
// them as constructor arguments in the layout of the reference members: `rotation` = m_rotation_matrix
// (component i at i*dim*dim), `shift` = m_origin_shift after compaction (component i at i*dim, cec2014.cpp:76-86),
// `shuffle` = m_shuffle (1-based).
class cuda_cec2014 : public cuda_udp_base
{
public:
    cuda_cec2014() = default; // pagmo requires default-constructible UDPs; unusable until assigned
    cuda_cec2014(unsigned prob_id, unsigned dim, std::vector<double> rotation, std::vector<double> shift,
                 std::vector<int> shuffle = {}, int device = 0)
        : m_prob_id(prob_id), m_dim(dim), m_rotation(std::move(rotation)), m_shift(std::move(shift)),
          m_shuffle(std::move(shuffle))
    {
        m_device = device;
        create();
    }
    const pagmo::vector_double &get_origin_shift() const // cec2014.hpp:104
    {
        return m_shift;
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_prob_id, m_dim, m_rotation, m_shift, m_shuffle, m_device);
        serialize_handle(ar);
    }

private:
    void create()
    {
        pgc_problem_desc d = detail::make_desc(PGC_CEC2014, m_prob_id, m_dim);
        d.rotation = m_rotation.data();
        d.rotation_len = m_rotation.size();
        d.shift = m_shift.data();
        d.shift_len = m_shift.size();
        static_assert(sizeof(int) == sizeof(int32_t), "pgc.h passes the shuffle as int32_t");
        d.shuffle = reinterpret_cast<const int32_t *>(m_shuffle.data());
        d.shuffle_len = m_shuffle.size();
        m_handle = std::make_shared<detail::problem_handle>(m_device, d);
    }
    unsigned m_prob_id = 0, m_dim = 0;
    std::vector<double> m_rotation, m_shift;
    std::vector<int> m_shuffle;
};

// rastrigin / ackley / griewank / schwefel / rosenbrock on the device (same constructor argument as the
// reference UDPs: the dimension).
template <int Family>
class cuda_simple : public cuda_udp_base
{
public:
    explicit cuda_simple(unsigned dim = (Family == PGC_ROSENBROCK ? 2u : 1u), int device = 0) : m_dim(dim)
    {
        m_device = device;
        m_handle = std::make_shared<detail::problem_handle>(m_device, detail::make_desc(Family, 0u, dim));
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_dim, m_device);
        serialize_handle(ar);
    }

private:
    unsigned m_dim;
};
using cuda_rastrigin = cuda_simple<PGC_RASTRIGIN>;
using cuda_ackley = cuda_simple<PGC_ACKLEY>;
using cuda_griewank = cuda_simple<PGC_GRIEWANK>;
using cuda_schwefel = cuda_simple<PGC_SCHWEFEL>;
using cuda_rosenbrock = cuda_simple<PGC_ROSENBROCK>;

// ZDT1-6 on the device; same constructor arguments as pagmo::zdt (zdt.hpp:145).
class cuda_zdt : public cuda_udp_base
{
public:
    explicit cuda_zdt(unsigned prob_id = 1u, unsigned param = 30u, int device = 0) : m_prob_id(prob_id), m_param(param)
    {
        m_device = device;
        m_handle = std::make_shared<detail::problem_handle>(m_device, detail::make_desc(PGC_ZDT, prob_id, param));
    }
    pagmo::vector_double::size_type get_nix() const // zdt.cpp:126-143: zdt5 is integer valued
    {
        return m_prob_id == 5u ? 30u + 5u * (m_param - 1u) : 0u;
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_prob_id, m_param, m_device);
        serialize_handle(ar);
    }

private:
    unsigned m_prob_id, m_param;
};

// DTLZ1-7 on the device; same constructor arguments as pagmo::dtlz (dtlz.hpp:105).
class cuda_dtlz : public cuda_udp_base
{
public:
    explicit cuda_dtlz(unsigned prob_id = 1u, pagmo::vector_double::size_type dim = 5u,
                       pagmo::vector_double::size_type fdim = 3u, unsigned alpha = 100u, int device = 0)
        : m_prob_id(prob_id), m_alpha(alpha), m_dim(static_cast<unsigned>(dim)), m_fdim(static_cast<unsigned>(fdim))
    {
        m_device = device;
        m_handle = std::make_shared<detail::problem_handle>(m_device,
                                                            detail::make_desc(PGC_DTLZ, prob_id, m_dim, m_fdim, alpha));
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_prob_id, m_dim, m_fdim, m_alpha, m_device);
        serialize_handle(ar);
    }

private:
    unsigned m_prob_id, m_alpha, m_dim, m_fdim;
};

// CEC2013 on the device.  Like cuda_cec2014 the data tables are constructor arguments, in the layout of the reference members:
// `rotation` = m_rotation_matrix = MD[dim] (10 matrices, cec2013.cpp:65-67), `shift` = m_origin_shift (flat, component i at i*dim).
class cuda_cec2013 : public cuda_udp_base
{
public:
    cuda_cec2013() = default;
    cuda_cec2013(unsigned prob_id, unsigned dim, std::vector<double> rotation, std::vector<double> shift, int device = 0)
        : m_prob_id(prob_id), m_dim(dim), m_rotation(std::move(rotation)), m_shift(std::move(shift))
    {
        m_device = device;
        pgc_problem_desc d = detail::make_desc(PGC_CEC2013, m_prob_id, m_dim);
        d.rotation = m_rotation.data();
        d.rotation_len = m_rotation.size();
        d.shift = m_shift.data();
        d.shift_len = m_shift.size();
        m_handle = std::make_shared<detail::problem_handle>(m_device, d);
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_prob_id, m_dim, m_rotation, m_shift, m_device);
        serialize_handle(ar);
    }

private:
    unsigned m_prob_id = 0, m_dim = 0;
    std::vector<double> m_rotation, m_shift;
};

// WFG1-9 on the device; same constructor arguments as pagmo::wfg (wfg.hpp:107).
class cuda_wfg : public cuda_udp_base
{
public:
    explicit cuda_wfg(unsigned prob_id = 1u, pagmo::vector_double::size_type dim_dvs = 5u, pagmo::vector_double::size_type dim_obj = 3u,
                      pagmo::vector_double::size_type dim_k = 4u, int device = 0)
        : m_prob_id(prob_id), m_dim_dvs(static_cast<unsigned>(dim_dvs)), m_dim_obj(static_cast<unsigned>(dim_obj)),
          m_dim_k(static_cast<unsigned>(dim_k))
    {
        m_device = device;
        m_handle = std::make_shared<detail::problem_handle>(m_device, detail::make_desc(PGC_WFG, prob_id, m_dim_dvs, m_dim_obj, m_dim_k));
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_prob_id, m_dim_dvs, m_dim_obj, m_dim_k, m_device);
        serialize_handle(ar);
    }

private:
    unsigned m_prob_id, m_dim_dvs, m_dim_obj, m_dim_k;
};

// Lennard-Jones cluster on the device; same constructor argument as pagmo::lennard_jones (lennard_jones.hpp:82).
class cuda_lennard_jones : public cuda_udp_base
{
public:
    explicit cuda_lennard_jones(unsigned atoms = 3u, int device = 0) : m_atoms(atoms)
    {
        m_device = device;
        m_handle = std::make_shared<detail::problem_handle>(m_device, detail::make_desc(PGC_LENNARD_JONES, 0u, atoms));
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_atoms, m_device);
        serialize_handle(ar);
    }

private:
    unsigned m_atoms;
};

// The two constrained UDPs with a device evaluator (fitness rows [objective | equalities | inequalities]); same constructor
// arguments as pagmo::hock_schittkowski_71 (none) and pagmo::luksan_vlcek1 (luksan_vlcek1.hpp:78: the dimension, >= 3).
class cuda_hock_schittkowski_71 : public cuda_udp_base
{
public:
    explicit cuda_hock_schittkowski_71(int device = 0)
    {
        m_device = device;
        m_handle = std::make_shared<detail::problem_handle>(m_device, detail::make_desc(PGC_HOCK_SCHITTKOWSKI_71, 0u, 4u));
    }
    pagmo::vector_double best_known() const // hock_schittkowski_71.cpp:141-144
    {
        return {1., 4.74299963, 3.82114998, 1.37940829};
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_device);
        serialize_handle(ar);
    }
};

class cuda_luksan_vlcek1 : public cuda_udp_base
{
public:
    explicit cuda_luksan_vlcek1(unsigned dim = 3u, int device = 0) : m_dim(dim)
    {
        m_device = device;
        m_handle = std::make_shared<detail::problem_handle>(m_device, detail::make_desc(PGC_LUKSAN_VLCEK1, 0u, dim));
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_dim, m_device);
        serialize_handle(ar);
    }

private:
    unsigned m_dim;
};

// pagmo::unconstrain on the device (reference include/pagmo/problems/unconstrain.hpp, src/problems/unconstrain.cpp:66-97,136-223):
// the inner problem is one of the constrained CUDA UDPs above (or a cuda_translate of one); same constructor arguments, checks
// and messages as the reference.  The reference's unconstrain wraps a pagmo::problem and reads its tolerances (problem::get_c_tol);
// a UDP has none, so they are the optional fourth argument (empty: all zero, problem's default).  The stock
// pagmo::unconstrain{cuda_udp, ...} works too (it penalizes on the host after the inner batch_fitness); this one keeps the penalty
// on the device, so the device generation loops run on it.
class cuda_unconstrain : public cuda_udp_base
{
public:
    cuda_unconstrain() = default;
    template <typename T>
    explicit cuda_unconstrain(const T &inner, const std::string &method = "death penalty",
                              const pagmo::vector_double &weights = pagmo::vector_double(),
                              const pagmo::vector_double &c_tol = pagmo::vector_double())
        : m_method(method), m_weights(weights)
    {
        int code = -1;
        if (method == "death penalty") code = PGC_UNCONSTRAIN_DEATH;
        else if (method == "kuri") code = PGC_UNCONSTRAIN_KURI;
        else if (method == "weighted") code = PGC_UNCONSTRAIN_WEIGHTED;
        else if (method == "ignore_c") code = PGC_UNCONSTRAIN_IGNORE_C;
        else if (method == "ignore_o") code = PGC_UNCONSTRAIN_IGNORE_O;
        const auto inner_handle = inner.shared_handle();
        if (inner_handle->nec() + inner_handle->nic() != 0u && code < 0) { // unconstrain.cpp:84-88 (an unconstrained inner fails first)
            pagmo_throw(std::invalid_argument, "The method " + method + " is not supported (did you misspell?)");
        }
        m_device = inner.device();
        m_handle = std::make_shared<detail::problem_handle>(inner_handle, static_cast<int>(PGC_UNCONSTRAIN), weights, c_tol,
                                                            code < 0 ? static_cast<int>(PGC_UNCONSTRAIN_DEATH) : code);
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_method, m_weights);
        serialize_handle(ar);
    }

private:
    std::string m_method;
    pagmo::vector_double m_weights;
};

// pagmo::translate on the device (reference include/pagmo/problems/translate.hpp, src/problems/translate.cpp:100-153): the
// inner problem is one of the CUDA UDPs above (or another cuda_translate); fitness(x) = inner.fitness(x - translation), bounds
// = inner bounds + translation.  The stock pagmo::translate{cuda_udp, t} works too (it de-shifts on the host and calls the inner
// batch_fitness); this one keeps the de-shifting on the device, so a device-resident generation loop never leaves it.
class cuda_translate : public cuda_udp_base
{
public:
    cuda_translate() = default; // like every UDP: default-constructible; unusable until assigned from a constructed one
    template <typename T>
    cuda_translate(const T &inner, const pagmo::vector_double &translation) : m_translation(translation)
    {
        m_device = inner.device();
        m_handle = std::make_shared<detail::problem_handle>(inner.shared_handle(), static_cast<int>(PGC_TRANSLATE), translation,
                                                            pagmo::vector_double{}, 0);
    }
    const pagmo::vector_double &get_translation() const // translate.hpp: get_translation()
    {
        return m_translation;
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_translation);
        serialize_handle(ar);
    }

private:
    pagmo::vector_double m_translation;
};

// pagmo::decompose on the device (reference include/pagmo/problems/decompose.hpp, src/problems/decompose.cpp:66-154;
// decompose_objectives, src/utils/multi_objective.cpp:582-638).  Same constructor arguments and checks as the reference;
// adapt_ideal = true is refused: the reference moves z after every single fitness call, in call order.
class cuda_decompose : public cuda_udp_base
{
public:
    cuda_decompose() = default;
    template <typename T>
    cuda_decompose(const T &inner, const pagmo::vector_double &weight, const pagmo::vector_double &z,
                   const std::string &method = "weighted", bool adapt_ideal = false)
        : m_weight(weight), m_z(z), m_method(method)
    {
        if (method != "weighted" && method != "tchebycheff" && method != "bi") { // decompose.cpp:80-83
            pagmo_throw(std::invalid_argument, "Decomposition method requested is: " + method
                                                   + " while only one of ['weighted', 'tchebycheff', 'bi'] are allowed");
        }
        if (adapt_ideal) {
            pagmo_throw(std::invalid_argument, "cuda_decompose: adapt_ideal is a sequential semantic (decompose.cpp:143-149: z moves "
                                               "after every fitness call); it is not available on the batch path");
        }
        m_device = inner.device();
        const int code = method == "weighted" ? PGC_DECOMPOSE_WEIGHTED : method == "tchebycheff" ? PGC_DECOMPOSE_TCHEBYCHEFF : PGC_DECOMPOSE_BI;
        m_handle = std::make_shared<detail::problem_handle>(inner.shared_handle(), static_cast<int>(PGC_DECOMPOSE), weight, z, code);
    }
    pagmo::vector_double get_z() const // decompose.cpp:208-211
    {
        return m_z;
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_weight, m_z, m_method);
        serialize_handle(ar);
    }

private:
    pagmo::vector_double m_weight, m_z;
    std::string m_method;
};

namespace detail
{
// The device problem behind a pagmo::problem: (1) a CUDA-backed UDP -> its own handle; (2) a stock pagmo UDP whose parameters
// are recoverable from its public interface -> a cached device twin; (3) anything else -> nullptr (callers throw: no CPU fallback).
class twin_cache
{
public:
    std::shared_ptr<problem_handle> find(const pagmo::problem &p, int device)
    {
        // a CUDA-backed UDP carries its own handle; asked for another device, its twin there (created once, then shared)
#define PGC_OWN_UDP(T)                                                                                                                      \
    if (p.is<T>()) {                                                                                                                        \
        auto own = p.extract<T>()->shared_handle();                                                                                         \
        return own->device() == device ? own : own->twin_on(device);                                                                        \
    }
        PGC_OWN_UDP(cuda_cec2014)
        PGC_OWN_UDP(cuda_cec2013)
        PGC_OWN_UDP(cuda_rastrigin)
        PGC_OWN_UDP(cuda_ackley)
        PGC_OWN_UDP(cuda_griewank)
        PGC_OWN_UDP(cuda_schwefel)
        PGC_OWN_UDP(cuda_rosenbrock)
        PGC_OWN_UDP(cuda_zdt)
        PGC_OWN_UDP(cuda_dtlz)
        PGC_OWN_UDP(cuda_wfg)
        PGC_OWN_UDP(cuda_lennard_jones)
        PGC_OWN_UDP(cuda_translate)
        PGC_OWN_UDP(cuda_decompose)
        PGC_OWN_UDP(cuda_hock_schittkowski_71)
        PGC_OWN_UDP(cuda_luksan_vlcek1)
        PGC_OWN_UDP(cuda_unconstrain)
#undef PGC_OWN_UDP
        // stock zdt / dtlz: the problem id is only visible through get_name() ("ZDT3", "DTLZ2": zdt.cpp:161-164,
        // dtlz.cpp:170-173); dtlz4's alpha is private and cannot be recovered; so is wfg's dim_k
        if (p.is<pagmo::zdt>() || p.is<pagmo::dtlz>()) {
            const std::string nm = p.get_name();
            const bool is_zdt = p.is<pagmo::zdt>();
            const unsigned id = static_cast<unsigned>(std::stoul(nm.substr(is_zdt ? 3 : 4)));
            const unsigned nx = static_cast<unsigned>(p.get_nx());
            if (is_zdt) {
                const unsigned param = (id == 5u) ? (nx - 30u) / 5u + 1u : nx; // zdt.cpp:115-118
                return twin(device, make_desc(PGC_ZDT, id, param));
            }
            if (id == 4u) {
                pagmo_throw(std::invalid_argument, "the CUDA path cannot evaluate a stock pagmo::dtlz with prob_id 4: its alpha "
                                                   "is private; construct a pagmo_cuda::cuda_dtlz instead");
            }
            return twin(device, make_desc(PGC_DTLZ, id, nx, static_cast<unsigned>(p.get_nobj()), 100u));
        }
        if (p.is<pagmo::lennard_jones>()) { // nx = 3*atoms - 6, lennard_jones.cpp:99-110
            return twin(device, make_desc(PGC_LENNARD_JONES, 0u, static_cast<unsigned>((p.get_nx() + 6u) / 3u)));
        }
        if (p.is<pagmo::hock_schittkowski_71>()) return twin(device, make_desc(PGC_HOCK_SCHITTKOWSKI_71, 0u, 4u));
        int family = 0; // stock UDPs that are fully described by (type, nx)
        if (p.is<pagmo::rastrigin>()) family = PGC_RASTRIGIN;
        else if (p.is<pagmo::ackley>()) family = PGC_ACKLEY;
        else if (p.is<pagmo::griewank>()) family = PGC_GRIEWANK;
        else if (p.is<pagmo::schwefel>()) family = PGC_SCHWEFEL;
        else if (p.is<pagmo::rosenbrock>()) family = PGC_ROSENBROCK;
        else if (p.is<pagmo::luksan_vlcek1>()) family = PGC_LUKSAN_VLCEK1;
        if (family == 0) return nullptr;
        return twin(device, make_desc(family, 0u, static_cast<unsigned>(p.get_nx())));
    }

private:
    std::shared_ptr<problem_handle> twin(int device, const pgc_problem_desc &d)
    {
        std::lock_guard<std::mutex> lk(m_mtx);
        auto &slot = m_twins[std::make_tuple(device, d.family, d.prob_id, d.dim, d.nobj, d.param)];
        if (!slot) slot = std::make_shared<problem_handle>(device, d);
        return slot;
    }
    std::mutex m_mtx;
    std::map<std::tuple<int, int, unsigned, unsigned, unsigned, unsigned>, std::shared_ptr<problem_handle>> m_twins;
};
} // namespace detail

// The UDBFE.  Dispatch order (detail::twin_cache::find): (1) a CUDA-backed UDP -> its own device problem; (2) a stock pagmo UDP
// whose parameters are recoverable from its public interface -> a cached device twin; (3) anything else -> throw.
class cuda_bfe
{
public:
    explicit cuda_bfe(int device = 0) : m_device(device), m_devices{device}, m_cache(std::make_shared<detail::twin_cache>()) {}
    // several devices: the batch is split into contiguous shards of individuals, one per device, evaluated concurrently (the
    // problem's tables are replicated; no communication between the devices - SURVEY 8e).  Each device has its own PCIe link, so
    // the host-vector path scales with the number of devices.
    explicit cuda_bfe(std::vector<int> devices)
        : m_device(devices.empty() ? 0 : devices.front()), m_devices(std::move(devices)), m_cache(std::make_shared<detail::twin_cache>())
    {
        if (m_devices.empty()) pagmo_throw(std::invalid_argument, "cuda_bfe: the list of devices is empty");
    }

    pagmo::vector_double operator()(const pagmo::problem &p, const pagmo::vector_double &dvs) const
    {
        // fevals are bumped by pagmo::bfe (bfe.cpp:94-107), never here (member_bfe.cpp:40-45 does the same)
        const auto h = m_cache->find(p, m_device);
        if (!h) {
            // no device evaluator and, by design, no CPU fallback
            pagmo_throw(std::invalid_argument,
                        "cuda_bfe cannot evaluate the problem '" + p.get_name()
                            + "': no CUDA evaluator exists for this UDP type (wrap it in a pagmo_cuda:: UDP, or use "
                              "thread_bfe); there is no CPU fallback");
        }
        if (m_devices.size() == 1u) return h->evaluate(dvs);
        const std::size_t nx = h->nx(), nf = h->nf();
        if (dvs.size() % nx != 0u) {
            pagmo_throw(std::invalid_argument, "cuda evaluator: a batch of " + std::to_string(dvs.size())
                                                   + " values is not a multiple of the problem dimension " + std::to_string(nx));
        }
        const std::size_t n = dvs.size() / nx, G = m_devices.size();
        pagmo::vector_double fvs(n * nf);
        std::vector<std::shared_ptr<detail::problem_handle>> hs(G);
        for (std::size_t g = 0; g < G; ++g) hs[g] = (m_devices[g] == h->device()) ? h : h->twin_on(m_devices[g]);
        std::vector<std::thread> workers;
        std::vector<std::exception_ptr> errors(G);
        for (std::size_t g = 0; g < G; ++g) {
            const std::size_t lo = n * g / G, hi = n * (g + 1) / G; // contiguous shard of individuals
            if (hi == lo) continue;
            workers.emplace_back([&, g, lo, hi]() {
                try {
                    hs[g]->evaluate_raw(dvs.data() + lo * nx, hi - lo, fvs.data() + lo * nf);
                } catch (...) {
                    errors[g] = std::current_exception();
                }
            });
        }
        for (auto &w : workers) w.join();
        for (auto &e : errors)
            if (e) std::rethrow_exception(e);
        return fvs;
    }
    std::string get_name() const
    {
        std::string devs;
        for (int d : m_devices) devs += (devs.empty() ? "" : ",") + std::to_string(d);
        return "CUDA batch fitness evaluator (sm_100a, device " + devs + ")";
    }
    pagmo::thread_safety get_thread_safety() const
    {
        return pagmo::thread_safety::basic;
    }
    template <typename Archive>
    void serialize(Archive &ar, unsigned)
    {
        pagmo::detail::archive(ar, m_device, m_devices);
    }

private:
    int m_device;
    std::vector<int> m_devices;
    std::shared_ptr<detail::twin_cache> m_cache;
};

} // namespace pagmo_cuda

PAGMO_S11N_BFE_EXPORT_KEY(pagmo_cuda::cuda_bfe)
PAGMO_S11N_PROBLEM_EXPORT_KEY(pagmo_cuda::cuda_cec2014)
PAGMO_S11N_PROBLEM_EXPORT_KEY(pagmo_cuda::cuda_rastrigin)
PAGMO_S11N_PROBLEM_EXPORT_KEY(pagmo_cuda::cuda_ackley)
PAGMO_S11N_PROBLEM_EXPORT_KEY(pagmo_cuda::cuda_griewank)
PAGMO_S11N_PROBLEM_EXPORT_KEY(pagmo_cuda::cuda_schwefel)
PAGMO_S11N_PROBLEM_EXPORT_KEY(pagmo_cuda::cuda_rosenbrock)
PAGMO_S11N_PROBLEM_EXPORT_KEY(pagmo_cuda::cuda_zdt)
PAGMO_S11N_PROBLEM_EXPORT_KEY(pagmo_cuda::cuda_dtlz)
PAGMO_S11N_PROBLEM_EXPORT_KEY(pagmo_cuda::cuda_cec2013)
PAGMO_S11N_PROBLEM_EXPORT_KEY(pagmo_cuda::cuda_wfg)
PAGMO_S11N_PROBLEM_EXPORT_KEY(pagmo_cuda::cuda_lennard_jones)

#endif
