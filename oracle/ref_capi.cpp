// oracle/ref_capi.cpp - plain-C handle API over the unmodified reference.  TEST INFRASTRUCTURE ONLY.
// See ref_capi.h.  Everything here just forwards to pagmo:: symbols compiled from /root/reference/src.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <pagmo/batch_evaluators/default_bfe.hpp>
#include <pagmo/batch_evaluators/thread_bfe.hpp>
#include <pagmo/bfe.hpp>
#include <pagmo/problem.hpp>
#include <pagmo/problems/ackley.hpp>
#include <pagmo/problems/cec2013.hpp>
#include <pagmo/problems/cec2014.hpp>
#include <pagmo/problems/decompose.hpp>
#include <pagmo/problems/dtlz.hpp>
#include <pagmo/problems/griewank.hpp>
#include <pagmo/problems/hock_schittkowski_71.hpp>
#include <pagmo/problems/lennard_jones.hpp>
#include <pagmo/problems/luksan_vlcek1.hpp>
#include <pagmo/problems/rastrigin.hpp>
#include <pagmo/problems/rosenbrock.hpp>
#include <pagmo/problems/schwefel.hpp>
#include <pagmo/problems/translate.hpp>
#include <pagmo/problems/unconstrain.hpp>
#include <pagmo/problems/wfg.hpp>
#include <pagmo/problems/zdt.hpp>
#include <pagmo/types.hpp>
#include <pagmo/utils/multi_objective.hpp>

#include "ref_capi.h"

namespace oracle_ref {
void ensure_cec2014_tables(unsigned, unsigned);
void ensure_cec2013_tables(unsigned);
const std::vector<double> &cec2014_rotation(unsigned, unsigned);
const std::vector<double> &cec2014_shift(unsigned);
const std::vector<int> &cec2014_shuffle(unsigned, unsigned);
const std::vector<double> &cec2013_md(unsigned);
const std::vector<double> &cec2013_shift();
}

struct ref_problem {
    pagmo::problem prob;
};

static thread_local std::string g_err;

template <typename F> static int guarded(F &&f)
{
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    } catch (...) {
        g_err = "unknown C++ exception";
        return 2;
    }
}

static std::vector<pagmo::vector_double> unflatten(const double *f, size_t n, size_t m)
{
    std::vector<pagmo::vector_double> out(n);
    for (size_t i = 0; i < n; ++i) out[i].assign(f + i * m, f + (i + 1) * m);
    return out;
}

struct threads_env_guard {
    explicit threads_env_guard(int n)
    {
        if (n > 0) setenv("ORACLE_TBB_THREADS", std::to_string(n).c_str(), 1);
        else unsetenv("ORACLE_TBB_THREADS");
    }
    ~threads_env_guard() { unsetenv("ORACLE_TBB_THREADS"); }
};

extern "C" void ref_set_error(const char *msg) { g_err = msg; }

extern "C" {

const char *ref_last_error(void) { return g_err.c_str(); }

int ref_problem_create(const char *family, unsigned p0, unsigned p1, unsigned p2, unsigned p3, ref_problem **out)
{
    return guarded([&] {
        const std::string fam(family);
        pagmo::problem pr;
        if (fam == "rastrigin") pr = pagmo::problem{pagmo::rastrigin{p0}};
        else if (fam == "ackley") pr = pagmo::problem{pagmo::ackley{p0}};
        else if (fam == "griewank") pr = pagmo::problem{pagmo::griewank{p0}};
        else if (fam == "schwefel") pr = pagmo::problem{pagmo::schwefel{p0}};
        else if (fam == "rosenbrock") pr = pagmo::problem{pagmo::rosenbrock{p0}};
        else if (fam == "cec2014") {
            // the reference ctor dereferences the table iterators before validating (cec2014.cpp:67-70)
            if (p0 >= 1u && p0 <= 30u && (p1 == 2u || p1 == 10u || p1 == 20u || p1 == 30u || p1 == 50u || p1 == 100u))
                oracle_ref::ensure_cec2014_tables(p0, p1);
            pr = pagmo::problem{pagmo::cec2014{p0, p1}};
        } else if (fam == "cec2013") {
            if (p1 >= 2u && p1 <= 100u) oracle_ref::ensure_cec2013_tables(p1);
            pr = pagmo::problem{pagmo::cec2013{p0, p1}};
        } else if (fam == "zdt") pr = pagmo::problem{pagmo::zdt{p0, p1}};
        else if (fam == "dtlz") pr = pagmo::problem{pagmo::dtlz{p0, p1, p2, p3}};
        else if (fam == "wfg") pr = pagmo::problem{pagmo::wfg{p0, p1, p2, p3}};
        else if (fam == "lennard_jones") pr = pagmo::problem{pagmo::lennard_jones{p0}};
        else if (fam == "hock_schittkowski_71") pr = pagmo::problem{pagmo::hock_schittkowski_71{}};
        else if (fam == "luksan_vlcek1") pr = pagmo::problem{pagmo::luksan_vlcek1{p0}};
        else throw std::invalid_argument("ref_problem_create: unknown family '" + fam + "'");
        *out = new ref_problem{std::move(pr)};
    });
}

int ref_problem_translate(const ref_problem *inner, const double *t, size_t len, ref_problem **out)
{
    return guarded([&] { *out = new ref_problem{pagmo::problem{pagmo::translate{inner->prob, pagmo::vector_double(t, t + len)}}}; });
}

int ref_problem_decompose(const ref_problem *inner, const double *w, const double *z, size_t len, const char *method, int adapt_ideal,
                          ref_problem **out)
{
    return guarded([&] {
        *out = new ref_problem{pagmo::problem{pagmo::decompose{inner->prob, pagmo::vector_double(w, w + len),
                                                               pagmo::vector_double(z, z + len), method, adapt_ideal != 0}}};
    });
}

int ref_problem_unconstrain(const ref_problem *inner, const char *method, const double *weights, size_t len, ref_problem **out)
{
    return guarded([&] {
        *out = new ref_problem{pagmo::problem{pagmo::unconstrain{inner->prob, method, pagmo::vector_double(weights, weights + len)}}};
    });
}

int ref_problem_set_c_tol(ref_problem *p, const double *tol, size_t len)
{
    return guarded([&] { p->prob.set_c_tol(pagmo::vector_double(tol, tol + len)); });
}

int ref_decompose_objectives(const double *f, const double *w, const double *z, size_t m, const char *method, double *out)
{
    return guarded([&] {
        *out = pagmo::decompose_objectives(pagmo::vector_double(f, f + m), pagmo::vector_double(w, w + m),
                                           pagmo::vector_double(z, z + m), method)[0];
    });
}

void ref_problem_destroy(ref_problem *p) { delete p; }
size_t ref_problem_nx(const ref_problem *p) { return p->prob.get_nx(); }
size_t ref_problem_nf(const ref_problem *p) { return p->prob.get_nf(); }
size_t ref_problem_nobj(const ref_problem *p) { return p->prob.get_nobj(); }
size_t ref_problem_nec(const ref_problem *p) { return p->prob.get_nec(); }
size_t ref_problem_nic(const ref_problem *p) { return p->prob.get_nic(); }
unsigned long long ref_problem_fevals(const ref_problem *p) { return p->prob.get_fevals(); }

int ref_problem_bounds(const ref_problem *p, double *lb, double *ub)
{
    return guarded([&] {
        const auto &l = p->prob.get_lb();
        const auto &u = p->prob.get_ub();
        std::copy(l.begin(), l.end(), lb);
        std::copy(u.begin(), u.end(), ub);
    });
}

int ref_problem_name(const ref_problem *p, char *buf, size_t buflen)
{
    return guarded([&] {
        const auto s = p->prob.get_name();
        if (buflen) {
            std::strncpy(buf, s.c_str(), buflen - 1);
            buf[buflen - 1] = 0;
        }
    });
}

int ref_problem_fitness(const ref_problem *p, const double *x, double *f)
{
    return guarded([&] {
        const pagmo::vector_double dv(x, x + p->prob.get_nx());
        const auto fv = p->prob.fitness(dv);
        std::copy(fv.begin(), fv.end(), f);
    });
}

int ref_problem_fitness_loop(const ref_problem *p, const double *dvs, size_t n, double *fvs)
{
    return guarded([&] {
        const auto nx = p->prob.get_nx(), nf = p->prob.get_nf();
        pagmo::vector_double dv(nx);
        for (size_t i = 0; i < n; ++i) {
            std::copy(dvs + i * nx, dvs + (i + 1) * nx, dv.begin());
            const auto fv = p->prob.fitness(dv);
            std::copy(fv.begin(), fv.end(), fvs + i * nf);
        }
    });
}

static int run_bfe(const pagmo::bfe &b, ref_problem *p, const double *dvs, size_t n, double *fvs, int nthreads)
{
    return guarded([&] {
        threads_env_guard g(nthreads);
        const pagmo::vector_double in(dvs, dvs + n * p->prob.get_nx());
        const auto out = b(p->prob, in);
        std::copy(out.begin(), out.end(), fvs);
    });
}

int ref_thread_bfe(ref_problem *p, const double *dvs, size_t n, double *fvs, int nthreads)
{
    return run_bfe(pagmo::bfe{pagmo::thread_bfe{}}, p, dvs, n, fvs, nthreads);
}

int ref_default_bfe(ref_problem *p, const double *dvs, size_t n, double *fvs, int nthreads)
{
    return run_bfe(pagmo::bfe{}, p, dvs, n, fvs, nthreads);
}

int ref_cec2014_tables(unsigned func, unsigned dim, double *Mr, double *Os, int *S)
{
    return guarded([&] {
        oracle_ref::ensure_cec2014_tables(func, dim);
        const auto &m = oracle_ref::cec2014_rotation(func, dim);
        const auto &o = oracle_ref::cec2014_shift(func);
        const auto &s = oracle_ref::cec2014_shuffle(func, dim);
        if (Mr) std::copy(m.begin(), m.end(), Mr);
        if (Os) std::copy(o.begin(), o.end(), Os);
        if (S) std::copy(s.begin(), s.end(), S);
    });
}

int ref_cec2013_tables(unsigned dim, double *Mr, double *Os)
{
    return guarded([&] {
        oracle_ref::ensure_cec2013_tables(dim);
        const auto &m = oracle_ref::cec2013_md(dim);
        const auto &o = oracle_ref::cec2013_shift();
        if (Mr) std::copy(m.begin(), m.end(), Mr);
        if (Os) std::copy(o.begin(), o.end(), Os);
    });
}

int ref_cec2014_origin_shift(const ref_problem *p, double *out, size_t cap, size_t *n)
{
    return guarded([&] {
        const auto *udp = p->prob.extract<pagmo::cec2014>();
        if (!udp) throw std::invalid_argument("not a cec2014 problem");
        const auto &v = udp->get_origin_shift();
        *n = v.size();
        if (v.size() > cap) throw std::invalid_argument("buffer too small");
        std::copy(v.begin(), v.end(), out);
    });
}

int ref_pareto_dominance(const double *a, const double *b, size_t m, int *out)
{
    return guarded([&] { *out = pagmo::pareto_dominance(pagmo::vector_double(a, a + m), pagmo::vector_double(b, b + m)) ? 1 : 0; });
}

int ref_fnds(const double *f, size_t n, size_t m, size_t *rank, size_t *dom_count, size_t *front_idx, size_t *front_off,
             size_t *nfronts, size_t *dl_idx, size_t *dl_off, size_t dl_cap)
{
    return guarded([&] {
        const auto r = pagmo::fast_non_dominated_sorting(unflatten(f, n, m));
        const auto &fronts = std::get<0>(r);
        const auto &dl = std::get<1>(r);
        const auto &dc = std::get<2>(r);
        const auto &rk = std::get<3>(r);
        if (rank) std::copy(rk.begin(), rk.end(), rank);
        if (dom_count) std::copy(dc.begin(), dc.end(), dom_count);
        size_t off = 0;
        for (size_t k = 0; k < fronts.size(); ++k) {
            if (front_off) front_off[k] = off;
            if (front_idx) std::copy(fronts[k].begin(), fronts[k].end(), front_idx + off);
            off += fronts[k].size();
        }
        if (front_off) front_off[fronts.size()] = off;
        if (nfronts) *nfronts = fronts.size();
        if (dl_off) {
            size_t o = 0;
            for (size_t i = 0; i < n; ++i) {
                dl_off[i] = o;
                if (dl_idx) {
                    if (o + dl[i].size() > dl_cap) throw std::invalid_argument("dom_list buffer too small");
                    std::copy(dl[i].begin(), dl[i].end(), dl_idx + o);
                }
                o += dl[i].size();
            }
            dl_off[n] = o;
        }
    });
}

int ref_crowding_distance(const double *f, size_t n, size_t m, double *out)
{
    return guarded([&] {
        const auto r = pagmo::crowding_distance(unflatten(f, n, m));
        std::copy(r.begin(), r.end(), out);
    });
}

int ref_sort_population_mo(const double *f, size_t n, size_t m, size_t *out)
{
    return guarded([&] {
        const auto r = pagmo::sort_population_mo(unflatten(f, n, m));
        std::copy(r.begin(), r.end(), out);
    });
}

int ref_select_best_N_mo(const double *f, size_t n, size_t m, size_t N, size_t *out, size_t *nout)
{
    return guarded([&] {
        const auto r = pagmo::select_best_N_mo(unflatten(f, n, m), N);
        std::copy(r.begin(), r.end(), out);
        if (nout) *nout = r.size();
    });
}

int ref_ideal(const double *f, size_t n, size_t m, double *out)
{
    return guarded([&] {
        const auto r = pagmo::ideal(unflatten(f, n, m));
        std::copy(r.begin(), r.end(), out);
    });
}

int ref_nadir(const double *f, size_t n, size_t m, double *out)
{
    return guarded([&] {
        const auto r = pagmo::nadir(unflatten(f, n, m));
        std::copy(r.begin(), r.end(), out);
    });
}

} // extern "C"
