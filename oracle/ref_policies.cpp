// oracle/ref_policies.cpp - C entry points over the UNMODIFIED reference migration policies (fair_replace, select_best).
// TEST INFRASTRUCTURE ONLY; see ref_capi.h.
#include <cstring>
#include <stdexcept>
#include <tuple>
#include <vector>

#include <pagmo/r_policies/fair_replace.hpp>
#include <pagmo/s_policies/select_best.hpp>
#include <pagmo/types.hpp>

#include "ref_capi.h"

extern "C" void ref_set_error(const char *what);

namespace
{
pagmo::individuals_group_t group_of(const unsigned long long *ids, const double *x, const double *f, size_t n, size_t nx, size_t nf)
{
    pagmo::individuals_group_t g;
    for (size_t i = 0; i < n; ++i) {
        std::get<0>(g).push_back(ids[i]);
        std::get<1>(g).emplace_back(x + i * nx, x + (i + 1) * nx);
        std::get<2>(g).emplace_back(f + i * nf, f + (i + 1) * nf);
    }
    return g;
}
size_t flatten(const pagmo::individuals_group_t &g, unsigned long long *ids, double *x, double *f, size_t nx, size_t nf)
{
    const size_t n = std::get<0>(g).size();
    for (size_t i = 0; i < n; ++i) {
        ids[i] = std::get<0>(g)[i];
        std::memcpy(x + i * nx, std::get<1>(g)[i].data(), nx * sizeof(double));
        std::memcpy(f + i * nf, std::get<2>(g)[i].data(), nf * sizeof(double));
    }
    return n;
}
} // namespace

extern "C" {

// fair_replace{rate}.replace(inds, nx, 0, nobj, 0, 0, {}, mig) (fair_replace.cpp:63-221).  rate_is_frac: fractional rate.
int ref_fair_replace(const unsigned long long *ids, const double *x, const double *f, size_t n, size_t nx, size_t nobj, int rate_is_frac,
                     double rate, const unsigned long long *mids, const double *mx, const double *mf, size_t nm, unsigned long long *ids_out,
                     double *x_out, double *f_out)
{
    try {
        const auto pol = rate_is_frac ? pagmo::fair_replace(rate) : pagmo::fair_replace(static_cast<pagmo::pop_size_t>(rate));
        const auto out = pol.replace(group_of(ids, x, f, n, nx, nobj), nx, 0, nobj, 0, 0, {}, group_of(mids, mx, mf, nm, nx, nobj));
        if (flatten(out, ids_out, x_out, f_out, nx, nobj) != n) throw std::runtime_error("fair_replace changed the population size");
        return 0;
    } catch (const std::exception &e) {
        ref_set_error(e.what());
        return 1;
    }
}

// the constrained branches: rows of nf = 1 + nec + nic doubles
int ref_fair_replace_con(const unsigned long long *ids, const double *x, const double *f, size_t n, size_t nx, size_t nec, size_t nic,
                         const double *tol, int rate_is_frac, double rate, const unsigned long long *mids, const double *mx, const double *mf,
                         size_t nm, unsigned long long *ids_out, double *x_out, double *f_out)
{
    try {
        const size_t nf = 1 + nec + nic;
        const auto pol = rate_is_frac ? pagmo::fair_replace(rate) : pagmo::fair_replace(static_cast<pagmo::pop_size_t>(rate));
        const auto out = pol.replace(group_of(ids, x, f, n, nx, nf), nx, 0, 1, nec, nic, pagmo::vector_double(tol, tol + nec + nic),
                                     group_of(mids, mx, mf, nm, nx, nf));
        if (flatten(out, ids_out, x_out, f_out, nx, nf) != n) throw std::runtime_error("fair_replace changed the population size");
        return 0;
    } catch (const std::exception &e) {
        ref_set_error(e.what());
        return 1;
    }
}
int ref_select_best_con(const unsigned long long *ids, const double *x, const double *f, size_t n, size_t nx, size_t nec, size_t nic,
                        const double *tol, int rate_is_frac, double rate, unsigned long long *ids_out, double *x_out, double *f_out, size_t *n_out)
{
    try {
        const size_t nf = 1 + nec + nic;
        const auto pol = rate_is_frac ? pagmo::select_best(rate) : pagmo::select_best(static_cast<pagmo::pop_size_t>(rate));
        *n_out = flatten(pol.select(group_of(ids, x, f, n, nx, nf), nx, 0, 1, nec, nic, pagmo::vector_double(tol, tol + nec + nic)), ids_out,
                         x_out, f_out, nx, nf);
        return 0;
    } catch (const std::exception &e) {
        ref_set_error(e.what());
        return 1;
    }
}

// select_best{rate}.select(inds, nx, 0, nobj, 0, 0, {}) (select_best.cpp:63-171); outputs sized n
int ref_select_best(const unsigned long long *ids, const double *x, const double *f, size_t n, size_t nx, size_t nobj, int rate_is_frac,
                    double rate, unsigned long long *ids_out, double *x_out, double *f_out, size_t *n_out)
{
    try {
        const auto pol = rate_is_frac ? pagmo::select_best(rate) : pagmo::select_best(static_cast<pagmo::pop_size_t>(rate));
        *n_out = flatten(pol.select(group_of(ids, x, f, n, nx, nobj), nx, 0, nobj, 0, 0, {}), ids_out, x_out, f_out, nx, nobj);
        return 0;
    } catch (const std::exception &e) {
        ref_set_error(e.what());
        return 1;
    }
}
}

// ---- hypervolume (src/utils/hypervolume.cpp, hv_algos/hv_hv2d.cpp, hv_hv3d.cpp) through pagmo::hypervolume with the reference's own
// choice of algorithm (get_best_compute / get_best_contributions: hv2d for 2 objectives, hv3d for 3, hvwfg beyond)
#include <pagmo/utils/hv_algos/hv_bf_approx.hpp>
#include <pagmo/utils/hv_algos/hv_bf_fpras.hpp>
#include <pagmo/utils/hypervolume.hpp>
namespace
{
std::vector<pagmo::vector_double> rows(const double *f, size_t n, size_t m)
{
    std::vector<pagmo::vector_double> v(n);
    for (size_t i = 0; i < n; ++i) v[i].assign(f + i * m, f + (i + 1) * m);
    return v;
}
} // namespace
extern "C" {
// the approximation algorithms through pagmo::hypervolume with an explicit algorithm object (reference defaults but eps / delta / seed)
int ref_hv_fpras(const double *f, size_t n, size_t m, const double *r, double eps, double delta, unsigned seed, double *out)
{
    try {
        pagmo::hypervolume hv(rows(f, n, m), false);
        pagmo::bf_fpras algo(eps, delta, seed);
        *out = hv.compute(pagmo::vector_double(r, r + m), algo);
        return 0;
    } catch (const std::exception &e) {
        ref_set_error(e.what());
        return 1;
    }
}
int ref_hv_approx_extreme(const double *f, size_t n, size_t m, const double *r, int greatest, int use_exact, double eps, double delta,
                          unsigned seed, size_t *out)
{
    try {
        pagmo::hypervolume hv(rows(f, n, m), false);
        pagmo::bf_approx algo(use_exact != 0, 1u, eps, delta, 0.775, 0.2, 0.1, 0.25, seed);
        const pagmo::vector_double rp(r, r + m);
        *out = static_cast<size_t>(greatest ? hv.greatest_contributor(rp, algo) : hv.least_contributor(rp, algo));
        return 0;
    } catch (const std::exception &e) {
        ref_set_error(e.what());
        return 1;
    }
}
int ref_hv_compute(const double *f, size_t n, size_t m, const double *r, double *out)
{
    try {
        *out = pagmo::hypervolume(rows(f, n, m), true).compute(pagmo::vector_double(r, r + m));
        return 0;
    } catch (const std::exception &e) {
        ref_set_error(e.what());
        return 1;
    }
}
int ref_hv_contributions(const double *f, size_t n, size_t m, const double *r, double *out)
{
    try {
        const auto c = pagmo::hypervolume(rows(f, n, m), true).contributions(pagmo::vector_double(r, r + m));
        std::memcpy(out, c.data(), n * sizeof(double));
        return 0;
    } catch (const std::exception &e) {
        ref_set_error(e.what());
        return 1;
    }
}
}
