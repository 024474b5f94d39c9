// cuda_hypervolume.hpp - header-only hypervolume algorithm that runs on the device, behind pagmo's own extension point for
// hypervolume computations: a class derived from pagmo::hv_algorithm (reference include/pagmo/utils/hv_algos/hv_algorithm.hpp:91-216),
// to be passed to hypervolume::compute / exclusive / least_contributor / greatest_contributor / contributions
// (reference src/utils/hypervolume.cpp:196-404), e.g.
//
//     pagmo::hypervolume hv{pop};                    // or hypervolume{points}
//     pagmo_cuda::cuda_hv algo{/*device=*/0};
//     double v = hv.compute(ref_point, algo);         // replaces hv2d / hv3d (hv_hv2d.cpp:59-84, hv_hv3d.cpp:107-166)
//     auto c   = hv.contributions(ref_point, algo);   // replaces hv2d::contributions / HyCon3D (hv_hv3d.cpp:170-343)
//
// 2 and 3 objectives run the device versions of hv2d / hv3d, 4 to 12 objectives the device WFG (replaces hvwfg, hv_hvwfg.cpp:64-117:
// one thread per term of the top-level sum / per (point, term) for the contributions); more make verify_before_compute throw.
// No CPU fallback.
#ifndef PAGMO_CUDA_CUDA_HYPERVOLUME_HPP
#define PAGMO_CUDA_CUDA_HYPERVOLUME_HPP

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <pagmo/exceptions.hpp>
#include <pagmo/rng.hpp>
#include <pagmo/types.hpp>
#include <pagmo/utils/hv_algos/hv_algorithm.hpp>

#include <pagmo_cuda/cuda_bfe.hpp>
#include <pagmo_cuda/pgc.h>

namespace pagmo_cuda
{

namespace detail
{
inline std::vector<double> flatten_points(const std::vector<pagmo::vector_double> &points, std::size_t m)
{
    std::vector<double> flat(points.size() * m);
    for (std::size_t i = 0; i < points.size(); ++i) {
        if (points[i].size() != m) pagmo_throw(std::invalid_argument, "cuda hypervolume: a point and the reference point differ in dimension");
        for (std::size_t d = 0; d < m; ++d) flat[i * m + d] = points[i][d];
    }
    return flat;
}
} // namespace detail

class cuda_hv final : public pagmo::hv_algorithm
{
public:
    explicit cuda_hv(int device = 0) : m_device(device), m_ctx(detail::device_context(device)) {}

    double compute(std::vector<pagmo::vector_double> &points, const pagmo::vector_double &r_point) const override
    {
        const auto flat = flatten(points, r_point.size());
        double hv = 0.0;
        std::lock_guard<std::mutex> lk(detail::device_mutex(m_device));
        detail::check(pgc_hv_compute_host(m_ctx.get(), flat.data(), points.size(), r_point.size(), r_point.data(), &hv), "pgc_hv_compute_host");
        return hv;
    }
    std::vector<double> contributions(std::vector<pagmo::vector_double> &points, const pagmo::vector_double &r_point) const override
    {
        const auto flat = flatten(points, r_point.size());
        std::vector<double> c(points.size());
        std::lock_guard<std::mutex> lk(detail::device_mutex(m_device));
        detail::check(pgc_hv_contributions_host(m_ctx.get(), flat.data(), points.size(), r_point.size(), r_point.data(), c.data()),
                      "pgc_hv_contributions_host");
        return c;
    }
    // exclusive / least_contributor / greatest_contributor: the base class derives them from compute(); one batched
    // contributions() call is cheaper than n hypervolumes
    double exclusive(unsigned p_idx, std::vector<pagmo::vector_double> &points, const pagmo::vector_double &r_point) const override
    {
        if (p_idx >= points.size()) pagmo_throw(std::invalid_argument, "Index of the individual is out of bounds.");
        return contributions(points, r_point)[p_idx];
    }
    unsigned long long least_contributor(std::vector<pagmo::vector_double> &points, const pagmo::vector_double &r_point) const override
    {
        const auto c = contributions(points, r_point);
        unsigned long long best = 0;
        for (std::size_t i = 1; i < c.size(); ++i)
            if (c[i] < c[best]) best = i;
        return best;
    }
    unsigned long long greatest_contributor(std::vector<pagmo::vector_double> &points, const pagmo::vector_double &r_point) const override
    {
        const auto c = contributions(points, r_point);
        unsigned long long best = 0;
        for (std::size_t i = 1; i < c.size(); ++i)
            if (c[i] > c[best]) best = i;
        return best;
    }
    void verify_before_compute(const std::vector<pagmo::vector_double> &points, const pagmo::vector_double &r_point) const override
    {
        if (r_point.size() < 2u || r_point.size() > 12u) {
            pagmo_throw(std::invalid_argument, "Algorithm cuda_hv works for 2 to 12 objectives");
        }
        hv_algorithm::assert_minimisation(points, r_point);
    }
    std::shared_ptr<pagmo::hv_algorithm> clone() const override
    {
        return std::shared_ptr<pagmo::hv_algorithm>(new cuda_hv(*this));
    }
    std::string get_name() const override
    {
        return "cuda_hv algorithm (sm_100a, device " + std::to_string(m_device) + ")";
    }

private:
    static std::vector<double> flatten(const std::vector<pagmo::vector_double> &points, std::size_t m)
    {
        std::vector<double> flat(points.size() * m);
        for (std::size_t i = 0; i < points.size(); ++i) {
            if (points[i].size() != m) pagmo_throw(std::invalid_argument, "cuda_hv: a point and the reference point differ in dimension");
            for (std::size_t d = 0; d < m; ++d) flat[i * m + d] = points[i][d];
        }
        return flat;
    }
    int m_device;
    std::shared_ptr<pgc_ctx> m_ctx;
};

// pagmo::bf_fpras (hv_bf_fpras.hpp:63): (eps, delta) approximation of the hypervolume on the device (pgc_hv_fpras_host).  Like the
// reference class it only computes: exclusive / least / greatest contributor throw.
class cuda_bf_fpras final : public pagmo::hv_algorithm
{
public:
    explicit cuda_bf_fpras(double eps = 1e-2, double delta = 1e-2, unsigned seed = pagmo::random_device::next(), int device = 0)
        : m_eps(eps), m_delta(delta), m_seed(seed), m_device(device), m_ctx(detail::device_context(device))
    {
        if (eps <= 0 || eps > 1) pagmo_throw(std::invalid_argument, "Epsilon needs to be a probability greater then zero");
        if (delta <= 0 || delta > 1) pagmo_throw(std::invalid_argument, "Delta needs to be a probability greater than zero");
    }
    double compute(std::vector<pagmo::vector_double> &points, const pagmo::vector_double &r_point) const override
    {
        const auto flat = detail::flatten_points(points, r_point.size());
        double hv = 0.0;
        std::lock_guard<std::mutex> lk(detail::device_mutex(m_device));
        detail::check(pgc_hv_fpras_host(m_ctx.get(), flat.data(), points.size(), r_point.size(), r_point.data(), m_eps, m_delta, m_seed++, &hv),
                      "pgc_hv_fpras_host");
        return hv;
    }
    double exclusive(unsigned, std::vector<pagmo::vector_double> &, const pagmo::vector_double &) const override
    {
        pagmo_throw(std::invalid_argument, "This method is not supported by the bf_fpras algorithm");
    }
    unsigned long long least_contributor(std::vector<pagmo::vector_double> &, const pagmo::vector_double &) const override
    {
        pagmo_throw(std::invalid_argument, "This method is not supported by the bf_fpras algorithm");
    }
    unsigned long long greatest_contributor(std::vector<pagmo::vector_double> &, const pagmo::vector_double &) const override
    {
        pagmo_throw(std::invalid_argument, "This method is not supported by the bf_fpras algorithm");
    }
    std::vector<double> contributions(std::vector<pagmo::vector_double> &, const pagmo::vector_double &) const override
    {
        pagmo_throw(std::invalid_argument, "This method is not supported by the bf_fpras algorithm");
    }
    void verify_before_compute(const std::vector<pagmo::vector_double> &points, const pagmo::vector_double &r_point) const override
    {
        hv_algorithm::assert_minimisation(points, r_point);
    }
    std::shared_ptr<pagmo::hv_algorithm> clone() const override { return std::shared_ptr<pagmo::hv_algorithm>(new cuda_bf_fpras(*this)); }
    std::string get_name() const override { return "bf_fpras algorithm (sm_100a, device " + std::to_string(m_device) + ")"; }

private:
    double m_eps, m_delta;
    mutable unsigned m_seed; // successive calls continue like the reference's engine: a new stream each
    int m_device;
    std::shared_ptr<pgc_ctx> m_ctx;
};

// pagmo::bf_approx (hv_bf_approx.hpp:73): the Bringmann-Friedrich approximation of the least / greatest contributor on the device
// (pgc_hv_approx_extreme_host).  Like the reference class it cannot compute the hypervolume itself.
class cuda_bf_approx final : public pagmo::hv_algorithm
{
public:
    explicit cuda_bf_approx(bool use_exact = true, unsigned trivial_subcase_size = 1, double eps = 1e-2, double delta = 1e-6,
                            double delta_multiplier = 0.775, double alpha = 0.2, double initial_delta_coeff = 0.1, double gamma = 0.25,
                            unsigned seed = pagmo::random_device::next(), int device = 0)
        : m_use_exact(use_exact), m_trivial(trivial_subcase_size), m_eps(eps), m_delta(delta), m_delta_multiplier(delta_multiplier), m_alpha(alpha),
          m_initial_delta_coeff(initial_delta_coeff), m_gamma(gamma), m_seed(seed), m_device(device), m_ctx(detail::device_context(device))
    {
        if (eps < 0 || eps > 1) pagmo_throw(std::invalid_argument, "Epsilon needs to be a probability.");
        if (delta < 0 || delta > 1) pagmo_throw(std::invalid_argument, "Delta needs to be a probability.");
    }
    double compute(std::vector<pagmo::vector_double> &, const pagmo::vector_double &) const override
    {
        pagmo_throw(std::invalid_argument, "This algorithm can just approximate extreme contributions but not the hypervolume itself.");
    }
    unsigned long long least_contributor(std::vector<pagmo::vector_double> &points, const pagmo::vector_double &r_point) const override
    {
        return extreme(points, r_point, 0);
    }
    unsigned long long greatest_contributor(std::vector<pagmo::vector_double> &points, const pagmo::vector_double &r_point) const override
    {
        return extreme(points, r_point, 1);
    }
    void verify_before_compute(const std::vector<pagmo::vector_double> &points, const pagmo::vector_double &r_point) const override
    {
        hv_algorithm::assert_minimisation(points, r_point);
    }
    std::shared_ptr<pagmo::hv_algorithm> clone() const override { return std::shared_ptr<pagmo::hv_algorithm>(new cuda_bf_approx(*this)); }
    std::string get_name() const override { return "Bringmann-Friedrich approximation method (sm_100a, device " + std::to_string(m_device) + ")"; }

private:
    unsigned long long extreme(const std::vector<pagmo::vector_double> &points, const pagmo::vector_double &r_point, int greatest) const
    {
        const auto flat = detail::flatten_points(points, r_point.size());
        std::size_t idx = 0;
        std::lock_guard<std::mutex> lk(detail::device_mutex(m_device));
        detail::check(pgc_hv_approx_extreme_host(m_ctx.get(), flat.data(), points.size(), r_point.size(), r_point.data(), greatest,
                                                 m_use_exact ? 1 : 0, m_trivial, m_eps, m_delta, m_delta_multiplier, m_alpha, m_initial_delta_coeff,
                                                 m_gamma, m_seed++, &idx),
                      "pgc_hv_approx_extreme_host");
        return idx;
    }
    bool m_use_exact;
    unsigned m_trivial;
    double m_eps, m_delta, m_delta_multiplier, m_alpha, m_initial_delta_coeff, m_gamma;
    mutable unsigned m_seed;
    int m_device;
    std::shared_ptr<pgc_ctx> m_ctx;
};

} // namespace pagmo_cuda

#endif
