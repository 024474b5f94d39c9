// oracle shim for boost::safe_numerics::safe<T>: an integer whose products are overflow-checked (unconstrain.cpp:247-254,
// the only use: `safe<size_type>(n_dvs) * nobj` as a vector size).
#ifndef ORACLE_SHIM_BOOST_SAFE_INTEGER_HPP
#define ORACLE_SHIM_BOOST_SAFE_INTEGER_HPP
#include <stdexcept>
#include <type_traits>
namespace boost { namespace safe_numerics {
template <typename T> class safe
{
    static_assert(std::is_integral<T>::value, "safe<T>: integral types only");
    T m_v;

public:
    constexpr safe(T v = T()) : m_v(v) {}
    constexpr operator T() const { return m_v; }
    template <typename U> friend safe operator*(safe a, U b)
    {
        T r;
        if (__builtin_mul_overflow(a.m_v, static_cast<T>(b), &r)) throw std::overflow_error("safe<T>: multiplication overflow");
        return safe(r);
    }
    template <typename U> friend safe operator+(safe a, U b)
    {
        T r;
        if (__builtin_add_overflow(a.m_v, static_cast<T>(b), &r)) throw std::overflow_error("safe<T>: addition overflow");
        return safe(r);
    }
};
}}
#endif
