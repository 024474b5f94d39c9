// oracle shim for boost::math::constants (cec2014.cpp:36, wfg.cpp:39). The long-double literals round
// to the same doubles Boost returns (nearest-double pi and e).
#ifndef ORACLE_SHIM_BOOST_CONSTANTS_HPP
#define ORACLE_SHIM_BOOST_CONSTANTS_HPP
namespace boost { namespace math { namespace constants {
template <typename T> constexpr T pi() { return static_cast<T>(3.141592653589793238462643383279502884L); }
template <typename T> constexpr T e() { return static_cast<T>(2.718281828459045235360287471352662498L); }
template <typename T> constexpr T half_pi() { return static_cast<T>(1.570796326794896619231321691639751442L); }
}}}
#endif
