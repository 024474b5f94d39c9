// oracle shim: boost::integral_constant family mapped onto the std one (problem.hpp:44, bfe.hpp:41).
#ifndef ORACLE_SHIM_BOOST_INTEGRAL_CONSTANT_HPP
#define ORACLE_SHIM_BOOST_INTEGRAL_CONSTANT_HPP
#include <type_traits>
namespace boost {
template <class T, T v> using integral_constant = std::integral_constant<T, v>;
using true_type = std::true_type;
using false_type = std::false_type;
}
#endif
