// oracle shim: streaming of boost::tuple in Boost's default format "(a b)".  TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_SHIM_BOOST_TUPLE_TUPLE_IO_HPP
#define ORACLE_SHIM_BOOST_TUPLE_TUPLE_IO_HPP
#include <ostream>
#include <boost/tuple/tuple.hpp>
namespace boost
{
template <typename A, typename B>
inline std::ostream &operator<<(std::ostream &os, const tuple<A, B> &t)
{
    return os << '(' << std::get<0>(t) << ' ' << std::get<1>(t) << ')';
}
} // namespace boost
#endif
