// oracle shim for boost::hash_combine (custom_comparisons.hpp:37).
#ifndef ORACLE_SHIM_BOOST_HASH_HPP
#define ORACLE_SHIM_BOOST_HASH_HPP
#include <cstddef>
#include <functional>
namespace boost {
template <class T> inline void hash_combine(std::size_t &seed, const T &v)
{
    seed ^= std::hash<T>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}
}
#endif
