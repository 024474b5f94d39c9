// oracle shim: boost::tuple / make_tuple for base_bgl_topology.cpp:273-278 (pairs of iterators / values).  TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_SHIM_BOOST_TUPLE_TUPLE_HPP
#define ORACLE_SHIM_BOOST_TUPLE_TUPLE_HPP
#include <tuple>
namespace boost
{
template <typename... T>
struct tuple : std::tuple<T...> {
    using std::tuple<T...>::tuple;
    tuple(const std::tuple<T...> &t) : std::tuple<T...>(t) {}
};
template <typename... T>
inline tuple<typename std::decay<T>::type...> make_tuple(T &&...t)
{
    return tuple<typename std::decay<T>::type...>(std::forward<T>(t)...);
}
} // namespace boost
#endif
