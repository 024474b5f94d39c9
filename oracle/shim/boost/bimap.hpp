// oracle/shim/boost/bimap.hpp - minimal stand-in for boost::bimap so that the reference's sga.cpp compiles unmodified
// (insert(value_type(l, r)), .left.at(l), .right.at(r) are all it uses, sga.cpp:73-110,169-171,313-325).  TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_SHIM_BOOST_BIMAP_HPP
#define ORACLE_SHIM_BOOST_BIMAP_HPP
#include <map>
#include <utility>

namespace boost
{
template <typename L, typename R>
class bimap
{
public:
    struct value_type {
        value_type(L l, R r) : first(std::move(l)), second(std::move(r)) {}
        L first;
        R second;
    };
    void insert(const value_type &v)
    {
        left.emplace(v.first, v.second);
        right.emplace(v.second, v.first);
    }
    std::map<L, R> left;
    std::map<R, L> right;
};
} // namespace boost
#endif
