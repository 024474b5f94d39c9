// oracle shim: boost::make_zip_iterator over a tuple of two iterators (base_bgl_topology.cpp:273-278).  TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_SHIM_BOOST_ITERATOR_ZIP_ITERATOR_HPP
#define ORACLE_SHIM_BOOST_ITERATOR_ZIP_ITERATOR_HPP
#include <iterator>
#include <boost/tuple/tuple.hpp>
namespace boost
{
template <typename A, typename B>
class zip_iterator2
{
public:
    using iterator_category = std::forward_iterator_tag;
    using value_type = tuple<typename std::iterator_traits<A>::value_type, typename std::iterator_traits<B>::value_type>;
    using difference_type = std::ptrdiff_t;
    using pointer = const value_type *;
    using reference = value_type;
    zip_iterator2(A a, B b) : m_a(a), m_b(b) {}
    value_type operator*() const { return value_type(*m_a, *m_b); }
    zip_iterator2 &operator++() { ++m_a; ++m_b; return *this; }
    bool operator==(const zip_iterator2 &o) const { return m_a == o.m_a; }
    bool operator!=(const zip_iterator2 &o) const { return m_a != o.m_a; }
private:
    A m_a;
    B m_b;
};
template <typename A, typename B>
inline zip_iterator2<A, B> make_zip_iterator(const tuple<A, B> &t)
{
    return zip_iterator2<A, B>(std::get<0>(t), std::get<1>(t));
}
} // namespace boost
#endif
