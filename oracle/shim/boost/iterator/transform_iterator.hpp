// oracle shim: boost::make_transform_iterator (base_bgl_topology.cpp:274-278).  TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_SHIM_BOOST_ITERATOR_TRANSFORM_ITERATOR_HPP
#define ORACLE_SHIM_BOOST_ITERATOR_TRANSFORM_ITERATOR_HPP
#include <iterator>
#include <utility>
namespace boost
{
template <typename F, typename It>
class transform_iterator
{
public:
    using iterator_category = std::forward_iterator_tag;
    using value_type = decltype(std::declval<F &>()(*std::declval<It &>()));
    using difference_type = std::ptrdiff_t;
    using pointer = const value_type *;
    using reference = value_type;
    transform_iterator(It it, F f) : m_it(it), m_f(f) {}
    value_type operator*() const { return m_f(*m_it); }
    transform_iterator &operator++() { ++m_it; return *this; }
    bool operator==(const transform_iterator &o) const { return m_it == o.m_it; }
    bool operator!=(const transform_iterator &o) const { return m_it != o.m_it; }
private:
    It m_it;
    mutable F m_f;
};
template <typename It, typename F>
inline transform_iterator<F, It> make_transform_iterator(It it, F f)
{
    return transform_iterator<F, It>(it, f);
}
} // namespace boost
#endif
