// oracle shim: boost::indirect_iterator (archipelago.hpp:147-148: iteration over vector<unique_ptr<island>> yielding island&).
// TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_SHIM_BOOST_ITERATOR_INDIRECT_ITERATOR_HPP
#define ORACLE_SHIM_BOOST_ITERATOR_INDIRECT_ITERATOR_HPP
#include <iterator>
#include <type_traits>
namespace boost
{
template <typename It>
class indirect_iterator
{
    using inner_ref = decltype(**std::declval<It &>());
public:
    using iterator_category = std::random_access_iterator_tag;
    using value_type = typename std::remove_cv<typename std::remove_reference<inner_ref>::type>::type;
    using difference_type = typename std::iterator_traits<It>::difference_type;
    using reference = inner_ref;
    using pointer = typename std::remove_reference<inner_ref>::type *;
    indirect_iterator() = default;
    indirect_iterator(It it) : m_it(it) {}
    template <typename Other, typename = typename std::enable_if<std::is_convertible<Other, It>::value>::type>
    indirect_iterator(const indirect_iterator<Other> &o) : m_it(o.base()) {}
    It base() const { return m_it; }
    reference operator*() const { return **m_it; }
    pointer operator->() const { return &**m_it; }
    reference operator[](difference_type n) const { return **(m_it + n); }
    indirect_iterator &operator++() { ++m_it; return *this; }
    indirect_iterator operator++(int) { auto t = *this; ++m_it; return t; }
    indirect_iterator &operator--() { --m_it; return *this; }
    indirect_iterator operator--(int) { auto t = *this; --m_it; return t; }
    indirect_iterator &operator+=(difference_type n) { m_it += n; return *this; }
    indirect_iterator &operator-=(difference_type n) { m_it -= n; return *this; }
    friend indirect_iterator operator+(indirect_iterator a, difference_type n) { return a += n; }
    friend indirect_iterator operator+(difference_type n, indirect_iterator a) { return a += n; }
    friend indirect_iterator operator-(indirect_iterator a, difference_type n) { return a -= n; }
    friend difference_type operator-(const indirect_iterator &a, const indirect_iterator &b) { return a.m_it - b.m_it; }
    friend bool operator==(const indirect_iterator &a, const indirect_iterator &b) { return a.m_it == b.m_it; }
    friend bool operator!=(const indirect_iterator &a, const indirect_iterator &b) { return a.m_it != b.m_it; }
    friend bool operator<(const indirect_iterator &a, const indirect_iterator &b) { return a.m_it < b.m_it; }
    friend bool operator>(const indirect_iterator &a, const indirect_iterator &b) { return a.m_it > b.m_it; }
    friend bool operator<=(const indirect_iterator &a, const indirect_iterator &b) { return a.m_it <= b.m_it; }
    friend bool operator>=(const indirect_iterator &a, const indirect_iterator &b) { return a.m_it >= b.m_it; }
private:
    It m_it{};
};
} // namespace boost
#endif
