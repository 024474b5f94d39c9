// oracle shim: minimal tbb::blocked_range (TBB is absent in this image).
// Used by the unmodified reference thread_bfe.cpp:129-137, bfe_impl.cpp:80-81,121-122, translate.cpp:129.
#ifndef ORACLE_SHIM_TBB_BLOCKED_RANGE_H
#define ORACLE_SHIM_TBB_BLOCKED_RANGE_H
#include <cstddef>
namespace tbb {
template <typename T> class blocked_range
{
public:
    using const_iterator = T;
    blocked_range(T b, T e, std::size_t grain = 1) : m_b(b), m_e(e), m_g(grain) {}
    T begin() const { return m_b; }
    T end() const { return m_e; }
    std::size_t size() const { return static_cast<std::size_t>(m_e - m_b); }
    std::size_t grainsize() const { return m_g; }
    bool empty() const { return !(m_b < m_e); }
private:
    T m_b, m_e;
    std::size_t m_g;
};
}
#endif
