// oracle shim: tbb::parallel_for restated on std::thread. Static contiguous partition of the range over
// ORACLE_TBB_THREADS (env) or hardware_concurrency() workers. The body is COPIED once per worker, which
// mirrors TBB copying the body on every range split: a lambda that captured a pagmo::problem by value
// (thread_bfe.cpp:135, thread_safety::basic) therefore gets one problem copy per block, as in the reference.
#ifndef ORACLE_SHIM_TBB_PARALLEL_FOR_H
#define ORACLE_SHIM_TBB_PARALLEL_FOR_H
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <exception>
#include <mutex>
#include <thread>
#include <vector>

#include <tbb/blocked_range.h>

namespace tbb {
namespace shim_detail {
inline unsigned n_workers()
{
    if (const char *e = std::getenv("ORACLE_TBB_THREADS")) {
        const long v = std::strtol(e, nullptr, 10);
        if (v > 0) return static_cast<unsigned>(v);
    }
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? hc : 1u;
}
}
template <typename T, typename Body> void parallel_for(const blocked_range<T> &r, const Body &body)
{
    if (r.empty()) return;
    const std::size_t n = r.size();
    const std::size_t nw = std::min<std::size_t>(shim_detail::n_workers(), n);
    if (nw <= 1) {
        Body local(body);
        local(r);
        return;
    }
    std::vector<std::thread> pool;
    std::exception_ptr first_exc;
    std::mutex mtx;
    const std::size_t chunk = n / nw, rem = n % nw;
    T lo = r.begin();
    for (std::size_t w = 0; w < nw; ++w) {
        const T hi = static_cast<T>(lo + chunk + (w < rem ? 1 : 0));
        pool.emplace_back([&body, &first_exc, &mtx, lo, hi]() {
            try {
                Body local(body);
                local(blocked_range<T>(lo, hi));
            } catch (...) {
                std::lock_guard<std::mutex> lk(mtx);
                if (!first_exc) first_exc = std::current_exception();
            }
        });
        lo = hi;
    }
    for (auto &t : pool) t.join();
    if (first_exc) std::rethrow_exception(first_exc);
}
}
#endif
