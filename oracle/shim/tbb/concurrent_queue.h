// oracle shim: tbb::concurrent_queue (island.cpp:202-260: a global cache of task_queue objects) on a mutex + std::queue.
// TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_SHIM_TBB_CONCURRENT_QUEUE_H
#define ORACLE_SHIM_TBB_CONCURRENT_QUEUE_H
#include <mutex>
#include <queue>
#include <utility>
namespace tbb
{
template <typename T>
class concurrent_queue
{
public:
    void push(T &&v)
    {
        std::lock_guard<std::mutex> lk(m_mtx);
        m_q.push(std::move(v));
    }
    bool try_pop(T &out)
    {
        std::lock_guard<std::mutex> lk(m_mtx);
        if (m_q.empty()) return false;
        out = std::move(m_q.front());
        m_q.pop();
        return true;
    }
private:
    std::mutex m_mtx;
    std::queue<T> m_q;
};
} // namespace tbb
#endif
