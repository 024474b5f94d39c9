// oracle shim: tbb::task_group::run_and_wait (thread_island.cpp:140-150, use_pool = true) - runs the task on the calling
// thread, which is what TBB may do as well.  TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_SHIM_TBB_TASK_GROUP_H
#define ORACLE_SHIM_TBB_TASK_GROUP_H
namespace tbb
{
class task_group
{
public:
    template <typename F>
    void run_and_wait(F &&f)
    {
        f();
    }
};
} // namespace tbb
#endif
