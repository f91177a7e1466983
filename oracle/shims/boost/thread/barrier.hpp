// shim: boost::barrier as used by the reference's ParallelExecutor (src/ik_parallel.h:37,47,59-74,84-86):
// constructor with a thread count and wait().  Standard semantics (a reusable rendezvous of `count` threads) on
// std::mutex / std::condition_variable.  Test infrastructure, see README.md.
#pragma once
#include <condition_variable>
#include <mutex>
namespace boost
{
class barrier
{
    std::mutex m_;
    std::condition_variable cv_;
    const unsigned count_;
    unsigned waiting_ = 0, generation_ = 0;

public:
    explicit barrier(unsigned count) : count_(count ? count : 1) {}
    barrier(const barrier&) = delete;
    bool wait()
    {
        std::unique_lock<std::mutex> lock(m_);
        const unsigned gen = generation_;
        if(++waiting_ == count_)
        {
            waiting_ = 0;
            generation_++;
            cv_.notify_all();
            return true;
        }
        cv_.wait(lock, [&] { return gen != generation_; });
        return false;
    }
};
}
