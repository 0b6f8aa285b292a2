// A few persistent helper threads for the O(bytes) host passes of a chunk (the row scan of
// the planner, the decode of packed fields).  Workers sleep on a condition variable between
// jobs; a forked child runs everything inline (threads do not survive fork).
#pragma once
#include <functional>

namespace spx {

class HostPool {
public:
    static HostPool& get();
    // Threads available to one job, caller included (SPX_HOST_THREADS, default 4 for the
    // scan; at most 16 and at most the cores of the affinity mask).
    int max_threads() const;
    // fn(part, n_parts) on n_parts = min(n_threads, max_threads()) threads (n_threads <= 0:
    // the default); returns when every part is done.  One job at a time.
    void run(const std::function<void(int, int)>& fn, int n_threads = 0);

private:
    HostPool();
    struct Impl;
    Impl* impl_;
};

}  // namespace spx
