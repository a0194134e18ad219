// Host-side helper shared by the loader and the TSV readers / writers.  Not installed.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstddef>
#include <thread>
#include <vector>

namespace soup {

// Runs fn(i) for i in [0, n) on up to `nthreads` threads (dynamic distribution, one item at a time).
template <typename F>
void parallel_for(unsigned nthreads, size_t n, F fn) {
    if (n == 0) return;
    const unsigned t_n = (unsigned)std::min<size_t>(std::max(1u, nthreads), n);
    if (t_n == 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < t_n; ++t)
        pool.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i); });
    for (auto& th : pool) th.join();
}

inline unsigned host_threads() { return std::max(1u, std::min(64u, std::thread::hardware_concurrency())); }

}  // namespace soup
