// Host-side helpers shared by the loader, the TSV readers / writers and the driver.  Not installed.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstddef>
#include <exception>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

namespace soup {

// Runs fn(i) for i in [0, n) on up to `nthreads` threads, the caller's included (dynamic distribution, one item at a time).
// An exception thrown by fn stops the distribution and is rethrown on the calling thread after every helper has joined; if helper
// threads cannot be created the work still completes on the threads that exist.
template <typename F>
void parallel_for(unsigned nthreads, size_t n, F fn) {
    if (n == 0) return;
    const unsigned t_n = (unsigned)std::min<size_t>(std::max(1u, nthreads), n);
    if (t_n == 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
    std::atomic<size_t> next{0};
    std::exception_ptr failure;
    std::mutex mu;
    auto worker = [&] {
        try {
            for (size_t i; (i = next.fetch_add(1)) < n;) fn(i);
        } catch (...) {
            next.store(n);
            std::lock_guard<std::mutex> g(mu);
            if (!failure) failure = std::current_exception();
        }
    };
    std::vector<std::thread> pool;
    try {
        pool.reserve(t_n - 1);
        for (unsigned t = 0; t + 1 < t_n; ++t) pool.emplace_back(worker);
    } catch (...) {}   // out of threads: go on with the ones that started
    worker();
    for (auto& th : pool) th.join();
    if (failure) std::rethrow_exception(failure);
}

inline unsigned host_threads() { return std::max(1u, std::min(64u, std::thread::hardware_concurrency())); }

}  // namespace soup

// Closes a function-try-block of an extern "C" entry point that returns a status: no C++ exception leaves the library.
#define SOUP_CATCH_RETURN(fn_name_)                                                                                                    \
    catch (const std::bad_alloc&) { soup::set_thread_error("%s: out of memory", fn_name_); return SOUP_ERR_NOMEM; }                   \
    catch (const std::exception& e_) { soup::set_thread_error("%s: %s", fn_name_, e_.what()); return SOUP_ERR_INVALID; }              \
    catch (...) { soup::set_thread_error("%s: unknown C++ exception", fn_name_); return SOUP_ERR_INVALID; }
// The same for entry points that take a `soup_ctx* ctx`: the message goes where soup_last_error(ctx) reads it.
#define SOUP_CATCH_RETURN_CTX(fn_name_)                                                                                                \
    catch (const std::bad_alloc&) { return ctx ? ctx->fail(SOUP_ERR_NOMEM, "%s: out of memory", fn_name_) : SOUP_ERR_NOMEM; }          \
    catch (const std::exception& e_) { return ctx ? ctx->fail(SOUP_ERR_INVALID, "%s: %s", fn_name_, e_.what()) : SOUP_ERR_INVALID; }   \
    catch (...) { return ctx ? ctx->fail(SOUP_ERR_INVALID, "%s: unknown C++ exception", fn_name_) : SOUP_ERR_INVALID; }
