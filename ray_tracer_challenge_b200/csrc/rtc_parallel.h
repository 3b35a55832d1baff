// rtc_parallel.h — the host half's worker threads.
//
// A one-shot Camera::render_b200 of a 100 k-primitive scene spends more time flattening and committing the scene on
// the host than rendering it on the device, and that work is a handful of passes over arrays of
// 100 k+ records whose cost is memory latency: it splits over threads cleanly.  One lazily started pool per process
// (per library: the header is shared by librtc_b200.so and librtc_host.so), workers asleep between jobs.  Size: the
// host's cores (at most 16), divided by LOCAL_WORLD_SIZE when torchrun runs one process per GPU; RTC_HOST_THREADS
// overrides it.
//
//   rtc::parallel_for(n, grain, [&](size_t begin, size_t end, int chunk) { ... });
//
// runs the body on contiguous chunks of [0, n) — at most 4 per thread, none smaller than `grain` unless n is — and
// returns when all are done; with one chunk (small n, RTC_HOST_THREADS=1, or a call from inside another body) it is a
// plain call on the caller's thread.  The first exception a body throws is rethrown on the caller's thread.
#pragma once

#include <unistd.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <exception>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace rtc {

class WorkerPool {
   public:
    static WorkerPool& instance() {
        static std::mutex guard;
        static WorkerPool* pool = nullptr;
        std::lock_guard<std::mutex> lock(guard);
        if (!pool || pool->pid_ != getpid()) pool = new WorkerPool();  // a forked child has none of the parent's threads
        return *pool;
    }
    int threads() const { return (int)workers_.size() + 1; }

    // job(i) for every i in [0, n_jobs), on the workers and the calling thread
    void run(int n_jobs, const std::function<void(int)>& job) {
        if (n_jobs <= 0) return;
        if (n_jobs == 1 || workers_.empty() || inside_job()) {
            for (int i = 0; i < n_jobs; i++) job(i);
            return;
        }
        std::lock_guard<std::mutex> one_at_a_time(run_mutex_);
        job_ = &job, n_jobs_ = n_jobs, error_ = nullptr;
        next_.store(0, std::memory_order_relaxed);
        pending_.store((int)workers_.size(), std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lock(m_);
            generation_.fetch_add(1, std::memory_order_release);
            if (sleepers_ > 0) wake_.notify_all();
        }
        work();
        // the passes of a commit follow one another within microseconds: spin briefly before sleeping, on both sides
        for (int spin = 0; spin < kSpins && pending_.load(std::memory_order_acquire) != 0; spin++) cpu_relax();
        if (pending_.load(std::memory_order_acquire) != 0) {
            std::unique_lock<std::mutex> lock(m_);
            done_.wait(lock, [&] { return pending_.load(std::memory_order_acquire) == 0; });
        }
        job_ = nullptr;
        if (error_) std::rethrow_exception(error_);
    }

   private:
    static constexpr int kSpins = 4000;  // ~20-40 us
    static void cpu_relax() {
#if defined(__SSE2__)
        _mm_pause();
#else
        std::this_thread::yield();
#endif
    }
    WorkerPool() : pid_(getpid()) {
        unsigned n = std::max(1u, std::thread::hardware_concurrency());
        // one process per GPU (torchrun): the ranks of a node share its cores
        if (const char* ranks = getenv("LOCAL_WORLD_SIZE")) n = std::max(1u, n / (unsigned)std::max(1, atoi(ranks)));
        n = std::min(16u, n);
        if (const char* env = getenv("RTC_HOST_THREADS")) n = (unsigned)std::max(1, atoi(env));
        for (unsigned i = 1; i < n; i++) {
            workers_.emplace_back([this] { worker(); });
            workers_.back().detach();  // they sleep on `wake_` for the life of the process
        }
    }
    static bool& inside_job() {
        static thread_local bool inside = false;
        return inside;
    }
    void work() {
        inside_job() = true;
        for (;;) {
            const int i = next_.fetch_add(1, std::memory_order_relaxed);
            if (i >= n_jobs_) break;
            try {
                (*job_)(i);
            } catch (...) {
                std::lock_guard<std::mutex> lock(m_);
                if (!error_) error_ = std::current_exception();
            }
        }
        inside_job() = false;
    }
    void worker() {
        unsigned long long seen = 0;
        for (;;) {
            for (int spin = 0; spin < kSpins && generation_.load(std::memory_order_acquire) == seen; spin++) cpu_relax();
            if (generation_.load(std::memory_order_acquire) == seen) {
                std::unique_lock<std::mutex> lock(m_);
                sleepers_++;
                wake_.wait(lock, [&] { return generation_.load(std::memory_order_acquire) != seen; });
                sleepers_--;
            }
            seen = generation_.load(std::memory_order_acquire);
            work();
            if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                std::lock_guard<std::mutex> lock(m_);
                done_.notify_one();
            }
        }
    }

    const pid_t pid_;
    std::vector<std::thread> workers_;
    std::mutex m_, run_mutex_;
    std::condition_variable wake_, done_;
    const std::function<void(int)>* job_ = nullptr;
    int n_jobs_ = 0;
    int sleepers_ = 0;  // under m_
    std::atomic<int> pending_{0};
    std::atomic<unsigned long long> generation_{0};
    std::atomic<int> next_{0};
    std::exception_ptr error_;
};

template <class Body>
inline void parallel_for(size_t n, size_t grain, Body&& body) {
    if (n == 0) return;
    if (n <= grain) {
        body((size_t)0, n, 0);
        return;
    }
    WorkerPool& pool = WorkerPool::instance();
    const size_t chunks = std::min<size_t>((n + grain - 1) / grain, (size_t)pool.threads() * 4);
    if (chunks <= 1 || pool.threads() == 1) {
        body((size_t)0, n, 0);
        return;
    }
    pool.run((int)chunks, [&](int c) { body(n * (size_t)c / chunks, n * (size_t)(c + 1) / chunks, c); });
}

// How many chunks parallel_for(n, grain, ...) will use (for per-chunk scratch sized before the call).
inline int parallel_chunks(size_t n, size_t grain) {
    if (n <= grain) return 1;
    WorkerPool& pool = WorkerPool::instance();
    const size_t chunks = std::min<size_t>((n + grain - 1) / grain, (size_t)pool.threads() * 4);
    return (chunks <= 1 || pool.threads() == 1) ? 1 : (int)chunks;
}

// memcpy of a large block by all threads (a 14 MB primitive array is a millisecond of one core's time)
inline void parallel_copy(void* dst, const void* src, size_t bytes, size_t grain = (size_t)1 << 20) {
    parallel_for(bytes, grain, [&](size_t b, size_t e, int) { memcpy((char*)dst + b, (const char*)src + b, e - b); });
}

// The same into a buffer the device is about to read over PCIe: non-temporal stores, so that no line of it is left
// dirty in a core's cache (a 22 MB upload written by sixteen cores with plain stores took 2.7 ms of DMA instead of 0.44).
inline void stream_copy_range(char* dst, const char* src, size_t n) {
#if defined(__SSE2__)
    while (n && (reinterpret_cast<uintptr_t>(dst) & 15)) *dst++ = *src++, n--;
    for (; n >= 64; n -= 64, dst += 64, src += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 32));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 48), d);
    }
    _mm_sfence();
#endif
    memcpy(dst, src, n);
}
inline void parallel_stream_copy(void* dst, const void* src, size_t bytes) {
    parallel_for(bytes, (size_t)1 << 20, [&](size_t b, size_t e, int) { stream_copy_range((char*)dst + b, (const char*)src + b, e - b); });
}

// Big blocks are kept between uses.  A one-shot render of a 100 k-primitive scene allocates ~100 MB of arrays and frees
// them a few milliseconds later; whether malloc hands such blocks back to the kernel (munmap, then page faults on every
// page of the next frame: + 4 ms per c4 frame) depends on the history of the process's heap — here it does not.
class BlockCache {
   public:
    static constexpr size_t kMinBytes = (size_t)1 << 20, kHeader = 64, kMaxHeld = (size_t)768 << 20;
    static BlockCache& instance() {
        static BlockCache* cache = new BlockCache();  // never destroyed: blocks may be returned during static destruction
        return *cache;
    }
    void* allocate(size_t bytes) {  // bytes >= kMinBytes
        const size_t want = (bytes + kMinBytes - 1) & ~(kMinBytes - 1);
        {
            std::lock_guard<std::mutex> lock(m_);
            size_t best = free_.size();
            for (size_t i = 0; i < free_.size(); i++)
                if (free_[i].first >= want && free_[i].first <= 2 * want && (best == free_.size() || free_[i].first < free_[best].first)) best = i;
            if (best < free_.size()) {
                char* block = free_[best].second;
                held_ -= free_[best].first;
                free_[best] = free_.back();
                free_.pop_back();
                return block + kHeader;
            }
        }
        char* block = static_cast<char*>(::operator new(want + kHeader));
        *reinterpret_cast<size_t*>(block) = want;
        return block + kHeader;
    }
    void deallocate(void* p) {
        char* block = static_cast<char*>(p) - kHeader;
        const size_t capacity = *reinterpret_cast<size_t*>(block);
        {
            std::lock_guard<std::mutex> lock(m_);
            if (held_ + capacity <= kMaxHeld) {
                free_.emplace_back(capacity, block);
                held_ += capacity;
                return;
            }
        }
        ::operator delete(block);
    }

   private:
    std::mutex m_;
    std::vector<std::pair<size_t, char*>> free_;
    size_t held_ = 0;
};

// std::allocator that leaves what a resize adds untouched (not even default member initialisers run): for arrays of
// plain records that the threads fill right after — a sequential fill of 14 MB first would cost what the threads save.
template <class T>
struct NoInit : std::allocator<T> {
    static_assert(std::is_trivially_destructible<T>::value && std::is_trivially_copyable<T>::value, "plain records only");
    template <class U>
    struct rebind {
        using other = NoInit<U>;
    };
    NoInit() = default;
    template <class U>
    NoInit(const NoInit<U>&) noexcept {}
    template <class U>
    void construct(U*) noexcept {}
    template <class U, class... A>
    void construct(U* p, A&&... a) {
        ::new (static_cast<void*>(p)) U(std::forward<A>(a)...);
    }
    // blocks of a MiB and more come from, and go back to, the process-wide cache above
    T* allocate(size_t n) {
        const size_t bytes = n * sizeof(T);
        if (bytes >= BlockCache::kMinBytes) return static_cast<T*>(BlockCache::instance().allocate(bytes));
        return static_cast<T*>(::operator new(bytes));
    }
    void deallocate(T* p, size_t n) noexcept {
        if (n * sizeof(T) >= BlockCache::kMinBytes)
            BlockCache::instance().deallocate(p);
        else
            ::operator delete(p);
    }
};
template <class T>
using RawVector = std::vector<T, NoInit<T>>;

}  // namespace rtc
