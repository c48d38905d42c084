// host_staging.h -- library-owned page-locked staging for callers whose buffers are plain pageable memory.
//
// The reference's callers hand over managed arrays (byte[] pinned with `fixed` / GCHandle, AlacContext.cs:194-195):
// pageable memory as far as CUDA is concerned.  cudaMemcpyAsync on such memory degrades to a synchronous
// bounce copy on the calling thread, which serialises the H2D / kernel / D2H overlap the pipeline is built on.
// So the library owns two small rings of page-locked slots per device: caller bytes are copied into a slot by
// a pool of host threads while the previous slot is in flight over PCIe (and the other way round for PCM), and
// nothing of the caller's is touched by the DMA engines.
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define ALACGPU_HAVE_STREAM_STORES 1
#endif

namespace alacgpu {

// dst[0, n) = src[0, n) with non-temporal stores: the PCM that leaves a ring slot for the caller's buffer is not
// read again soon, and an ordinary store would first pull every destination line into the cache (the end-to-end
// path from / to pageable memory is bound by host DRAM traffic: every byte crosses it three to four times).
inline void copy_streaming(uint8_t *dst, const uint8_t *src, uint64_t n)
{
#ifdef ALACGPU_HAVE_STREAM_STORES
    uint64_t head = (16u - (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
    if (head > n) head = n;
    if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
    const uint64_t body = n & ~(uint64_t)63;
    for (uint64_t i = 0; i < body; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 32));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 48), d);
    }
    _mm_sfence();
    if (n > body) memcpy(dst + body, src + body, n - body);
#else
    memcpy(dst, src, n);
#endif
}

// A few host threads that copy memory in parallel (one memcpy per thread saturates one core's load/store
// bandwidth, ~10 GB/s; PCIe Gen5 moves ~55 GB/s each way).
class CopyPool {
public:
    explicit CopyPool(int threads)
    {
        for (int i = 0; i < threads; i++) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (std::thread &t : workers_) t.join();
    }
    int threads() const { return (int)workers_.size(); }
    // dst[0, len) = src[0, len), split over the pool; returns when every part is done.  `streaming`: the
    // destination is written with non-temporal stores (copy_streaming)
    void copy(uint8_t *dst, const uint8_t *src, uint64_t len, bool streaming = false)
    {
        const int parts = (int)std::min<uint64_t>((uint64_t)workers_.size(), (len + kMinPart - 1) / kMinPart);
        if (parts <= 1) { if (streaming) copy_streaming(dst, src, len); else memcpy(dst, src, len); return; }
        std::atomic<int> left{parts};
        std::mutex dm;
        std::condition_variable dcv;
        const uint64_t step = ((len + parts - 1) / parts + 4095) & ~4095ull;
        {
            std::lock_guard<std::mutex> g(m_);
            for (int p = 0; p < parts; p++) {
                const uint64_t lo = std::min<uint64_t>(len, step * p), hi = std::min<uint64_t>(len, step * (p + 1));
                q_.push_back([=, &left, &dm, &dcv] {
                    if (hi > lo) { if (streaming) copy_streaming(dst + lo, src + lo, hi - lo); else memcpy(dst + lo, src + lo, hi - lo); }
                    if (left.fetch_sub(1) == 1) {
                        std::lock_guard<std::mutex> g2(dm);
                        dcv.notify_one();
                    }
                });
            }
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> lk(dm);
        dcv.wait(lk, [&] { return left.load() == 0; });
    }

private:
    static constexpr uint64_t kMinPart = 1 << 20;
    void run()
    {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (stop_ && q_.empty()) return;
                job = std::move(q_.front());
                q_.pop_front();
            }
            job();
        }
    }
    std::vector<std::thread> workers_;
    std::deque<std::function<void()>> q_;
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;
};

// kSlots page-locked buffers of `slot_bytes` each, used round-robin; slot i may be reused once ev[i] has completed.
struct PinnedRing {
    static constexpr int kSlots = 4;
    uint8_t *buf[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    bool busy[kSlots] = {false, false, false, false};
    uint64_t slot_bytes = 0;
    int next = 0;
    cudaError_t ensure(uint64_t bytes)
    {
        if (slot_bytes >= bytes) return cudaSuccess;
        release();
        for (int i = 0; i < kSlots; i++) {
            if (cudaError_t e = cudaMallocHost(reinterpret_cast<void **>(&buf[i]), bytes)) { release(); return e; }
            if (cudaError_t e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)) { release(); return e; }
        }
        slot_bytes = bytes;
        return cudaSuccess;
    }
    void release()
    {
        for (int i = 0; i < kSlots; i++) {
            if (buf[i]) cudaFreeHost(buf[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
            buf[i] = nullptr; ev[i] = nullptr; busy[i] = false;
        }
        slot_bytes = 0;
        next = 0;
    }
};

// A counter one thread advances and another waits for ("chunk c's event has been recorded"): an event must be
// RECORDED before another stream can be made to wait for it, so the recording thread publishes its progress.
class Progress {
public:
    void reset() { std::lock_guard<std::mutex> g(m_); v_ = 0; failed_ = false; }
    void set(uint64_t v) { { std::lock_guard<std::mutex> g(m_); v_ = v; } cv_.notify_all(); }
    void fail() { { std::lock_guard<std::mutex> g(m_); failed_ = true; } cv_.notify_all(); }
    bool wait_above(uint64_t v)      // false if the producer failed
    {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return failed_ || v_ > v; });
        return !failed_;
    }
private:
    std::mutex m_;
    std::condition_variable cv_;
    uint64_t v_ = 0;
    bool failed_ = false;
};

}  // namespace alacgpu
