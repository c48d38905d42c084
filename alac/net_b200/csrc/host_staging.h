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

namespace alacgpu {

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
    // dst[0, len) = src[0, len), split over the pool; returns when every part is done
    void copy(uint8_t *dst, const uint8_t *src, uint64_t len)
    {
        const int parts = (int)std::min<uint64_t>((uint64_t)workers_.size(), (len + kMinPart - 1) / kMinPart);
        if (parts <= 1) { memcpy(dst, src, len); return; }
        std::atomic<int> left{parts};
        std::mutex dm;
        std::condition_variable dcv;
        const uint64_t step = ((len + parts - 1) / parts + 4095) & ~4095ull;
        {
            std::lock_guard<std::mutex> g(m_);
            for (int p = 0; p < parts; p++) {
                const uint64_t lo = std::min<uint64_t>(len, step * p), hi = std::min<uint64_t>(len, step * (p + 1));
                q_.push_back([=, &left, &dm, &dcv] {
                    if (hi > lo) memcpy(dst + lo, src + lo, hi - lo);
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
