// Host-only exercise of the staging helpers of libalacgpu (alac/net_b200/csrc/host_staging.h): the copy pool
// that moves caller bytes in and out of the page-locked rings, and the Progress hand-over between the stager,
// the issuer and the drainer threads.  Built with -fsanitize=thread when the toolchain has it
// (tests/test_host_staging_cpu.py), so a data race in either fails the CPU test tier.  No CUDA call is made.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "host_staging.h"

using namespace alacgpu;

static int check_copy_pool()
{
    CopyPool pool(6);
    std::mt19937_64 rng(7);
    for (int round = 0; round < 40; round++) {
        const uint64_t len = round < 4 ? (uint64_t)round : (rng() % (24u << 20)) + 1;
        std::vector<uint8_t> src(len + 64), dst(len + 64, 0xEE);
        for (uint64_t i = 0; i < src.size(); i++) src[i] = (uint8_t)(rng() >> 13);
        const uint64_t so = rng() % 64, doff = rng() % 64;
        const uint64_t n = len > std::max(so, doff) ? len - std::max(so, doff) : 0;
        pool.copy(dst.data() + doff, src.data() + so, n, /*streaming=*/(round & 1) != 0);
        for (uint64_t i = 0; i < n; i++)
            if (dst[doff + i] != src[so + i]) { printf("copy mismatch at %llu of %llu\n", (unsigned long long)i, (unsigned long long)n); return 1; }
        for (uint64_t i = 0; i < doff; i++) if (dst[i] != 0xEE) { printf("wrote before the destination\n"); return 1; }
        for (uint64_t i = doff + n; i < dst.size(); i++) if (dst[i] != 0xEE) { printf("wrote past the destination\n"); return 1; }
    }
    // two callers at once (the stager and the drainer share the pool)
    std::vector<uint8_t> a(9 << 20, 1), b(9 << 20, 0), c(7 << 20, 2), d(7 << 20, 0);
    std::thread t1([&] { for (int k = 0; k < 8; k++) pool.copy(b.data(), a.data(), a.size()); });
    std::thread t2([&] { for (int k = 0; k < 8; k++) pool.copy(d.data(), c.data(), c.size()); });
    t1.join();
    t2.join();
    if (b != a || d != c) { printf("concurrent copies differ\n"); return 1; }
    return 0;
}

// stager -> issuer -> drainer: chunk c may only be launched once its event is "recorded", and drained once it is
// "launched"; a failure upstream must release everyone downstream
static int check_progress()
{
    const uint64_t chunks = 2000;
    Progress staged, launched;
    std::vector<int> payload(chunks, 0), order;
    std::thread stager([&] { for (uint64_t c = 0; c < chunks; c++) { payload[c] = (int)c + 1; staged.set(c + 1); } });
    std::thread drainer([&] {
        for (uint64_t c = 0; c < chunks; c++) {
            if (!launched.wait_above(c)) return;
            order.push_back(payload[c]);
        }
    });
    for (uint64_t c = 0; c < chunks; c++) {
        if (!staged.wait_above(c)) { printf("stager reported a failure\n"); return 1; }
        if (payload[c] != (int)c + 1) { printf("chunk %llu launched before it was staged\n", (unsigned long long)c); return 1; }
        payload[c] = -payload[c];
        launched.set(c + 1);
    }
    stager.join();
    drainer.join();
    if (order.size() != chunks) { printf("drainer saw %zu chunks\n", order.size()); return 1; }
    for (uint64_t c = 0; c < chunks; c++) if (order[c] != -((int)c + 1)) { printf("drained before launched\n"); return 1; }
    // failure path
    Progress p;
    p.reset();
    std::thread waiter([&] { if (p.wait_above(5)) printf("wait returned success after a failure\n"); });
    p.set(3);
    p.fail();
    waiter.join();
    return 0;
}

int main()
{
    if (check_copy_pool()) return 1;
    if (check_progress()) return 1;
    printf("staging ok\n");
    return 0;
}
