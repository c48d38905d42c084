// K2 -- adaptive FIR predictor reconstruction (sign-LMS), in place on the
// stream-major residual planes.
//
// Replaces PredictorDecompressFirAdapt (ALACDecoder/AlacFile.cs:256-336).
//
// Mapping.  The recurrence is serial in the sample index but independent per
// (frame, channel) "stream".  One LANE owns one stream and walks its own row
// of the plane four samples at a time (one 16-byte load and one 16-byte store
// per four samples, prefetched two blocks ahead).  The work per sample is
// proportional to the stream's predictor order, so a pre-pass (k0s_order_sort)
// sorts the chunk's active streams by DESCENDING order and pads every order
// class to whole warps: a warp never mixes orders, each warp runs the code for
// exactly its order (template on the tap count, tap weights as immediates),
// and the heaviest warps are scheduled first (blocks are issued in order),
// which is what bounds the makespan when a batch has fewer streams than the
// GPU has lanes.
//
// The per-sample body is STRAIGHT-LINE code, identical for every lane:
//   * coefficients c[M] and the last M+1 outputs H[M+1] live in registers,
//     statically indexed, fully unrolled over the taps;
//   * the data-dependent early exit of the coefficient update
//     (AlacFile.cs:322) is the predicate "running error still positive" on a
//     sign-normalised error E = sign(err) * err;
//   * warm-up samples and zero residuals run the same code with E = 0 and a
//     select on the output; delta mode (order 31, :268-282) is its own class.
#pragma once
#include "alacgpu_device.cuh"
#include "alacgpu_kernels.h"
#include "lpc_tap.cuh"

namespace alacgpu {

// ---- order sort ------------------------------------------------------------------
// key(stream) = predictor order if K2 has work on it (1..31), else "inactive".  One block:
// shared-memory histogram -> descending offsets -> scatter.  perm[0..n_active) = stream ids
// (slot*2 + ch), heaviest first; perm_count[0] = n_active.
constexpr int kSortThreads = 1024;

__device__ __forceinline__ int lpc_key(const FrameDesc &d, int ch)
{
    const bool active = d.status == FS_OK && !(d.flags & FF_ESCAPE) && (ch == 0 || (d.flags & FF_STEREO)) &&
                        d.order[ch] != 0 && d.n > 1;                 // order 0: output == residual (:261-267)
    if (!active) return -1;
    // delta mode (31) has no taps: lightest class, sorted last
    return d.order[ch] == 31 ? 0 : d.order[ch];
}

// Streams that set the batch's critical path when the chunk is small (fewer frames than lanes): the
// LAST channel of a frame with a high predictor order.  Its residuals only start to appear once
// the entropy lane has finished the other channel, and its order-30 recurrence is the slowest thing
// in the pipeline, so these streams get FOUR (or eight) lanes each (lpc_lanes) instead of one.  The rest of the
// streams have slack and stay one lane per stream.
// `use_quads` packs the two thresholds: bits 0..7 = smallest order that gets four lanes on the LAST
// channel of a frame, bits 8..15 = the same for the other channel; 0 = never.
__device__ __forceinline__ bool lpc_quad(const FrameDesc &d, int ch, int key, int use_quads)
{
    const int last = (d.flags & FF_STEREO) ? 1 : 0;
    const int thr = ch == last ? (use_quads & 255) : ((use_quads >> 8) & 255);
    return thr != 0 && key >= thr;
}

// perm[0 .. n_rest): one-lane streams, heaviest order first, every order class padded with kNoStream to a
// multiple of 32 (a warp never mixes orders); perm[quad_base(n) .. + n_quad): multi-lane streams, heaviest
// first; perm_count[0] = n_rest (padded), perm_count[1] = n_quad.
constexpr uint32_t kPermPad = 32u * 32u;            // room for the padding of the 31 one-lane classes
__host__ __device__ __forceinline__ uint32_t quad_base(uint32_t n_frames) { return 2u * n_frames + kPermPad; }
constexpr uint32_t kNoStream = 0xFFFFFFFFu;

__global__ void __launch_bounds__(kSortThreads)
k0s_order_sort(const FrameDesc *__restrict__ desc, uint32_t n_frames, uint32_t *__restrict__ perm,
               uint32_t *__restrict__ perm_count, uint8_t *__restrict__ lpc_flag, const int use_quads)
{
    __shared__ uint32_t hist[64], cursor[64], padded;
    if (threadIdx.x < 64) hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n_streams = n_frames * 2u;
    for (uint32_t s = threadIdx.x; s < n_streams; s += kSortThreads) {
        const FrameDesc d = desc[s >> 1];
        const int key = lpc_key(d, (int)(s & 1u));
        if (key >= 0) atomicAdd(&hist[key + (lpc_quad(d, (int)(s & 1u), key, use_quads) ? 32 : 0)], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int key = 30; key >= 0; --key) { cursor[key] = acc; acc += (hist[key] + 31u) & ~31u; }
        cursor[31] = 0;
        perm_count[0] = padded = acc;
        acc = 0;
        for (int key = 30; key >= 0; --key) { cursor[32 + key] = quad_base(n_frames) + acc; acc += hist[32 + key]; }
        perm_count[1] = acc;
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < padded; k += kSortThreads) perm[k] = kNoStream;
    __syncthreads();
    for (uint32_t s = threadIdx.x; s < n_streams; s += kSortThreads) {
        const FrameDesc d = desc[s >> 1];
        const int key = lpc_key(d, (int)(s & 1u));
        if (key >= 0) perm[atomicAdd(&cursor[key + (lpc_quad(d, (int)(s & 1u), key, use_quads) ? 32 : 0)], 1u)] = s;
        lpc_flag[s] = key >= 0 ? 1 : 0;
    }
}

#ifndef ALACGPU_LPC_STREAMS_PER_WARP
#define ALACGPU_LPC_STREAMS_PER_WARP 32
#endif
constexpr int kK2Threads = 128;    // four LPC warps per block; blocks are issued heaviest first

__device__ __forceinline__ void st_relaxed(uint32_t *p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Fused kernel only: block until the entropy lane that produces this stream has published at
// least `need` residuals.  Bounded: a producer that never shows up (a bug, not a data
// condition) flags the lane instead of hanging the GPU.
constexpr uint32_t kSpinLimit = 1u << 20;   // x (up to 4 us of back-off sleep + one L2 round trip): ~4 s, two orders of magnitude above any legitimate wait -- only a genuine fault
// Warp-uniform on purpose: every lane polls its own stream's word, but the loop exit is a
// warp vote, so the lanes leave together (a per-lane spin loop lets the warp fall apart and
// run the tap code lane by lane afterwards).
__device__ __forceinline__ void wait_avail(const uint32_t *prog, const uint32_t need, uint32_t &avail,
                                           const bool active, bool &stalled)
{
    // Back-off: a consumer of channel B sits here for the whole of channel A's entropy decode, and every
    // poll costs issue slots (and an L1 invalidation) that the entropy lanes of the same SM need.
    uint32_t ns = 256;
    for (uint32_t spins = 0;; ++spins) {
        if (active && avail < need) avail = ld_acquire(prog);
        const bool ok = !active || avail >= need;
        if (__all_sync(0xffffffffu, ok)) break;
        if (spins >= kSpinLimit) {
            stalled = stalled || !ok;
            avail = 0xFFFFFFFFu;
            break;
        }
        __nanosleep(ns);
        ns = min(ns * 2u, 4096u);
    }
}

// All 32 lanes run this; `active` gates memory traffic only.  Every active lane's order is exactly M
// (1..30).
//   row    : the lane's row of the plane (16-byte aligned), n samples (0 if inactive)
//   nmax   : warp maximum of n
template <int M, bool kPoll, bool kPublish>
__device__ __noinline__ bool lpc_warp(int32_t *row, const int n, const int nmax, const int rss, const int q,
                                      const int16_t *__restrict__ coef16, const bool active,
                                      const uint32_t *prog, uint32_t *done)
{
    int32_t c[M], H[M + 1];                        // H[j] = o[i-1-j]; H[M] is the base o[i-1-M]
#pragma unroll
    for (int j = 0; j < M; j++) c[j] = active ? (int32_t)coef16[j] : 0;
#pragma unroll
    for (int j = 0; j <= M; j++) H[j] = 0;

    const int32_t rnd = (int32_t)(1u << ((q - 1) & 31));            // :306 (quant 0 -> 1 << 31)
    // sign * ((val*sign) >> quant) = (|val| + r) >> quant with r = 0 for a positive error and
    // r = 2^quant - 1 for a negative one (arithmetic shift of the negated magnitude, :328-329)
    const uint32_t rneg = (1u << q) - 1u;
    const int sh = (32 - rss) & 31;
    int4 *row4 = reinterpret_cast<int4 *>(row);
    const int nblk = (n + 3) >> 2, nblk_max = (nmax + 3) >> 2;

    // residual blocks are fetched two blocks (eight samples) ahead
    uint32_t avail = kPoll ? 0u : 0xFFFFFFFFu;
    bool stalled = false;
    if (kPoll) wait_avail(prog, (uint32_t)min(nblk, 10) * 4u, avail, active, stalled);
    int4 cur = active ? __ldcg(row4) : make_int4(0, 0, 0, 0);
    int4 nx1 = (active && nblk > 1) ? __ldcg(row4 + 1) : make_int4(0, 0, 0, 0);
    H[0] = cur.x;                                                   // first sample always copies (:259-260)
    for (int b = 0; b < nblk_max; b++) {
        // every 8 blocks: make sure the 32 residuals after the ones already granted are there
        if (kPoll && (b & 7) == 7) wait_avail(prog, (uint32_t)min(nblk, b + 11) * 4u, avail, active, stalled);
        const int4 nx2 = (active && b + 2 < nblk) ? __ldcg(row4 + b + 2) : make_int4(0, 0, 0, 0);
        // The four samples of a block run through ONE copy of the tap code (the body is ~11 M
        // instructions; unrolling it four times would overflow the instruction cache), so the
        // block's residuals / outputs are moved with selects instead of static indices.
        int32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
#pragma unroll 1
        for (int u = 0; u < 4; u++) {
            const int i = b * 4 + u;
            const int32_t e = u == 0 ? cur.x : (u == 1 ? cur.y : (u == 2 ? cur.z : cur.w));
            int32_t o = H[0];
            if (i != 0) {
                const bool main = i > M;                                    // warm-up covers i = 1..M (:284-293)
                const int32_t base = H[M];                                  // o[i-1-M]
                const int32_t nsg = e < 0 ? 1 : -1;                         // -sign(err)
                const int32_t sgbase = e < 0 ? (int32_t)(0u - (uint32_t)base) : base;
                int32_t E = main ? (e < 0 ? (int32_t)(0u - (uint32_t)e) : e) : 0;   // sign(err) * err
                const uint32_t r = e < 0 ? rneg : 0u;
                uint32_t acc = 0;
                lpc_taps<M, M - 1>(c, H, E, acc, nsg, sgbase, r, (uint32_t)q);
                const int32_t sum = (int32_t)(acc * (uint32_t)nsg);         // sum of (buf[b+order-j]-buf[b])*coef[j]
                int32_t v = (int32_t)((uint32_t)rnd + (uint32_t)sum) >> q;  // :306-307
                v = (int32_t)((uint32_t)v + (uint32_t)base + (uint32_t)e);  // :308
                const int32_t w = (int32_t)((uint32_t)H[0] + (uint32_t)e);  // warm-up (:288)
                const int32_t x = main ? v : w;
                o = (int32_t)((uint32_t)x << sh) >> sh;                     // :309-310
#pragma unroll
                for (int j = M; j > 0; --j) H[j] = H[j - 1];
                H[0] = o;
            }
            o0 = u == 0 ? o : o0; o1 = u == 1 ? o : o1; o2 = u == 2 ? o : o2; o3 = u == 3 ? o : o3;
        }
        if (active && b < nblk) row4[b] = make_int4(o0, o1, o2, o3);
        if (kPublish && (b & 7) == 7) {              // hand-off to the pack warps, every 32 samples
            __threadfence();
            if (active && b < nblk) st_relaxed(done, (uint32_t)(b + 1) * 4u);
        }
        cur = nx1;
        nx1 = nx2;
    }
    if (kPublish) {
        __threadfence();
        if (active) st_relaxed(done, 0xFFFFFFFFu);
    }
    return stalled;
}

// Delta mode (order 31, AlacFile.cs:268-282): o[i] = o[i-1] + e[i], no taps.
template <bool kPoll, bool kPublish>
__device__ __noinline__ bool lpc_delta(int32_t *row, const int n, const int nmax, const int rss, const bool active,
                                       const uint32_t *prog, uint32_t *done)
{
    const int sh = (32 - rss) & 31;
    int4 *row4 = reinterpret_cast<int4 *>(row);
    const int nblk = (n + 3) >> 2, nblk_max = (nmax + 3) >> 2;
    uint32_t avail = kPoll ? 0u : 0xFFFFFFFFu;
    bool stalled = false;
    int32_t prev = 0;
    for (int b = 0; b < nblk_max; b++) {
        if (kPoll && (b & 7) == 0) wait_avail(prog, (uint32_t)min(nblk, b + 8) * 4u, avail, active, stalled);
        int4 v = (active && b < nblk) ? __ldcg(row4 + b) : make_int4(0, 0, 0, 0);
        if (b == 0) prev = v.x;
        else prev = v.x = (int32_t)((uint32_t)(prev + v.x) << sh) >> sh;
        prev = v.y = (int32_t)((uint32_t)(prev + v.y) << sh) >> sh;
        prev = v.z = (int32_t)((uint32_t)(prev + v.z) << sh) >> sh;
        prev = v.w = (int32_t)((uint32_t)(prev + v.w) << sh) >> sh;
        if (active && b < nblk) row4[b] = v;
        if (kPublish && (b & 7) == 7) {
            __threadfence();
            if (active && b < nblk) st_relaxed(done, (uint32_t)(b + 1) * 4u);
        }
    }
    if (kPublish) {
        __threadfence();
        if (active) st_relaxed(done, 0xFFFFFFFFu);
    }
    return stalled;
}

// ---- four (or eight) lanes per stream ---------------------------------------------------------------------
// Lane r (0..3) of a quad owns taps j = r*T + t, t < T (T*4 >= order); eight streams per warp.  The
// reference's early-exit loop over the taps (AlacFile.cs:322-331: newest coefficient index first, stop
// when the running error changes sign) becomes a prefix problem: every tap's step
// ((|d| + r) >> q) * (order - p) is computed unconditionally, the quad scans the steps in loop order
// (lane 3's taps first), and tap p updates iff the error minus the steps BEFORE it is still positive.
// Steps are clamped to 2^25 > max |error| (|e| <= 2^24 for rss <= 25), so the sums cannot wrap and a
// clamped step still ends the loop.  Surplus taps (j >= order) carry weight 0, coefficient 0 and an
// unreachable threshold, so they contribute nothing and never update.
//
// L = 8 lanes per stream (four streams per warp, T*8 >= order) is the same code with one more level in the scan
// and in the sum: for a batch so small that every warp has a scheduler to itself, a warp's time per sample
// is its instruction count (one instruction per two cycles), and half the taps per lane is ~40 % fewer
// instructions (`use_quads` bit 16, chosen per chunk by the runtime).
template <int L, int T, bool kPoll, bool kPublish>
__device__ __noinline__ bool lpc_lanes(int32_t *row, const int n, const int nmax, const int rss, const int ord,
                                       const int q, const int16_t *__restrict__ coef16, const bool active,
                                       int32_t *ring /* this stream's column of a [32][8] shared ring */,
                                       const uint32_t *prog, uint32_t *done)
{
    static_assert(L == 4 || L == 8, "four or eight lanes per stream");
    const int lane = threadIdx.x & 31;
    const int r4 = lane & (L - 1);
    constexpr int32_t kClamp = 1 << 25;
    int32_t c[T], H[T], wgt[T], thr[T];
#pragma unroll
    for (int t = 0; t < T; t++) {
        const int j = r4 * T + t;
        const bool valid = active && j < ord;
        c[t] = valid ? (int32_t)coef16[j] : 0;
        wgt[t] = valid ? ord - j : 0;                   // (order - p), AlacFile.cs:329
        thr[t] = valid ? 0 : 0x7fffffff;
        H[t] = 0;
        asm volatile("" : "+r"(wgt[t]), "+r"(thr[t]));  // keep as register operands
    }
    const int32_t rnd = (int32_t)(1u << ((q - 1) & 31));            // :306 (quant 0 -> 1 << 31)
    const uint32_t rneg = (1u << q) - 1u;                           // see lpc_warp
    const int sh = (32 - rss) & 31;
    int4 *row4 = reinterpret_cast<int4 *>(row);
    const int nblk = (n + 3) >> 2, nblk_max = (nmax + 3) >> 2;

    uint32_t avail = kPoll ? 0u : 0xFFFFFFFFu;
    bool stalled = false;
    if (kPoll) wait_avail(prog, (uint32_t)min(nblk, 10) * 4u, avail, active, stalled);
    int4 cur = active ? __ldcg(row4) : make_int4(0, 0, 0, 0);
    int4 nx1 = (active && nblk > 1) ? __ldcg(row4 + 1) : make_int4(0, 0, 0, 0);
    int32_t prev = cur.x;                                           // o[i-1]; first sample always copies (:259-260)
    if (r4 == 0) { H[0] = prev; ring[0] = prev; }
    __syncwarp();
    for (int b = 0; b < nblk_max; b++) {
        if (kPoll && (b & 7) == 7) wait_avail(prog, (uint32_t)min(nblk, b + 11) * 4u, avail, active, stalled);
        const int4 nx2 = (active && b + 2 < nblk) ? __ldcg(row4 + b + 2) : make_int4(0, 0, 0, 0);
        // The block's four samples are unrolled (unlike lpc_warp's: a quad lane has at most 8 taps, so four
        // copies of the body still fit the instruction cache): no selects on the block's residuals / outputs,
        // and the history shift turns into register renaming.  A lone warp issues one instruction per two
        // cycles, and these streams are the critical path of a small batch.
        int32_t ob[4] = {0, 0, 0, 0};
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = b * 4 + u;
            const int32_t e = u == 0 ? cur.x : (u == 1 ? cur.y : (u == 2 ? cur.z : cur.w));
            int32_t o = prev;
            if (i != 0) {
                const int32_t base = ring[((uint32_t)(i - 1 - ord) & 31u) * 8];      // o[i-1-ord]
                const bool main = i > ord;                                  // warm-up covers i = 1..ord (:284-293)
                const int32_t nsg = e < 0 ? 1 : -1;                         // -sign(err)
                const int32_t sgbase = e < 0 ? (int32_t)(0u - (uint32_t)base) : base;
                const int32_t E0 = main ? (e < 0 ? (int32_t)(0u - (uint32_t)e) : e) : 0;   // sign(err) * err
                const uint32_t rr = e < 0 ? rneg : 0u;
                // pass 1: dot product and the unconditional steps of this lane's taps
                uint32_t acc = 0;
                int32_t dp[T], st[T];
                int32_t mine = 0;
#pragma unroll
                for (int t = T - 1; t >= 0; --t) {
                    dp[t] = (int32_t)((uint32_t)H[t] * (uint32_t)nsg + (uint32_t)sgbase);
                    acc += (uint32_t)c[t] * (uint32_t)dp[t];
                    const uint32_t mag = (uint32_t)abs(dp[t]) + rr;
                    st[t] = (int32_t)min((mag >> q) * (uint32_t)wgt[t], (uint32_t)kClamp);
                    mine += st[t];
                }
                // steps taken before this lane's first tap: the totals of the group's HIGHER lanes
                const int32_t t1 = __shfl_down_sync(0xffffffffu, mine, 1, L);
                const int32_t a1 = mine + (r4 < L - 1 ? t1 : 0);
                const int32_t t2 = __shfl_down_sync(0xffffffffu, a1, 2, L);
                int32_t incl = a1 + (r4 < L - 2 ? t2 : 0);
                if (L == 8) {
                    const int32_t t4 = __shfl_down_sync(0xffffffffu, incl, 4, L);
                    incl += r4 < 4 ? t4 : 0;
                }
                // (warm-up samples read a base that is not there yet: their steps are garbage, never used)
                int32_t rem = main ? E0 - (incl - mine) : -1;
                // pass 2: sign-LMS update of the taps the reference's loop would have reached
#pragma unroll
                for (int t = T - 1; t >= 0; --t) {
                    const int32_t sg = max(min(dp[t], 1), -1);
                    c[t] -= rem > thr[t] ? sg : 0;
                    rem -= st[t];
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1, L);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2, L);
                if (L == 8) acc += __shfl_xor_sync(0xffffffffu, acc, 4, L);
                const int32_t sum = (int32_t)(acc * (uint32_t)nsg);         // sum of (buf[b+order-j]-buf[b])*coef[j]
                int32_t v = (int32_t)((uint32_t)rnd + (uint32_t)sum) >> q;  // :306-307
                v = (int32_t)((uint32_t)v + (uint32_t)base + (uint32_t)e);  // :308
                const int32_t w = (int32_t)((uint32_t)prev + (uint32_t)e);  // warm-up (:288)
                const int32_t x = main ? v : w;
                o = (int32_t)((uint32_t)x << sh) >> sh;                     // :309-310
                // history: every lane shifts by one tap; lane r takes lane r-1's oldest value
                const int32_t from_below = __shfl_up_sync(0xffffffffu, H[T - 1], 1, L);
#pragma unroll
                for (int t = T - 1; t > 0; --t) H[t] = H[t - 1];
                H[0] = r4 == 0 ? o : from_below;
                if (r4 == 0) ring[((uint32_t)i & 31u) * 8] = o;
                __syncwarp();
                prev = o;
            }
            ob[u] = o;
        }
        if (active && r4 == 0 && b < nblk) row4[b] = make_int4(ob[0], ob[1], ob[2], ob[3]);
        if (kPublish && (b & 7) == 7) {              // hand-off to the pack warps, every 32 samples
            __threadfence();
            if (active && r4 == 0 && b < nblk) st_relaxed(done, (uint32_t)(b + 1) * 4u);
        }
        cur = nx1;
        nx1 = nx2;
    }
    if (kPublish) {
        __threadfence();
        if (active && r4 == 0) st_relaxed(done, 0xFFFFFFFFu);
    }
    return stalled;
}

// One LPC warp.  The first warps take the four-lane (eight-lane) streams, eight (four) per warp, the
// others 32 one-lane streams of ONE order each.  hist_warp: this warp's 4 KB of shared memory (the
// multi-lane warps' output rings).
template <bool kPoll, bool kPublish>
__device__ __forceinline__ void lpc_role(const ChunkArgs &a, uint32_t warp, int32_t *hist_warp)
{
    const int lane = threadIdx.x & 31;
    const uint32_t n_active = a.perm_count[0], n_quad = a.perm_count[1];
    const bool wide = (a.use_quads >> 16) & 1;                  // eight lanes per stream instead of four
    const uint32_t per_warp = wide ? 4u : 8u;                    // multi-lane streams of one warp
    const uint32_t quad_warps = (n_quad + per_warp - 1u) / per_warp;
    const bool quad = warp < quad_warps;
    if (!quad) warp -= quad_warps;
    const uint32_t group = wide ? (uint32_t)(lane >> 3) : (uint32_t)(lane >> 2);
    const uint32_t idx = quad ? warp * per_warp + group : warp * 32u + (uint32_t)lane;
    if (!quad && warp * 32u >= n_active) return;
    uint32_t sid = kNoStream;
    if (quad ? idx < n_quad : idx < n_active) sid = quad ? a.perm[quad_base(a.n) + idx] : a.perm[idx];
    // (checked build) list sizes and stream ids inside the chunk, rows inside the planes
    if (!ALACGPU_CHECK(a.check, n_active <= 2u * a.n + kPermPad && n_quad <= 2u * a.n, CK_LIST)) sid = kNoStream;
    if (sid != kNoStream && !(ALACGPU_CHECK(a.check, sid < 2u * a.n, CK_LIST) &&
                              ALACGPU_CHECK(a.check, ((uint64_t)sid + 1u) * a.ns * 4u <= a.plane_bytes, CK_PLANE)))
        sid = kNoStream;
    const bool active = sid != kNoStream;
    int n = 0, rss = 32, ord = 0, q = 0;
    const int16_t *coef16 = nullptr;
    int32_t *row = nullptr;
    const uint32_t *prog = nullptr;
    uint32_t *done = nullptr;
    uint64_t f = 0;
    if (active) {
        f = a.f0 + (sid >> 1);
        const int ch = (int)(sid & 1u);
        const FrameDesc d = a.desc[f];
        n = d.n; rss = d.rss; ord = d.order[ch]; q = d.quant[ch];
        coef16 = a.coefs[f].c[ch];
        row = a.planes + (uint64_t)sid * a.ns;
        prog = a.progress + sid;
        done = a.lpc_done + sid;
    }
    const int maxo = __reduce_max_sync(0xffffffffu, ord);      // one-lane warps: THE order of the warp (31 = delta mode)
    const int nmax = __reduce_max_sync(0xffffffffu, n);
    bool stalled = false;
    if (quad) {
        int32_t *ring = hist_warp + group;
        if (!active) ord = 1;
#define ALACGPU_LPCL(LL, TT) stalled = lpc_lanes<LL, TT, kPoll, kPublish>(row, n, nmax, rss, ord, q, coef16, active, ring, prog, done)
        if (wide) {
            if (maxo <= 8) ALACGPU_LPCL(8, 1);
            else if (maxo <= 16) ALACGPU_LPCL(8, 2);
            else if (maxo <= 24) ALACGPU_LPCL(8, 3);
            else ALACGPU_LPCL(8, 4);
        } else if (maxo <= 8) ALACGPU_LPCL(4, 2);
        else if (maxo <= 12) ALACGPU_LPCL(4, 3);
        else if (maxo <= 16) ALACGPU_LPCL(4, 4);
        else if (maxo <= 20) ALACGPU_LPCL(4, 5);
        else if (maxo <= 24) ALACGPU_LPCL(4, 6);
        else if (maxo <= 28) ALACGPU_LPCL(4, 7);
        else ALACGPU_LPCL(4, 8);
#undef ALACGPU_LPCL
        if (kPoll && active && stalled && (lane & (wide ? 7 : 3)) == 0) { a.desc[f].status = FS_INTERNAL; atomicAdd(a.faults, 1u); }
        return;
    }
#define ALACGPU_LPC(MM) case MM: stalled = lpc_warp<MM, kPoll, kPublish>(row, n, nmax, rss, q, coef16, active, prog, done); break
    switch (maxo) {
        ALACGPU_LPC(1); ALACGPU_LPC(2); ALACGPU_LPC(3); ALACGPU_LPC(4); ALACGPU_LPC(5); ALACGPU_LPC(6);
        ALACGPU_LPC(7); ALACGPU_LPC(8); ALACGPU_LPC(9); ALACGPU_LPC(10); ALACGPU_LPC(11); ALACGPU_LPC(12);
        ALACGPU_LPC(13); ALACGPU_LPC(14); ALACGPU_LPC(15); ALACGPU_LPC(16); ALACGPU_LPC(17); ALACGPU_LPC(18);
        ALACGPU_LPC(19); ALACGPU_LPC(20); ALACGPU_LPC(21); ALACGPU_LPC(22); ALACGPU_LPC(23); ALACGPU_LPC(24);
        ALACGPU_LPC(25); ALACGPU_LPC(26); ALACGPU_LPC(27); ALACGPU_LPC(28); ALACGPU_LPC(29); ALACGPU_LPC(30);
        case 31: stalled = lpc_delta<kPoll, kPublish>(row, n, nmax, rss, active, prog, done); break;
        default: break;
    }
#undef ALACGPU_LPC
    if (kPoll && active && stalled) { a.desc[f].status = FS_INTERNAL; atomicAdd(a.faults, 1u); }   // never expected: see wait_avail
}

}  // namespace alacgpu
