// K1 -- adaptive Golomb-Rice entropy decode.
//
// Replaces Readbits/Readbit/Unreadbits (ALACDecoder/AlacFile.cs:101-152),
// CountLeadingZeros (:154-191), EntropyDecodeValue (:193-212) and
// EntropyRiceDecode (:214-252).
//
// Mapping.  A frame's Rice streams are one serial chain: every symbol's
// length depends on the running history, and channel B starts at the bit
// where channel A ends (AlacFile.cs:643 then :653 share one cursor).  So the
// unit of parallelism is the FRAME: one lane per frame, channel A then B, and
// the stage's run time is (symbols per frame) x (cycles per symbol) whatever
// the batch size.  Everything here serves a short per-symbol path:
//
//   * Lock step over the OUTPUT index: a lane inside a zero run emits its
//     zeros one per step instead of jumping ahead (AlacFile.cs:238-245 writes
//     them in a burst), so the common case is one straight-line block for
//     every lane and four residuals leave the lane as one 16-byte store into
//     its own row of the stream-major plane.
//   * The operands that depend on the field just read (new cursor, new
//     history) are one or two instructions after it: the field is
//     (w >> (p - k)) & m with p = bfind(~w); value + signModifier = A + max(e,1)
//     with A = x*m + signModifier - 1 computed beside it; the history is
//     max(e,1)*mult + (A*mult + h - ((h*mult) >> 9)).
//   * The rare paths -- the raw field after nine 1-bits (:198-202) and the
//     zero-run length symbol after a small history (:231-249) -- sit in one
//     plain divergent block.  (Measured alternatives, all byte-exact: a warp
//     vote in front of the block costs ~100 cycles per sample in WARPSYNC/VOTE;
//     a per-lane "one field per iteration" state machine without lock step
//     needs twice the instructions per sample; an 8x unrolled body misses the
//     instruction cache.  See DESIGN.md section 3.)
//   * bitstream: each lane owns a 256-byte ring in shared memory, filled by
//     16-byte cp.async copies every eight samples (no register ever waits on
//     HBM); the cursor keeps two byte-swapped words in registers plus one
//     prefetched word, so the 32-bit window at the cursor is ONE funnel shift.
//   * k <= 22 always ((history >> 9) + 3 < 2^23; zero-run k <= 16), so prefix +
//     terminator + k bits fit the 32-bit window.  "Read k bits, un-read one if
//     the value is <= 1" (AlacFile.cs:205-210) is "consume k-1 bits".
//   * CountLeadingZeros' clz(0) == 40 quirk (AlacFile.cs:190) is kept in the
//     zero-run k, the only place a zero argument can reach it.
//
// Error policy (shared with the oracle): the cursor only moves forward, so
// "some symbol ended past the frame's last bit" is decided once from the
// final cursor (OVERRUN outranks a HISTORY / RUN_OVERFLOW fault).  The arena
// carries enough tail padding for a lane that runs past its frame.
#pragma once
#include "alacgpu_device.cuh"
#include "alacgpu_kernels.h"

namespace alacgpu {

constexpr int kRingChunks = 16;              // 256 B of bitstream per lane
constexpr int kRingBytes = kRingChunks * 16;
constexpr int kFlushEvery = 8;               // samples between ring top-ups
constexpr int kK1Threads = 128;

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gptr)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t smem_addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_addr) : "memory");
    return v;
}
// position of the most significant 1 bit (31 - clz), -1 for 0
__device__ __forceinline__ int flo(uint32_t v)
{
    return 31 - __clz((int)v);
}

// Bit cursor over the lane's ring.  A lane consumes at most 59 bits per sample (9 ones + 25
// raw bits, plus a zero-run symbol of 9 + 16), i.e. < 4 chunks per period of eight samples,
// and reads two words ahead of its cursor: everything it touches during a period lies within
// chunk(cursor at the previous top-up) + 4 + 4 + 1 < kRingChunks and was requested at least
// one period earlier, so the wait for the PREVIOUS period's copies is normally free.
struct BitCursor {
    const uint8_t *base;    // 16-byte aligned global address of chunk 0
    uint32_t ring;          // shared-space byte address of this lane's ring (256-byte aligned)
    const uint32_t *ringw;  // the same ring as 64 words (plain pointer: the per-sample refill below is an
                            // ordinary predicated LDS -- an `asm volatile` there makes the compiler
                            // re-converge the whole warp (WARPSYNC.ALL) after every divergent block)
    uint32_t rw;            // index (mod 64) of the next ring word to prefetch into `nn`
    uint32_t cur, nxt;      // byte-swapped words holding bits [32*w, 32*w+64) at the cursor's word w
    uint32_t nn;            // raw word w+2
    uint32_t off;           // cursor bit within `cur`, 0..31
    uint32_t words;         // words entered since init (cursor word = word0 + words)
    uint32_t word0, off0;   // cursor at init
    uint32_t filled;        // chunks [0, filled) have been requested

    __device__ __forceinline__ void top_up()
    {
        const uint32_t want = ((word0 + words) >> 2) + kRingChunks;
        while (filled < want) {
            cp_async16(ring + ((filled & (kRingChunks - 1)) << 4), base + ((uint64_t)filled << 4));
            ++filled;
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void init(const uint8_t *arena, uint64_t abs_bit, const uint8_t *ring_ptr)
    {
        const uint64_t byte = abs_bit >> 3;
        base = arena + (byte & ~15ull);
        ring = (uint32_t)__cvta_generic_to_shared(ring_ptr);
        ringw = reinterpret_cast<const uint32_t *>(ring_ptr);
        const uint32_t pos = (uint32_t)(byte & 15) * 8u + (uint32_t)(abs_bit & 7);
        word0 = pos >> 5;
        off = off0 = pos & 31;
        words = 0;
        filled = 0;
        top_up();
        cp_async_wait<0>();
        cur = bswap32(lds32(ring + ((word0 * 4u) & (kRingBytes - 1))));
        nxt = bswap32(lds32(ring + (((word0 + 1) * 4u) & (kRingBytes - 1))));
        nn = lds32(ring + (((word0 + 2) * 4u) & (kRingBytes - 1)));
        rw = (word0 + 3) & (kRingBytes / 4 - 1);
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(nxt, cur, off); }

    // move the cursor to bit t (0..63) of the current word pair.  The word prefetched here was
    // requested two top-ups ago (see above), so the load needs no ordering against the current
    // period's cp.async wait beyond the compiler barrier that wait already is.
    __device__ __forceinline__ void seek(uint32_t t)
    {
        const bool rf = t >= 32u;
        off = t & 31u;
        cur = rf ? nxt : cur;
        // nxt = rf ? bswap(nn) : nxt in one PRMT: selector 0x0123 reverses nn, 0x7654 passes nxt
        nxt = __byte_perm(nn, nxt, rf ? 0x0123u : 0x7654u);
        words += rf ? 1u : 0u;
        if (rf) nn = ringw[rw];
        rw = (rw + (rf ? 1u : 0u)) & (kRingBytes / 4 - 1);
    }
    __device__ __forceinline__ uint32_t consumed() const { return words * 32u + off - off0; }
};

// Progress hand-off to the LPC warps of the fused kernel (k12_decode.cu): a lane publishes how
// many residuals of its stream are in the plane (every 32 samples: fence, then a relaxed
// store the consumer reads with ld.acquire) and 0xFFFFFFFF when the channel is complete.
__device__ __forceinline__ void publish(uint32_t *p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
constexpr uint32_t kStreamDone = 0xFFFFFFFFu;

// One block of 128 threads = 4 entropy warps.  `block` is the index among the entropy blocks;
// ring_smem: kRingBytes * kK1Threads bytes, 256-byte aligned.
template <bool kPublish>
__device__ __forceinline__ void entropy_block(const ChunkArgs &a, const int lanes_log2, const uint32_t block,
                                              uint8_t *ring_smem)
{
    const int lane = threadIdx.x & 31;
    const int S = 1 << lanes_log2;
    const uint32_t gw = (block * kK1Threads + threadIdx.x) >> 5;
    const uint32_t slot = gw * (uint32_t)S + (uint32_t)lane;
    // Every lane stays in the loops (warp-uniform trip counts); a lane without work runs with
    // n == 0 and commits nothing.
    bool work = lane < S && slot < a.n;
    const uint64_t f = a.f0 + (work ? slot : 0u);
    const FrameDesc d = a.desc[f];
    work = work && d.status == FS_OK && !(d.flags & FF_ESCAPE);   // escape frames are read directly by K3
    const FrameRef ref = a.refs[f];
    const TrackCfg cfg = a.cfgs[ref.track];
    int n = work ? (int)d.n : 0;
    const int rss = d.rss;
    const int kmod = cfg.rice_kmodifier;
    const uint32_t kmask = (1u << kmod) - 1u;                // AlacFile.cs:483,:643
    const int ech = (d.flags & FF_STEREO) ? 2 : 1;
    const int ech_max = __reduce_max_sync(0xffffffffu, n ? ech : 0);
    const int nmax = __reduce_max_sync(0xffffffffu, n);

    BitCursor br;
    br.init(a.arena, work ? ref.off * 8ull + d.data_bit : 0ull, ring_smem + threadIdx.x * (uint32_t)kRingBytes);

    // Loop counters are kept in ordinary (per-thread) registers on purpose: if they live in the
    // uniform datapath, ptxas must re-converge the warp (WARPSYNC.ALL, ~50 cycles) after the
    // divergent rare-path block of EVERY sample before it may touch them again.
    int lane_opaque;
    asm("mov.u32 %0, %1;" : "=r"(lane_opaque) : "r"(lane));
    const int zero = lane_opaque - lane;

    uint8_t status = FS_OK;
    for (int c = 0; c < ech_max; c++) {
        int4 *row = reinterpret_cast<int4 *>(a.planes + ((uint64_t)(work ? slot : 0u) * 2u + (uint32_t)c) * a.ns);
        const uint32_t mult = (uint32_t)((int32_t)d.rice_mod[c & 1] * (cfg.rice_history_mult / 4));   // :483
        int nc = c < ech ? n : 0;                // samples this lane decodes in this channel
        int32_t h = cfg.rice_initial_history;    // :216
        uint32_t sm1 = 0xFFFFFFFFu;              // signModifier - 1
        uint32_t zcnt = 0;                       // zeros of the current run still to emit
        int k = min(flo((uint32_t)((h >> 9) + 3)), kmod);        // :221-222
        uint32_t m0 = (1u << k) - 1u;
        uint32_t *prog = a.progress + ((uint64_t)(work ? slot : 0u) * 2u + (uint32_t)c);

        for (int i0 = zero; i0 < nmax; i0 += kFlushEvery) {
            br.top_up();
            cp_async_wait<1>();                  // everything but the group just committed
#pragma unroll 1
            for (int i4 = i0; i4 < i0 + kFlushEvery; i4 += 4) {
              int32_t out[4];
#pragma unroll
              for (int u = 0; u < 4; u++) {
                const int i = i4 + u;
                // ---- common case, one straight-line block: every lane evaluates the Rice symbol at
                // its cursor; lanes inside a zero run or past their last sample commit nothing ----
                const bool live = zcnt == 0 && i < nc;
                const uint32_t w = br.peek();
                const int p = flo(~w);                           // bit index of the first 0 bit
                const uint32_t e = (w >> ((p - k) & 31)) & m0;   // k bits after the terminator (:205)
                const uint32_t em = max(e, 1u);
                const uint32_t A = (uint32_t)(31 - p) * m0 + sm1;                     // :206, :224
                uint32_t dv = A + em;                                                 // x*m + max(e,1) - 1 + signModifier
                const bool ok = live && w < 0xFF800000u;         // fewer than nine 1 bits (:198)
                // Rice consumes x + k bits, one more if e >= 2 (:210)
                br.seek(br.off + (ok ? (uint32_t)(31 + k - p) + (e >= 2u ? 1u : 0u) : 0u));
                const int32_t hb = h - ((int32_t)((uint32_t)h * mult) >> 9);
                const int32_t hn = (int32_t)(em * mult + (A * mult + (uint32_t)hb));
                h = ok ? (dv > 0xFFFFu ? 0xFFFF : hn) : h;                            // :229
                int32_t val = (int32_t)(dv >> 1) ^ -(int32_t)(dv & 1u);               // :225-226
                val = ok ? val : 0;
                sm1 = ok ? 0xFFFFFFFFu : sm1;
                zcnt -= (zcnt != 0) ? 1u : 0u;                                        // :240-243, one zero per step
                // ---- rare per lane, but some lane of the warp needs it every few samples: the
                // nine-ones escape (:198-202) and the zero-run symbol after a small history
                // (:231-249).  A plain divergent branch: a warp vote in front of it costs more
                // (vote + re-synchronisation, ~100 cycles per sample measured) than it saves. ----
                const bool need = live && (!ok || h < 128);
                {
                    if (__builtin_expect(need, 0)) {
                        if (!ok) {
                            br.seek(br.off + 9u);
                            dv = (br.peek() >> (32 - rss)) + (sm1 + 1u);
                            br.seek(br.off + (uint32_t)rss);
                            val = (int32_t)(dv >> 1) ^ -(int32_t)(dv & 1u);
                            h = dv > 0xFFFFu ? 0xFFFF : (int32_t)(dv * mult + (uint32_t)hb);
                            sm1 = 0xFFFFFFFFu;
                        }
                        if (h < 0) {
                            status = FS_HISTORY;                                      // reference: garbage k
                            nc = n = 0;
                        } else if (h < 128 && i + 1 < nc) {                           // :231
                            const int kz = (h == 0 ? 40 : 31 - flo((uint32_t)h)) + ((h + 16) >> 6) - 24;   // :234 (clz(0) == 40)
                            const uint32_t mz = (1u << kz) - 1u;
                            const uint32_t wz = br.peek();
                            uint32_t block;
                            if (wz >= 0xFF800000u) {                                  // :198-202 with 16 raw bits
                                br.seek(br.off + 9u);
                                block = br.peek() >> 16;
                                br.seek(br.off + 16u);
                            } else {
                                const int pz = flo(~wz);
                                const uint32_t ez = (wz >> ((pz - kz) & 31)) & mz;
                                block = (uint32_t)(31 - pz) * (mz & kmask) + max(ez, 1u) - 1u;   // :236
                                br.seek(br.off + (uint32_t)(31 + kz - pz) + (ez >= 2u ? 1u : 0u));
                            }
                            if (block > 0 && (uint32_t)i + 1u + block > (uint32_t)kMaxFrameSamples) {
                                status = FS_RUN_OVERFLOW;                             // reference: IndexOutOfRange
                                nc = n = 0;
                            }
                            zcnt = block;
                            sm1 = block > 0xFFFFu ? 0xFFFFFFFFu : 0u;                 // :233,:246
                            h = 0;                                                    // :248
                        }
                    }
                }
                out[u] = val;
                k = min(flo((uint32_t)((h >> 9) + 3)), kmod);                         // :221-222
                m0 = (1u << k) - 1u;
              }
              // four residuals leave the lane as one 16-byte store into its row
              if (i4 < nc) row[i4 >> 2] = make_int4(out[0], out[1], out[2], out[3]);
              if (kPublish && ((i4 + 4) & 31) == 0) {        // warp-uniform
                  __threadfence();
                  if (i4 < nc) publish(prog, (uint32_t)i4 + 4u);
              }
            }
        }
        if (kPublish) {
            __threadfence();
            if (work && c < ech) publish(prog, kStreamDone);
        }
    }
    cp_async_wait<0>();            // nothing may land in the ring after this block's shared memory is reused
    if (work) {
        // The cursor is monotone: it ended past the frame iff some symbol did.
        if (d.data_bit + br.consumed() > ref.len * 8u) status = FS_OVERRUN;
        if (status != FS_OK) a.desc[f].status = status;
    }
}

}  // namespace alacgpu
