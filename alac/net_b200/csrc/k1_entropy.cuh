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
// the batch size.  A batch like BASELINE configs[1] gives every SM
// sub-partition ONE such warp, and a lone warp on B200 issues in order, one
// instruction per two cycles, four to five cycles behind the instruction it
// depends on, ~25 cycles for a taken branch and more for a re-convergence.
// Hence the shape of the loop:
//
//   * ONE SYMBOL PER STEP, NO BRANCH.  A step decodes the field at the lane's
//     cursor, whatever it is: a value symbol, a zero-run length symbol
//     (:231-249), or the raw field that follows nine 1-bits (:198-202; the
//     nine bits are their own step).  The kind is per-lane state, applied with
//     selects, so lanes of one warp sit at different output indices and the
//     step is the same straight-line code for all of them.  (r1 history: a
//     lock-step-by-output-index loop with the run / escape handling in a
//     divergent block spent 41 % of its time in that block and 14 % on
//     re-convergence, because with 32 frames per warp SOME lane needs it in
//     42 % of the steps.)
//   * A zero run advances the lane's output index; the planes are cleared
//     before the launch (launch_k1 / k12 / k123), so skipped residuals are 0
//     as in the reference's cleared buffer (:238-245).
//   * No FLO on the dependency chain (~30 cycles on the XU pipe): the count
//     of leading 1-bits and floor(log2) of the history come from the exponent
//     of an exactly representable float (one LOP3 + FADD + shift).
//   * bitstream: each lane owns a 256-byte ring in shared memory, topped up
//     with predicated 16-byte cp.async copies every kPeriod steps (no register
//     ever waits on HBM); the cursor keeps two byte-swapped words in registers
//     plus one prefetched word, so the 32-bit window at the cursor is ONE
//     funnel shift and a step moves the cursor by at most one word.
//   * k <= 22 always ((history >> 9) + 3 < 2^23; zero-run k <= 16), so prefix +
//     terminator + k bits fit the 32-bit window.  "Read k bits, un-read one if
//     the value is <= 1" (AlacFile.cs:205-210) is "consume k-1 bits".
//   * CountLeadingZeros' clz(0) == 40 quirk (AlacFile.cs:190) is kept in the
//     zero-run k, the only place a zero argument can reach it.
//
// Error policy (shared with the oracle): the cursor only moves forward, so
// "some symbol ended past the frame's last bit" is decided once from the
// final cursor (OVERRUN outranks a HISTORY / RUN_OVERFLOW fault); a faulted
// lane stops where the fault was found.  The arena carries enough tail
// padding for a lane that runs past its frame.
#pragma once
#include "alacgpu_device.cuh"
#include "alacgpu_kernels.h"

namespace alacgpu {

constexpr int kRingChunks = 16;              // 256 B of bitstream per lane
constexpr int kRingBytes = kRingChunks * 16;
constexpr int kRingWords = kRingBytes / 4;
constexpr int kPeriod = 16;                  // steps between ring top-ups
constexpr int kTopUpMax = 5;                 // a lane enters at most kPeriod words = 4 chunks (+1 straddle) per period
constexpr int kAhead = 9;                    // chunks requested beyond the chunk of the next prefetch word (see BitCursor)
constexpr int kK1Threads = 128;

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gptr)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async16_if(uint32_t smem_addr, const void *gptr, uint32_t pred)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                 ::"r"(smem_addr), "l"(gptr), "r"(pred) : "memory");
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
// 127 + floor(log2(v)) for 0 < v < 2^23, 0 for v == 0: the exponent field of float(v), built without a
// conversion instruction (2^23 + v is exactly representable; subtracting 2^23 normalises it).
__device__ __forceinline__ uint32_t exp_of(uint32_t bits_4b /* 0x4B000000 | v */)
{
    return __float_as_uint(__uint_as_float(bits_4b) - 8388608.0f) >> 23;
}

// Bit cursor over the lane's ring.  `cur:nxt` hold the 64 bits at the cursor's word, `nn` the raw
// word after them; the ring itself is only read at the prefetch index `wpos` (cursor word + 3).  A
// step moves the cursor by at most 32 bits, so during the kPeriod steps after a top-up and the
// kPeriod steps after the next one -- whose copies are only waited for at the top-up after that --
// the prefetch index stays below wpos + 2 * kPeriod, i.e. within chunk(wpos) + 8: the top-up asks
// for everything below chunk(wpos) + kAhead, at most kTopUpMax new chunks per period, and the live
// window (9 chunks) fits the 16-chunk ring.
struct BitCursor {
    const uint8_t *base;    // 16-byte aligned global address of chunk 0
    uint32_t ring;          // shared-space byte address of this lane's ring (256-byte aligned)
    const uint32_t *ringw;  // the same ring as 64 words (plain pointer: the per-step refill is an ordinary
                            // predicated LDS)
    uint32_t wpos;          // absolute index (from chunk 0) of the next word to prefetch into `nn`
    uint32_t cur, nxt;      // byte-swapped words holding bits [32*w, 32*w+64) at the cursor's word w = wpos - 3
    uint32_t nn;            // raw word w+2
    uint32_t off;           // cursor bit within `cur`, 0..31
    uint32_t pos0;          // cursor at init, in bits from chunk 0
    uint32_t filled;        // chunks [0, filled) have been requested

    __device__ __forceinline__ void top_up()
    {
        const uint32_t want = (wpos >> 2) + kAhead;
#pragma unroll
        for (int j = 0; j < kTopUpMax; j++) {
            const uint32_t go = filled < want ? 1u : 0u;
            cp_async16_if(ring + ((filled & (kRingChunks - 1)) << 4), base + ((uint64_t)filled << 4), go);
            filled += go;
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void init(const uint8_t *arena, uint64_t abs_bit, const uint8_t *ring_ptr)
    {
        const uint64_t byte = abs_bit >> 3;
        base = arena + (byte & ~15ull);
        ring = (uint32_t)__cvta_generic_to_shared(ring_ptr);
        ringw = reinterpret_cast<const uint32_t *>(ring_ptr);
        pos0 = (uint32_t)(byte & 15) * 8u + (uint32_t)(abs_bit & 7);
        const uint32_t word0 = pos0 >> 5;
        off = pos0 & 31;
        wpos = word0 + 3;
        for (filled = 0; filled < (wpos >> 2) + kAhead; ++filled)
            cp_async16(ring + ((filled & (kRingChunks - 1)) << 4), base + ((uint64_t)filled << 4));
        cp_async_commit();
        cp_async_wait<0>();
        cur = bswap32(lds32(ring + word0 * 4u));
        nxt = bswap32(lds32(ring + (word0 + 1) * 4u));
        nn = lds32(ring + (word0 + 2) * 4u);
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(nxt, cur, off); }

    // move the cursor to bit t (0..63) of the current word pair.  The word prefetched here was
    // requested at least one top-up before the last one (see above), so the load needs no ordering
    // against the current period's cp.async wait beyond the compiler barrier that wait already is.
    __device__ __forceinline__ void seek(uint32_t t)
    {
        const bool rf = t >= 32u;
        off = t & 31u;
        cur = rf ? nxt : cur;
        // nxt = rf ? bswap(nn) : nxt in one PRMT: selector 0x0123 reverses nn, 0x7654 passes nxt
        nxt = __byte_perm(nn, nxt, rf ? 0x0123u : 0x7654u);
        if (rf) nn = ringw[wpos & (kRingWords - 1)];
        wpos += rf ? 1u : 0u;
    }
    __device__ __forceinline__ uint32_t consumed() const { return (wpos - 3u) * 32u + off - pos0; }
};

// Progress hand-off to the LPC warps of the fused kernel (k12_decode.cu): a lane publishes how
// many residuals of its stream are in the plane (every 32 steps: fence, then a relaxed store the
// consumer reads with ld.acquire) and 0xFFFFFFFF when the channel is complete.
__device__ __forceinline__ void publish(uint32_t *p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
constexpr uint32_t kStreamDone = 0xFFFFFFFFu;

// One block of 128 threads = 4 entropy warps.  `block` is the index among the entropy blocks;
// ring_smem: kRingBytes * kK1Threads bytes, 256-byte aligned.  The planes must be zero on entry.
template <bool kPublish>
__device__ __forceinline__ void entropy_block(const ChunkArgs &a, const int lanes_log2, const uint32_t block,
                                              uint8_t *ring_smem)
{
    const int lane = threadIdx.x & 31;
    const int S = 1 << lanes_log2;
    const uint32_t gw = (block * kK1Threads + threadIdx.x) >> 5;
    const uint32_t slot = gw * (uint32_t)S + (uint32_t)lane;
    // Every lane stays in the loop (the exit is a warp vote); a lane without work has no channels.
    bool work = lane < S && slot < a.n;
    const uint64_t f = a.f0 + (work ? slot : 0u);
    const FrameDesc d = a.desc[f];
    work = work && d.status == FS_OK && !(d.flags & FF_ESCAPE) && d.n > 0;   // escape frames are read directly by K3
    const FrameRef ref = a.refs[f];
    const TrackCfg cfg = a.cfgs[ref.track];
    const uint32_t nc = work ? (uint32_t)d.n : 0u;            // samples per channel
    const uint32_t rssh = 32u - (uint32_t)d.rss;              // the raw field after nine 1 bits is rss bits (:198-202)
    const uint32_t kmod = (uint32_t)cfg.rice_kmodifier;
    const uint32_t kmask = (1u << kmod) - 1u;                 // AlacFile.cs:483,:643
    const uint32_t kcap = kmod + 127u;
    const int32_t h0 = cfg.rice_initial_history;              // :216

    BitCursor br;
    br.init(a.arena, work ? ref.off * 8ull + d.data_bit : 0ull, ring_smem + threadIdx.x * (uint32_t)kRingBytes);

    // per-lane decode state
    uint32_t chans = work ? ((d.flags & FF_STEREO) ? 2u : 1u) : 0u;    // channels still to finish, the current one included
    int32_t *row = a.planes + (uint64_t)(work ? slot : 0u) * 2u * a.ns;
    uint32_t *prog = a.progress + (uint64_t)(work ? slot : 0u) * 2u;
    uint32_t mult = (uint32_t)((int32_t)d.rice_mod[0] * (cfg.rice_history_mult / 4));   // :483
    uint32_t i = 0;                          // output index of the next value
    int32_t h = h0;
    uint32_t smm1 = 0xFFFFFFFFu;             // signModifier - 1
    uint32_t k = min(exp_of(0x4B000000u | (uint32_t)((h0 >> 9) + 3)), kcap) - 127u;   // :221-222
    uint32_t mk = (1u << k) - 1u;            // mask of the k-bit field
    uint32_t mm = mk;                        // multiplier of the unary part (:206; & kmask for a run length, :236)
    bool runmode = false;                    // the symbol at the cursor is a zero-run length (:234-236)
    bool israw = false;                      // the field at the cursor is the raw one after nine 1 bits
    bool active = work;                      // the lane has a symbol to decode in its current channel
    uint8_t status = FS_OK;

    for (uint32_t period = 0;; ++period) {
        br.top_up();
        cp_async_wait<1>();                  // everything but the group just committed
        if (kPublish && (period & 1u)) {     // every 32 steps
            __threadfence();
            if (active) publish(prog, i);
        }
        // channel hand-over: a lane that ended its channel (or faulted) during the last period
        if (__any_sync(0xffffffffu, chans != 0u && !active)) {
            if (chans != 0u && !active) {
                const bool dead = status != FS_OK;
                if (kPublish) {              // a faulted lane releases the consumers of all its streams
                    __threadfence();
                    publish(prog, kStreamDone);
                    if (dead && chans == 2u) publish(prog + 1, kStreamDone);
                }
                chans = dead ? 0u : chans - 1u;
                if (chans != 0u) {           // channel B starts where A ended (:653)
                    row += a.ns;
                    prog += 1;
                    mult = (uint32_t)((int32_t)d.rice_mod[1] * (cfg.rice_history_mult / 4));
                    i = 0;
                    h = h0;
                    smm1 = 0xFFFFFFFFu;
                    k = min(exp_of(0x4B000000u | (uint32_t)((h0 >> 9) + 3)), kcap) - 127u;
                    mk = (1u << k) - 1u;
                    mm = mk;
                    runmode = israw = false;
                    active = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, chans != 0u)) break;

#pragma unroll 4
        for (int u = 0; u < kPeriod; u++) {
            // ---- the field at the cursor -------------------------------------------------------
            const uint32_t w = br.peek();
            const uint32_t ex = exp_of(((w >> 23) ^ 0x1FFu) | 0x4B000000u);   // 127 + flo(~w >> 23); 0: nine 1 bits
            const bool esc = ex == 0u;                                        // :198
            const uint32_t x = 135u - ex;                                     // leading 1 bits (0..8)
            const uint32_t s1 = x + k + 1u;                                   // unary part, terminator, k bits
            const uint32_t e = __funnelshift_l(w, 0u, s1) & mk;               // the k bits (:205)
            const uint32_t em = max(e, 1u);
            const uint32_t rice = x * mm + smm1 + em;                         // :206-210 (+ signModifier, :224)
            const uint32_t rawsh = runmode ? 16u : rssh;                      // :236 reads 16 raw bits, :224 rss
            const uint32_t rawv = (w >> rawsh) + (smm1 + 1u);
            const uint32_t dv = israw ? rawv : rice;
            uint32_t cons = s1 - (e < 2u ? 1u : 0u);                          // x + k, one more if e >= 2 (:210)
            cons = esc ? 9u : cons;
            cons = israw ? 32u - rawsh : cons;
            cons = active ? cons : 0u;
            br.seek(br.off + cons);
            const bool pend = active && esc && !israw;                        // raw field next step
            const bool done = active && !pend;
            const bool isval = done && !runmode;
            const bool isrun = done && runmode;
            // ---- a value: output, history (:225-229) -------------------------------------------
            if (isval) row[i] = (int32_t)(dv >> 1) ^ -(int32_t)(dv & 1u);
            const int32_t hb = h - ((int32_t)((uint32_t)h * mult) >> 9);
            int32_t hn = (int32_t)(dv * mult + (uint32_t)hb);
            hn = dv > 0xFFFFu ? 0xFFFF : hn;
            const uint32_t i1 = i + (isval ? 1u : 0u) + (isrun ? dv : 0u);    // a run of dv zeros is skipped (:240-245)
            const bool hfault = isval && hn < 0;                              // reference: garbage k
            const bool rfault = isrun && dv != 0u && i + dv > (uint32_t)kMaxFrameSamples;   // reference: IndexOutOfRange
            const bool torun = isval && (uint32_t)hn < 128u && i1 < nc;       // :231
            // ---- parameters of the next symbol -------------------------------------------------
            h = isval ? (torun ? 0 : hn) : h;                                 // :248 (h stays 0 through the run symbol)
            smm1 = isval ? 0xFFFFFFFFu : isrun ? (dv > 0xFFFFu ? 0xFFFFFFFFu : 0u) : smm1;   // :233,:246
            const uint32_t kv = min(exp_of((uint32_t)((h >> 9) + 0x4B000003)), kcap) - 127u;   // :221-222
            const uint32_t hz = (uint32_t)hn & 127u;
            const uint32_t kz = hz == 0u ? 16u : 134u - exp_of(hz | 0x4B000000u) + ((hz + 16u) >> 6);   // :234 (clz(0) == 40)
            k = torun ? kz : kv;
            mk = (1u << k) - 1u;
            mm = torun ? (mk & kmask) : mk;
            runmode = pend ? runmode : torun;
            israw = pend;
            if (hfault) status = FS_HISTORY;
            if (rfault) status = FS_RUN_OVERFLOW;
            i = i1;
            active = active && !hfault && !rfault && i1 < nc;
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
