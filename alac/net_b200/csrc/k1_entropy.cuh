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
#ifdef ALACGPU_CHECKED
    const uint8_t *lim;     // one past the last staged byte (incl. tail padding)
    uint32_t *chk;
#endif

    template <bool kCommit = true>
    __device__ __forceinline__ void top_up()
    {
        const uint32_t want = (wpos >> 2) + kAhead;
#pragma unroll
        for (int j = 0; j < kTopUpMax; j++) {
            uint32_t go = filled < want ? 1u : 0u;
#ifdef ALACGPU_CHECKED
            if (go && !ALACGPU_CHECK(chk, base + ((uint64_t)filled << 4) + 16 <= lim, CK_ARENA)) go = 0u;
#endif
            cp_async16_if(ring + ((filled & (kRingChunks - 1)) << 4), base + ((uint64_t)filled << 4), go);
            filled += go;
        }
        if (kCommit) cp_async_commit();
    }
    __device__ __forceinline__ void init(const uint8_t *arena, uint64_t abs_bit, const uint8_t *ring_ptr,
                                         const uint64_t arena_bytes = 0, uint32_t *check = nullptr)
    {
#ifdef ALACGPU_CHECKED
        lim = arena + arena_bytes;
        chk = check;
        if (!ALACGPU_CHECK(chk, (abs_bit >> 3) + 16u * (kAhead + 2) <= arena_bytes, CK_ARENA)) abs_bit = 0;
#endif
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
// many residuals of its stream are in the plane (every 64 steps: fence, then a relaxed store the
// consumer reads with ld.acquire) and 0xFFFFFFFF when the channel is complete.
__device__ __forceinline__ void publish(uint32_t *p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
constexpr uint32_t kStreamDone = 0xFFFFFFFFu;

// ---- one decode step, in PTX ------------------------------------------------------------------------
// The step is written in PTX so that it stays ~85 straight-line instructions (nvcc's version of the
// same C juggled ten live booleans through P2R/R2P and came out at 150).  Operands:
//   %0 cur  %1 nxt  %2 nn  %3 off  %4 wpos           bit cursor (BitCursor)
//   %5 i    %6 nc   %7 h   %8 smm1 (signModifier-1)  output index, samples per channel, history
//   %9 kk (k + 127)  %10 mk ((1<<k)-1)  %11 mm       parameters of the symbol at the cursor
//   %12 R  %13 W                                      kind of that symbol: run length / raw field (0|1)
//   %14 ring  %15 mult  %16 rssh (32-rss)  %17 kcap (kmod+127)  %18 kmask  %19 kk after a run  %20 row
// Predicates inside: pA lane active, pesc nine 1 bits, pP raw field pending, pV a value completed,
// pU a run length completed, pT a run-length symbol is next.
#define ALACGPU_ENTROPY_STEP                                                                              \
    /* the field at the cursor */                                                                        \
    "shf.l.wrap.b32 w, %1, %0, %3;\n\t"                                                                   \
    "shr.u32 t0, w, 23;\n\t"                                                                              \
    "lop3.b32 fx, t0, 0x1FF, 0x4B000000, 0xBE;\n\t"   /* (~w >> 23) | 2^23-as-float */                   \
    "add.rn.f32 fx, fx, 0fCB000000;\n\t"                                                                  \
    "shr.b32 ex, fx, 23;\n\t"                          /* 127 + flo(~w >> 23); 0: nine 1 bits (:198) */   \
    "setp.eq.u32 pesc, ex, 0;\n\t"                                                                        \
    "sub.u32 x, 135, ex;\n\t"                          /* leading 1 bits */                               \
    "sub.u32 s0, %9, ex;\n\t"                                                                             \
    "add.u32 s0, s0, 8;\n\t"                           /* x + k */                                        \
    "add.u32 s1, s0, 1;\n\t"                                                                              \
    "shf.l.wrap.b32 e, w, 0, s1;\n\t"                                                                     \
    "and.b32 e, e, %10;\n\t"                           /* the k bits after the terminator (:205) */       \
    "max.u32 em, e, 1;\n\t"                                                                               \
    "setp.ge.u32 pbig, e, 2;\n\t"                                                                         \
    "mad.lo.u32 rice, x, %11, %8;\n\t"                                                                    \
    "add.u32 rice, rice, em;\n\t"                      /* :206-210 (+ signModifier, :224) */              \
    "selp.u32 rsh, 16, %16, pR;\n\t"                   /* raw field: 16 bits (:236) or rss (:224) */      \
    "shr.u32 rawv, w, rsh;\n\t"                                                                           \
    "add.u32 rawv, rawv, %8;\n\t"                                                                         \
    "add.u32 rawv, rawv, 1;\n\t"                                                                          \
    "selp.u32 dv, rawv, rice, pW;\n\t"                                                                    \
    /* bits consumed: x + k (+1 if e >= 2, :210) | 9 | the raw field | 0 for an idle lane */             \
    "setp.lt.u32 pA, %5, %6;\n\t"                                                                         \
    "not.pred nA, pA;\n\t"                                                                                \
    "sub.u32 alt, 32, rsh;\n\t"                                                                           \
    "selp.u32 alt, alt, 9, pW;\n\t"                                                                       \
    "selp.u32 alt, alt, 0, pA;\n\t"                                                                       \
    "or.pred palt, pesc, pW;\n\t"                                                                         \
    "or.pred palt, palt, nA;\n\t"                                                                         \
    "add.u32 tb, %3, alt;\n\t"                                                                            \
    "add.u32 ta, %3, s0;\n\t"                                                                             \
    "@pbig add.u32 ta, ta, 1;\n\t"                                                                        \
    "selp.u32 t, tb, ta, palt;\n\t"                                                                       \
    /* move the cursor (BitCursor::seek) */                                                              \
    "setp.ge.u32 prf, t, 32;\n\t"                                                                         \
    "and.b32 %3, t, 31;\n\t"                                                                              \
    "selp.u32 sel, 0x0123, 0x7654, prf;\n\t"                                                              \
    "selp.u32 %0, %1, %0, prf;\n\t"                                                                       \
    "prmt.b32 %1, %2, %1, sel;\n\t"                                                                       \
    "and.b32 wa, %4, 63;\n\t"                                                                             \
    "shl.b32 wa, wa, 2;\n\t"                                                                              \
    "add.u32 wa, wa, %14;\n\t"                                                                            \
    "@prf ld.shared.u32 %2, [wa];\n\t"                                                                    \
    "@prf add.u32 %4, %4, 1;\n\t"                                                                         \
    /* what was completed */                                                                             \
    "not.pred nW, pW;\n\t"                                                                                \
    "and.pred pP, pA, pesc;\n\t"                                                                          \
    "and.pred pP, pP, nW;\n\t"                                                                            \
    "not.pred nP, pP;\n\t"                                                                                \
    "and.pred q0, pA, nP;\n\t"                                                                            \
    "and.pred pU, q0, pR;\n\t"                                                                            \
    "not.pred nR, pR;\n\t"                                                                                \
    "and.pred pV, q0, nR;\n\t"                                                                            \
    /* a value: output (:225-226) and history (:229) */                                                  \
    "and.b32 t1, dv, 1;\n\t"                                                                              \
    "neg.s32 t1, t1;\n\t"                                                                                 \
    "shr.u32 t2, dv, 1;\n\t"                                                                              \
    "xor.b32 t2, t2, t1;\n\t"                                                                             \
    "mad.wide.u32 ad, %5, 4, %20;\n\t"                                                                    \
    "@pV st.global.u32 [ad], t2;\n\t"                                                                     \
    "mul.lo.u32 t3, %7, %15;\n\t"                                                                         \
    "shr.s32 t3, t3, 9;\n\t"                                                                              \
    "sub.s32 t3, %7, t3;\n\t"                                                                             \
    "mad.lo.u32 hn, dv, %15, t3;\n\t"                                                                     \
    "setp.gt.u32 pbv, dv, 0xFFFF;\n\t"                                                                    \
    "selp.s32 hn, 0xFFFF, hn, pbv;\n\t"                                                                   \
    /* output index: one value, or a run of dv zeros skipped (:240-245) */                               \
    "add.u32 isum, %5, dv;\n\t"                                                                           \
    "@pV add.u32 %5, %5, 1;\n\t"                                                                          \
    "@pU mov.u32 %5, isum;\n\t"                                                                           \
    "setp.lt.and.u32 pT, hn, 128, pV;\n\t"             /* :231 */                                         \
    "setp.lt.and.u32 pT, %5, %6, pT;\n\t"                                                                 \
    "setp.lt.and.s32 pF, hn, 0, pV;\n\t"               /* negative history: the lane stops here */        \
    "@pF mov.u32 %6, 0;\n\t"                                                                              \
    "@pV mov.s32 %7, hn;\n\t"                                                                             \
    "@pT mov.s32 %7, 0;\n\t"                           /* :248 */                                         \
    "@pV mov.u32 %8, 0xFFFFFFFF;\n\t"                                                                     \
    "selp.u32 t4, 0xFFFFFFFF, 0, pbv;\n\t"                                                                \
    "@pU mov.u32 %8, t4;\n\t"                          /* :233, :246 */                                   \
    /* k of the next symbol: value (:221-222), run length (:234, clz(0) == 40), or after a run */        \
    "shr.s32 t5, hn, 9;\n\t"                                                                              \
    "add.s32 fk, t5, 0x4B000003;\n\t"                                                                     \
    "add.rn.f32 fk, fk, 0fCB000000;\n\t"                                                                  \
    "shr.b32 t5, fk, 23;\n\t"                                                                             \
    "min.u32 kkv, t5, %17;\n\t"                                                                           \
    "bfind.u32 t6, hn;\n\t"                                                                               \
    "add.u32 t7, hn, 16;\n\t"                                                                             \
    "shr.u32 t7, t7, 6;\n\t"                                                                              \
    "sub.u32 t7, t7, t6;\n\t"                                                                             \
    "add.u32 t7, t7, 134;\n\t"                                                                            \
    "setp.eq.u32 pz, hn, 0;\n\t"                                                                          \
    "selp.u32 t7, 143, t7, pz;\n\t"                                                                       \
    "selp.u32 kn, %19, kkv, pU;\n\t"                                                                      \
    "selp.u32 %9, t7, kn, pT;\n\t"                                                                        \
    "shf.l.wrap.b32 t8, 2, 2, %9;\n\t"                 /* 1 << (kk - 127) */                              \
    "sub.u32 %10, t8, 1;\n\t"                                                                             \
    "selp.u32 t9, %18, 0xFFFFFFFF, pT;\n\t"                                                               \
    "and.b32 %11, %10, t9;\n\t"                        /* :236 masks the multiplier */                    \
    /* kind of the next field */                                                                         \
    "and.pred q0, pP, pR;\n\t"                                                                            \
    "and.pred q1, nP, pT;\n\t"                                                                            \
    "or.pred pR, q0, q1;\n\t"                                                                             \
    "mov.pred pW, pP;\n\t"

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
    // (checked build) the lane's two plane rows lie inside the slot's planes
    if (work && !ALACGPU_CHECK(a.check, ((uint64_t)slot * 2u + 2u) * a.ns * 4u <= a.plane_bytes && d.n <= a.ns, CK_PLANE)) work = false;
    const FrameRef ref = a.refs[f];
    const TrackCfg cfg = a.cfgs[ref.track];
    const uint32_t n = work ? (uint32_t)d.n : 0u;             // samples per channel
    const uint32_t rssh = 32u - (uint32_t)d.rss;              // the raw field after nine 1 bits is rss bits (:198-202)
    const uint32_t kmod = (uint32_t)cfg.rice_kmodifier;
    const uint32_t kmask = (1u << kmod) - 1u;                 // AlacFile.cs:483,:643
    const uint32_t kcap = kmod + 127u;
    const uint32_t kk_after_run = min(128u, kcap);            // history 0: k = min(flo(3), kmod)
    const int32_t h0 = cfg.rice_initial_history;              // :216
    const uint32_t kk0 = min(exp_of(0x4B000000u | (uint32_t)((h0 >> 9) + 3)), kcap);   // :221-222

    BitCursor br;
    br.init(a.arena, work ? ref.off * 8ull + d.data_bit : 0ull, ring_smem + threadIdx.x * (uint32_t)kRingBytes,
            a.arena_bytes, a.check);

    // per-lane decode state
    uint32_t chans = work ? ((d.flags & FF_STEREO) ? 2u : 1u) : 0u;    // channels still to finish, the current one included
    int32_t *row = a.planes + (uint64_t)(work ? slot : 0u) * 2u * a.ns;
    uint32_t *prog = a.progress + (uint64_t)(work ? slot : 0u) * 2u;
    uint32_t mult = (uint32_t)((int32_t)d.rice_mod[0] * (cfg.rice_history_mult / 4));   // :483
    uint32_t i = 0;                          // output index of the next value
    uint32_t nc = n;                         // 0 once the lane has stopped on a fault
    int32_t h = h0;
    uint32_t smm1 = 0xFFFFFFFFu;             // signModifier - 1
    uint32_t kk = kk0;                       // k + 127
    uint32_t mk = (1u << (kk - 127u)) - 1u;  // mask of the k-bit field
    uint32_t mm = mk;                        // multiplier of the unary part (:206; & kmask for a run length, :236)
    uint32_t R = 0, W = 0;                   // the field at the cursor is a zero-run length (:234-236) / a raw field
    uint8_t status = FS_OK;

    for (uint32_t period = 0;; ++period) {
        br.top_up();
        cp_async_wait<1>();                  // everything but the group just committed
        const bool active = i < nc;
        if (kPublish && (period & 3u) == 3u) {   // every 64 steps (the fence waits for the lane's stores: ~15 % of the
                                                 // stage when done every 32)
            __threadfence();
            if (active) publish(prog, i);
        }
        // channel hand-over: a lane that ended its channel (or faulted) during the last period
        if (__any_sync(0xffffffffu, chans != 0u && !active)) {
            if (chans != 0u && !active) {
                if (h < 0) status = FS_HISTORY;                                   // reference: garbage k
                else if (i > (uint32_t)kMaxFrameSamples) status = FS_RUN_OVERFLOW;   // reference: IndexOutOfRange
                const bool dead = status != FS_OK;
                if (kPublish) {              // a faulted lane releases the consumers of all its streams
                    __threadfence();
                    publish(prog, kStreamDone);
                    if (dead && chans == 2u) publish(prog + 1, kStreamDone);
                }
                chans = dead ? 0u : chans - 1u;
                nc = 0;
                if (chans != 0u) {           // channel B starts where A ended (:653)
                    row += a.ns;
                    prog += 1;
                    mult = (uint32_t)((int32_t)d.rice_mod[1] * (cfg.rice_history_mult / 4));
                    i = 0;
                    nc = n;
                    h = h0;
                    smm1 = 0xFFFFFFFFu;
                    kk = kk0;
                    mk = (1u << (kk - 127u)) - 1u;
                    mm = mk;
                    R = W = 0;
                }
            }
        }
        if (!__any_sync(0xffffffffu, chans != 0u)) break;

#pragma unroll 1
        for (int u = 0; u < kPeriod; u += 4) {
            asm volatile(
                "{\n\t"
                ".reg .pred pR, pW, pA, nA, pesc, pbig, palt, prf, pP, nP, nW, nR, pV, pU, pT, pF, pbv, pz, q0, q1;\n\t"
                ".reg .b32 w, t0, fx, ex, x, s0, s1, e, em, rice, rsh, rawv, dv, alt, tb, ta, t, sel, wa;\n\t"
                ".reg .b32 t1, t2, t3, hn, isum, t4, t5, fk, kkv, t6, t7, kn, t8, t9;\n\t"
                ".reg .b64 ad;\n\t"
                "setp.ne.u32 pR, %12, 0;\n\t"
                "setp.ne.u32 pW, %13, 0;\n\t"
                ALACGPU_ENTROPY_STEP ALACGPU_ENTROPY_STEP ALACGPU_ENTROPY_STEP ALACGPU_ENTROPY_STEP
                "selp.u32 %12, 1, 0, pR;\n\t"
                "selp.u32 %13, 1, 0, pW;\n\t"
                "}"
                : "+r"(br.cur), "+r"(br.nxt), "+r"(br.nn), "+r"(br.off), "+r"(br.wpos), "+r"(i), "+r"(nc), "+r"(h),
                  "+r"(smm1), "+r"(kk), "+r"(mk), "+r"(mm), "+r"(R), "+r"(W)
                : "r"(br.ring), "r"(mult), "r"(rssh), "r"(kcap), "r"(kmask), "r"(kk_after_run), "l"(row)
                : "memory");
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
