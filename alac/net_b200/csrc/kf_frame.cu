// KF -- the frame-lane decode path for machine-filling chunks.
//
// Replaces, for chunks with (many) more frames than the GPU has lanes, the whole of
// AlacFile.DecodeFrame behind the header parse (ALACDecoder/AlacFile.cs:476-719):
// EntropyRiceDecode (:214-252), PredictorDecompressFirAdapt (:256-336), Deinterlace16/24
// (:338-421), the mono packers (:527-575) and AlacContext.FormatSamples (AlacContext.cs:214-256).
//
// Why a second mapping.  The stream-lane path (k12_decode.cu) splits a frame over an entropy
// lane, LPC lanes and pack threads so that a SMALL batch finishes early; the price is that every
// residual and every predicted sample crosses HBM (planes cleared, written 4 bytes at a time,
// re-read, re-written, re-read: r1 measured 2.58 GB of DRAM traffic for 0.62 GB of algorithmic
// bytes, and 21.8 sectors per store request).  When a chunk holds more frames than the machine has
// lanes there is nothing to gain from splitting a frame: throughput is instructions per sample.
// Here ONE LANE owns one channel of one frame from bitstream to output:
//
//   phase A (kf_frames<false>): lane = frame.  Entropy-decode channel A one symbol per step
//       (the branch-free step of k1_entropy.cuh, with the residual left in a register instead of
//       a plane) and run the predictor on it in the same lane (coefficients and history in
//       registers, templated on the order).  A mono element is un-mixed/packed on the spot and its
//       PCM leaves through a lane-private shared-memory ring in 16-byte stores.  A stereo element
//       cannot be packed yet (channel B starts where A's bits end, AlacFile.cs:643,:653), so its
//       channel-A samples go to a HALF-WIDTH plane -- only the low 16 bits matter for 16-bit PCM
//       (AlacContext.cs:234-238 truncates), 32 bits for 24-bit PCM -- and the lane records where
//       channel B's bits start.
//   phase B (kf_frames<true>): lane = stereo frame.  Entropy + predictor of channel B, channel A
//       streamed back through a cp.async ring, un-mix, wasted-byte merge, PCM.
//   pack-only frames (uncompressed / failed at the header): k3's pack code over a work list.
//
// A warp must be homogeneous in predictor order (the tap code is straight-line, templated on the
// order), so a counting sort groups the chunk's frames by (kind, order of the phase's channel) and
// pads every class to whole warps; heavy orders come first in the grid.  The two channels of a
// frame have independent orders, hence two phases with two work lists rather than one lane doing
// A then B (which would need 32 x 32 classes).
//
// HBM traffic per 16-bit stereo sample-frame: compressed bytes once, 2 + 2 bytes of channel-A
// plane, 4 bytes of PCM -- no clear, no residual plane, no second pass over the samples.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "k1_entropy.cuh"
#include "k3_pack.cuh"
#include "lpc_tap.cuh"

namespace alacgpu {

constexpr int kKfThreads = 64;                 // two warps per block: every structure is lane-private
constexpr int kKfPeriod = 16;                  // iterations between ring top-ups / flushes
constexpr uint32_t kStageStride = 144;         // bytes per lane: 128-byte ring + 16 (lanes 8 apart share banks: 4-way)
constexpr uint32_t kStageRing = 128;
// channel-A ring of phase B: 256 bytes + 16 per lane when the device holds 24-bit tracks (4-byte plane samples),
// 128 + 16 when it holds only 16-bit ones (a sixth block then fits the SM's shared memory)
constexpr uint32_t kARingStride = 272;
__host__ __device__ constexpr uint32_t kf_aring_stride(uint32_t kf_row, uint32_t ns) { return kf_row == ns * 2u ? 144u : kARingStride; }
constexpr uint32_t kKfWarpSmemA = kRingBytes * 32 + kStageStride * 32;                 // 12800

constexpr uint32_t kNoSlot = 0xFFFFFFFFu;

// ---- classes ---------------------------------------------------------------------------------
// kind: bit 0 container has two channels, bit 1 24-bit samples, bit 2 stereo element
__device__ __forceinline__ uint32_t kf_kind(const FrameDesc &d, const TrackCfg &cfg)
{
    return (cfg.num_channels == 2 ? 1u : 0u) | (cfg.sample_size == 24 ? 2u : 0u) | ((d.flags & FF_STEREO) ? 4u : 0u);
}
struct KfClass { int a, b; bool e; };          // class in the phase-A / phase-B list (-1: not on it); pack-only
__device__ __forceinline__ KfClass kf_classify(const FrameDesc &d, const TrackCfg &cfg)
{
    KfClass k{-1, -1, false};
    if (d.out_len == 0) return k;                                  // nothing to emit
    if (!(d.status == FS_OK && !(d.flags & FF_ESCAPE) && d.n > 0)) { k.e = true; return k; }
    const uint32_t kind = kf_kind(d, cfg);
    k.a = (int)(kind * 32u + d.order[0]);
    if (d.flags & FF_STEREO) k.b = (int)(kind * 32u + d.order[1]);
    return k;
}
// (grid order of the classes, kf_scan: heaviest predictor first -- 30 .. 1 -- then delta mode, then order 0)

// kf_count: [0] phase-A entries (padded), [1] phase-B entries (padded), [2] pack-only frames, then
// hist[2][256] at 8, cursor[2][256] behind it, and the per-SM schedule of the two phases (see kf_frames):
// seg_bound[2][kMaxSeg + 1] (warp index where segment s of the class-sorted list starts), seg_next[2][kMaxSeg]
// (warps handed out per segment), sm_seg[2][kMaxSm] (segment claimed by SM id: 0 none, 1 being claimed,
// s + 2), seg_claim[2] (segments claimed so far).
constexpr uint32_t kMaxSeg = 256, kMaxSm = 512;
constexpr uint32_t kHist = 8, kCursor = kHist + 2 * kKfClasses;
constexpr uint32_t kSegBound = kCursor + 2 * kKfClasses, kSegNext = kSegBound + 2 * (kMaxSeg + 1);
constexpr uint32_t kSmSeg = kSegNext + 2 * kMaxSeg, kSegClaim = kSmSeg + 2 * kMaxSm;
static_assert(kSegClaim + 2 <= kKfCountWords, "kf_count scratch too small");

__global__ void __launch_bounds__(256)
kf_hist(const FrameDesc *__restrict__ desc, const FrameRef *__restrict__ refs, const TrackCfg *__restrict__ cfgs,
        uint32_t n, uint32_t *__restrict__ cnt)
{
    __shared__ uint32_t h[2 * kKfClasses];
    for (uint32_t k = threadIdx.x; k < 2 * kKfClasses; k += 256) h[k] = 0;
    __syncthreads();
    const uint32_t s = blockIdx.x * 256 + threadIdx.x;
    if (s < n) {
        const FrameDesc d = desc[s];
        const KfClass k = kf_classify(d, cfgs[refs[s].track]);
        if (k.a >= 0) atomicAdd(&h[k.a], 1u);
        if (k.b >= 0) atomicAdd(&h[kKfClasses + k.b], 1u);
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < 2 * kKfClasses; k += 256)
        if (h[k]) atomicAdd(&cnt[kHist + k], h[k]);
}

// class r-th in grid order; instructions one warp issues per sample of that class (entropy step + taps + output)
__device__ __forceinline__ uint32_t kf_class_at(uint32_t r)
{
    const uint32_t rr = r / 8u, kind = r % 8u;
    const uint32_t order = rr == 31 ? 0u : (rr == 30 ? 31u : 30u - rr);
    return kind * 32u + order;
}
__device__ __forceinline__ uint32_t kf_class_weight(uint32_t cls)
{
    const uint32_t order = cls & 31u;
    return 209u + (order == 31u ? 10u : 10u * order);      // measured (r2, 128 tracks of one order each): 21.7 ms + 1.04 ms per tap
}

__global__ void __launch_bounds__(256)
kf_scan(uint32_t *__restrict__ cnt, const uint32_t list_cap, const uint32_t nseg)
{
    // one thread per list: class -> start of its (padded) range, in grid order; then the list is cut into
    // `nseg` contiguous segments of equal estimated work, one per SM (kf_frames)
    if (threadIdx.x < 2) {
        const uint32_t l = threadIdx.x;
        uint32_t acc = 0;
        unsigned long long work = 0;
        for (uint32_t r = 0; r < kKfClasses; r++) {
            const uint32_t cls = kf_class_at(r);
            cnt[kCursor + l * kKfClasses + cls] = l * list_cap + acc;
            const uint32_t padded = (cnt[kHist + l * kKfClasses + cls] + 31u) & ~31u;
            acc += padded;
            work += (unsigned long long)(padded / 32u) * kf_class_weight(cls);
        }
        cnt[l] = acc;
        uint32_t *bound = cnt + kSegBound + l * (kMaxSeg + 1);
        uint32_t s = 1, warp0 = 0;
        unsigned long long cum = 0;
        bound[0] = 0;
        for (uint32_t r = 0; r < kKfClasses && s < nseg; r++) {
            const uint32_t cls = kf_class_at(r);
            const uint32_t warps = ((cnt[kHist + l * kKfClasses + cls] + 31u) & ~31u) / 32u;
            const unsigned long long w = kf_class_weight(cls);
            while (s < nseg && work * s <= (cum + warps * w) * nseg) {      // target(s) = work * s / nseg falls inside this class
                const unsigned long long target = work * s / nseg;
                bound[s++] = warp0 + (uint32_t)((target > cum ? target - cum : 0ull) / w);
            }
            cum += warps * w;
            warp0 += warps;
        }
        for (; s <= nseg; s++) bound[s] = acc / 32u;
    }
}

__global__ void __launch_bounds__(256)
kf_scatter(const FrameDesc *__restrict__ desc, const FrameRef *__restrict__ refs, const TrackCfg *__restrict__ cfgs,
           uint32_t n, uint32_t *__restrict__ cnt, uint32_t *__restrict__ list, const uint32_t list_cap)
{
    // block-local ranks in shared memory, one global reservation per class and block
    __shared__ uint32_t h[2 * kKfClasses], base[2 * kKfClasses], ecount, ebase;
    for (uint32_t k = threadIdx.x; k < 2 * kKfClasses; k += 256) h[k] = 0;
    if (threadIdx.x == 0) ecount = 0;
    __syncthreads();
    const uint32_t s = blockIdx.x * 256 + threadIdx.x;
    KfClass k{-1, -1, false};
    uint32_t ra = 0, rb = 0, re = 0;
    if (s < n) {
        const FrameDesc d = desc[s];
        k = kf_classify(d, cfgs[refs[s].track]);
        if (k.a >= 0) ra = atomicAdd(&h[k.a], 1u);
        if (k.b >= 0) rb = atomicAdd(&h[kKfClasses + k.b], 1u);
        if (k.e) re = atomicAdd(&ecount, 1u);
    }
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < 2 * kKfClasses; c += 256)
        if (h[c]) base[c] = atomicAdd(&cnt[kCursor + c], h[c]);
    if (threadIdx.x == 0 && ecount) ebase = atomicAdd(&cnt[2], ecount);
    __syncthreads();
    if (k.a >= 0) list[base[k.a] + ra] = s;
    if (k.b >= 0) list[base[kKfClasses + k.b] + rb] = s;
    if (k.e) list[2u * list_cap + ebase + re] = s;
}

// ---- lane-private output stage -----------------------------------------------------------------
// A lane writes its bytes (PCM, or channel-A plane samples) into its own 128-byte shared-memory
// ring at the position the byte will have in memory modulo 128, and every period moves the whole
// 16-byte groups out with one LDS.128 + STG.128 each.  No other thread touches the ring, so no
// barrier is involved.  `head`: a frame whose PCM does not start on a 16-byte boundary (it follows
// a frame with an odd sample count) shares its first group with its predecessor: byte stores there.
// g[b] = ring byte b for b in [lo, hi): the partial groups at the two ends of a frame
__device__ __noinline__ void stage_bytes(uint8_t *g, const uint32_t s, const uint32_t lo, const uint32_t hi)
{
#pragma unroll 1
    for (uint32_t b = lo; b < hi; b++) {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(s + (b & 127u)) : "memory");
        g[b] = (uint8_t)v;
    }
}

struct OutStage {
    uint8_t *g;          // global address of position 0 (16-byte aligned)
    uint32_t s;          // shared-space address of the ring
    uint32_t p, F, head; // next byte position; flushed up to (multiple of 16); first own byte of group 0 (0 once it is out)
#ifdef ALACGPU_CHECKED
    uint64_t room;       // bytes from position 0 to the end of the destination buffer
    uint32_t *chk;
    int what;
#endif

    __device__ __forceinline__ void init(uint8_t *dst, uint32_t saddr, const uint8_t *buf = nullptr, uint64_t buf_bytes = 0,
                                         uint32_t *check = nullptr, int code = 0)
    {
#ifdef ALACGPU_CHECKED
        chk = check;
        what = code;
        const bool inside = ALACGPU_CHECK(chk, dst >= buf && dst <= buf + buf_bytes, code);
        if (!inside) dst = const_cast<uint8_t *>(buf);
        room = (uint64_t)(buf + buf_bytes - dst) + ((uintptr_t)dst & 15u);
#endif
        head = (uint32_t)((uintptr_t)dst & 15u);
        g = dst - head;
        s = saddr;
        asm volatile("" : "+r"(s));     // keep the ring address in a register (ptxas otherwise re-derives it from %tid every sample)
        p = head;
        F = 0;
    }
    // predicated stores: a lane without a sample this round stores nothing and does not advance
    __device__ __forceinline__ void put8(uint32_t v, uint32_t on)
    {
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.u8 [%0], %1;\n\t}"
                     ::"r"(s + (p & (kStageRing - 1))), "r"(v), "r"(on) : "memory");
        p += on;
    }
    __device__ __forceinline__ void put16(uint32_t v, uint32_t on)
    {
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.u16 [%0], %1;\n\t}"
                     ::"r"(s + (p & (kStageRing - 1))), "h"((uint16_t)v), "r"(on) : "memory");
        p += on << 1;
    }
    __device__ __forceinline__ void put32(uint32_t v, uint32_t on)
    {
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.u32 [%0], %1;\n\t}"
                     ::"r"(s + (p & (kStageRing - 1))), "r"(v), "r"(on) : "memory");
        p += on << 2;
    }
    // whole groups below p; at most kGroups per call (a period adds at most 16 x 6 bytes)
    template <int kGroups>
    __device__ __forceinline__ void flush()
    {
        if (head != 0u && p >= 16u) {           // rare: group 0 is shared with the previous frame's last bytes
            stage_bytes(g, s, head, 16u);
            head = 0u;
            F = 16u;
        }
        const uint32_t go = head == 0u ? 1u : 0u;
#ifdef ALACGPU_CHECKED
        ALACGPU_CHECK(chk, p - F <= kStageRing + 15u, CK_RING);      // nothing was overwritten before it left
#endif
#pragma unroll
        for (int k = 0; k < kGroups; k++) {
            uint32_t on = (go != 0u && F + 16u <= p) ? 1u : 0u;
#ifdef ALACGPU_CHECKED
            if (on && !ALACGPU_CHECK(chk, (uint64_t)F + 16u <= room, what)) on = 0u;
#endif
            uint4 v;
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %6, 0;\n\t"
                         "@q ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n\t"
                         "@q st.global.v4.u32 [%5], {%0,%1,%2,%3};\n\t}"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(s + (F & (kStageRing - 1))), "l"(g + F), "r"(on) : "memory");
            F += on << 4;
        }
    }
    __device__ __forceinline__ void finish()       // the last, partial group (and group 0 of a very short frame)
    {
        const uint32_t lo = max(F, head);
#ifdef ALACGPU_CHECKED
        if (lo < p && !ALACGPU_CHECK(chk, (uint64_t)p <= room, what)) { F = p; return; }
#endif
        if (lo < p) stage_bytes(g, s, lo, p);
        F = p;
    }
};

// Channel-A samples of a stereo frame coming back in phase B: a lane-private cp.async ring over the
// lane's plane row (16 chunks of 16 bytes), topped up every period like the bitstream ring.
struct PlaneRing {
    const uint8_t *base; // plane row (16-byte aligned)
    uint32_t s;          // shared-space address of the ring
    uint32_t filled;     // chunks requested so far
    uint32_t limit;      // chunks that hold samples of this frame
    uint32_t mask;       // ring bytes - 1 (255 or 127)
    uint32_t ahead;      // chunks requested beyond the one being read: two periods of samples + the straddle
    __device__ __forceinline__ void init(const uint8_t *row, uint32_t saddr, uint32_t bytes, uint32_t ring_bytes)
    {
        base = row; s = saddr; filled = 0; limit = (bytes + 15u) >> 4;
        mask = ring_bytes - 1u;
        ahead = ring_bytes == 256u ? 10u : 6u;       // 256: 32 four-byte samples = 8 chunks (+2); 128: 32 two-byte samples = 4 (+2)
        asm volatile("" : "+r"(s));
    }
    // `byte`: offset of the next sample to be read; a period reads at most 16 samples (64 bytes) and the
    // copies issued here are only waited for at the NEXT top-up: ask for two periods + the straddle
    __device__ __forceinline__ void top_up(uint32_t byte)
    {
        const uint32_t want = min((byte >> 4) + ahead, limit);
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const uint32_t go = filled < want ? 1u : 0u;
            cp_async16_if(s + ((filled << 4) & mask), base + ((uint64_t)filled << 4), go);
            filled += go;
        }
    }
    __device__ __forceinline__ void prime()        // first two periods' worth, waited for by the caller
    {
        for (; filled < min(ahead, limit); ++filled) cp_async16(s + ((filled << 4) & mask), base + ((uint64_t)filled << 4));
    }
    __device__ __forceinline__ uint32_t get16(uint32_t byte) const
    {
        uint32_t v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(s + (byte & mask)) : "memory");
        return v;
    }
    __device__ __forceinline__ uint32_t get32(uint32_t byte) const { return lds32(s + (byte & mask)); }
};

// ---- one entropy step with the residual left in a register ------------------------------------------
// k1_entropy.cuh's ALACGPU_ENTROPY_STEP with three changes: a lane that still holds unconsumed residuals
// (%14 cnt: 1 after a value, r after a run of r zeros) sits the step out; a completed value goes to %15 instead of a plane; a completed
// zero-run length becomes %14 = the number of zero residuals still to hand out (clipped to the frame)
// while the symbol index jumps as before.  Operands:
//   %0 cur %1 nxt %2 nn %3 off %4 wpos | %5 i %6 nc %7 h %8 smm1 %9 kk %10 mk %11 mm %12 R %13 W
//   %14 cnt %15 e | %16 ring %17 mult %18 rssh %19 kcap %20 kmask %21 kk after a run
#define ALACGPU_ENTROPY_STEP_F                                                                            \
    "setp.ne.u32 pR, %12, 0;\n\t"                                                                         \
    "setp.ne.u32 pW, %13, 0;\n\t"                                                                         \
    "shf.l.wrap.b32 w, %1, %0, %3;\n\t"                                                                   \
    "shr.u32 t0, w, 23;\n\t"                                                                              \
    "lop3.b32 fx, t0, 0x1FF, 0x4B000000, 0xBE;\n\t"                                                       \
    "add.rn.f32 fx, fx, 0fCB000000;\n\t"                                                                  \
    "shr.b32 ex, fx, 23;\n\t"                                                                             \
    "setp.eq.u32 pesc, ex, 0;\n\t"                                                                        \
    "sub.u32 x, 135, ex;\n\t"                                                                             \
    "sub.u32 s0, %9, ex;\n\t"                                                                             \
    "add.u32 s0, s0, 8;\n\t"                                                                              \
    "add.u32 s1, s0, 1;\n\t"                                                                              \
    "shf.l.wrap.b32 ee, w, 0, s1;\n\t"                                                                    \
    "and.b32 ee, ee, %10;\n\t"                                                                            \
    "max.u32 em, ee, 1;\n\t"                                                                              \
    "setp.ge.u32 pbig, ee, 2;\n\t"                                                                        \
    "mad.lo.u32 rice, x, %11, %8;\n\t"                                                                    \
    "add.u32 rice, rice, em;\n\t"                                                                         \
    "selp.u32 rsh, 16, %18, pR;\n\t"                                                                      \
    "shr.u32 rawv, w, rsh;\n\t"                                                                           \
    "add.u32 rawv, rawv, %8;\n\t"                                                                         \
    "add.u32 rawv, rawv, 1;\n\t"                                                                          \
    "selp.u32 dv, rawv, rice, pW;\n\t"                                                                    \
    "setp.lt.u32 pA, %5, %6;\n\t"                                                                         \
    "setp.eq.and.u32 pA, %14, 0, pA;\n\t"              /* a lane with residuals in hand (a value, or zeros of a run) waits */ \
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
    "setp.ge.u32 prf, t, 32;\n\t"                                                                         \
    "and.b32 %3, t, 31;\n\t"                                                                              \
    "selp.u32 sel, 0x0123, 0x7654, prf;\n\t"                                                              \
    "selp.u32 %0, %1, %0, prf;\n\t"                                                                       \
    "prmt.b32 %1, %2, %1, sel;\n\t"                                                                       \
    "and.b32 wa, %4, 63;\n\t"                                                                             \
    "shl.b32 wa, wa, 2;\n\t"                                                                              \
    "add.u32 wa, wa, %16;\n\t"                                                                            \
    "@prf ld.shared.u32 %2, [wa];\n\t"                                                                    \
    "@prf add.u32 %4, %4, 1;\n\t"                                                                         \
    "not.pred nW, pW;\n\t"                                                                                \
    "and.pred pP, pA, pesc;\n\t"                                                                          \
    "and.pred pP, pP, nW;\n\t"                                                                            \
    "not.pred nP, pP;\n\t"                                                                                \
    "and.pred q0, pA, nP;\n\t"                                                                            \
    "and.pred pU, q0, pR;\n\t"                                                                            \
    "not.pred nR, pR;\n\t"                                                                                \
    "and.pred pV, q0, nR;\n\t"                                                                            \
    "and.b32 t1, dv, 1;\n\t"                                                                              \
    "neg.s32 t1, t1;\n\t"                                                                                 \
    "shr.u32 t2, dv, 1;\n\t"                                                                              \
    "xor.b32 t2, t2, t1;\n\t"                                                                             \
    "@pV mov.b32 %15, t2;\n\t"                         /* the residual (:225-226) */                      \
    "@pV mov.u32 %14, 1;\n\t"                                                                             \
    "mul.lo.u32 t3, %7, %17;\n\t"                                                                         \
    "shr.s32 t3, t3, 9;\n\t"                                                                              \
    "sub.s32 t3, %7, t3;\n\t"                                                                             \
    "mad.lo.u32 hn, dv, %17, t3;\n\t"                                                                     \
    "setp.gt.u32 pbv, dv, 0xFFFF;\n\t"                                                                    \
    "selp.s32 hn, 0xFFFF, hn, pbv;\n\t"                                                                   \
    "add.u32 isum, %5, dv;\n\t"                                                                           \
    "min.u32 tz, isum, %6;\n\t"                        /* zeros of the run that lie inside the frame */   \
    "sub.u32 tz, tz, %5;\n\t"                                                                             \
    "@pU mov.u32 %14, tz;\n\t"                                                                            \
    "@pV add.u32 %5, %5, 1;\n\t"                                                                          \
    "@pU mov.u32 %5, isum;\n\t"                                                                           \
    "setp.lt.and.u32 pT, hn, 128, pV;\n\t"                                                                \
    "setp.lt.and.u32 pT, %5, %6, pT;\n\t"                                                                 \
    "setp.lt.and.s32 pF, hn, 0, pV;\n\t"                                                                  \
    "@pF mov.u32 %6, 0;\n\t"                                                                              \
    "@pV mov.s32 %7, hn;\n\t"                                                                             \
    "@pT mov.s32 %7, 0;\n\t"                                                                              \
    "@pV mov.u32 %8, 0xFFFFFFFF;\n\t"                                                                     \
    "selp.u32 t4, 0xFFFFFFFF, 0, pbv;\n\t"                                                                \
    "@pU mov.u32 %8, t4;\n\t"                                                                             \
    "shr.s32 t5, hn, 9;\n\t"                                                                              \
    "add.s32 fk, t5, 0x4B000003;\n\t"                                                                     \
    "add.rn.f32 fk, fk, 0fCB000000;\n\t"                                                                  \
    "shr.b32 t5, fk, 23;\n\t"                                                                             \
    "min.u32 kkv, t5, %19;\n\t"                                                                           \
    "bfind.u32 t6, hn;\n\t"                                                                               \
    "add.u32 t7, hn, 16;\n\t"                                                                             \
    "shr.u32 t7, t7, 6;\n\t"                                                                              \
    "sub.u32 t7, t7, t6;\n\t"                                                                             \
    "add.u32 t7, t7, 134;\n\t"                                                                            \
    "setp.eq.u32 pz, hn, 0;\n\t"                                                                          \
    "selp.u32 t7, 143, t7, pz;\n\t"                                                                       \
    "selp.u32 kn, %21, kkv, pU;\n\t"                                                                       \
    "selp.u32 kn, t7, kn, pT;\n\t"                                                                         \
    "and.pred pC, pA, nP;\n\t"                         /* a symbol was completed: only then k moves on */  \
    "@pC mov.u32 %9, kn;\n\t"                                                                             \
    "shf.l.wrap.b32 t8, 2, 2, %9;\n\t"                                                                    \
    "sub.u32 %10, t8, 1;\n\t"                                                                             \
    "selp.u32 t9, %20, 0xFFFFFFFF, pT;\n\t"                                                               \
    "and.b32 t9, %10, t9;\n\t"                                                                            \
    "@pC mov.u32 %11, t9;\n\t"                                                                            \
    "and.pred q0, pP, pR;\n\t"                                                                            \
    "and.pred q1, nP, pT;\n\t"                                                                            \
    "or.pred q0, q0, q1;\n\t"                                                                             \
    "@pA selp.u32 %12, 1, 0, q0;\n\t"                  /* a waiting lane keeps the kind of its next field */ \
    "@pA selp.u32 %13, 1, 0, pP;\n\t"

// ---- the predictor, one sample in one lane --------------------------------------------------------
// M = 1..30: adaptive FIR (AlacFile.cs:284-334); M = 31: delta mode (:268-282); M = 0: identity (:261-267).
template <int M>
struct LaneLpc {
    static constexpr int kTaps = (M >= 1 && M <= 30) ? M : 1;
    int32_t c[kTaps], H[kTaps + 1];      // H[j] = o[i-1-j]; H[kTaps] is the base o[i-1-M]
    int32_t rnd;
    uint32_t rneg, q;
    int sh;

    __device__ __forceinline__ void init(const int16_t *__restrict__ coef16, const int quant, const int rss, const bool active)
    {
#pragma unroll
        for (int j = 0; j < kTaps; j++) c[j] = (active && M >= 1 && M <= 30) ? (int32_t)coef16[j] : 0;
#pragma unroll
        for (int j = 0; j <= kTaps; j++) H[j] = 0;
        q = (uint32_t)quant;
        rnd = (int32_t)(1u << ((quant - 1) & 31));          // :306 (quant 0 -> 1 << 31)
        rneg = (1u << quant) - 1u;                          // see k2_lpc.cuh lpc_warp
        sh = (32 - rss) & 31;
    }
    // sample i with residual e.  `hv`: the lane really has a sample this round; a lane without one has finished
    // its channel (the predictor only runs when every unfinished lane holds a residual), so its state may move
    // freely -- only its coefficient update is switched off, to keep the arithmetic in range
    __device__ __forceinline__ int32_t step(const int32_t e, const uint32_t i, const bool hv)
    {
        if constexpr (M == 0) {
            return e;
        } else if constexpr (M == 31) {
            const int32_t x = (int32_t)((uint32_t)H[0] + (uint32_t)e);
            int32_t o = (int32_t)((uint32_t)x << sh) >> sh;
            o = i == 0 ? e : o;                                         // :259-260
            H[0] = o;
            return o;
        } else {
            const bool main = i > (uint32_t)M;                          // warm-up covers i = 1..M (:284-293)
            const int32_t base = H[M];
            // sign(err) as arithmetic on the FMA pipe (the ALU pipe is the one this kernel saturates): with
            // m = err >> 31 (0 or -1), sign = 2 m + 1, -sign = -2 m - 1
            const int32_t m = e >> 31;
            int32_t sg, nsg;
            asm("mad.lo.s32 %0, %1, 2, 1;" : "=r"(sg) : "r"(m));
            asm("mad.lo.s32 %0, %1, -2, -1;" : "=r"(nsg) : "r"(m));      // -sign(err)
            const int32_t sgbase = (int32_t)((uint32_t)sg * (uint32_t)base);
            const int32_t mag = (int32_t)((uint32_t)sg * (uint32_t)e);   // sign(err) * err
            int32_t E = (main && hv) ? mag : 0;
            const uint32_t r = (uint32_t)m & rneg;
            uint32_t acc = 0;
            lpc_taps<M, M - 1>(c, H, E, acc, nsg, sgbase, r, q);
            const int32_t sum = (int32_t)(acc * (uint32_t)nsg);
            int32_t v = (int32_t)((uint32_t)rnd + (uint32_t)sum) >> q;  // :306-307
            v = (int32_t)((uint32_t)v + (uint32_t)base + (uint32_t)e);  // :308
            const int32_t w = (int32_t)((uint32_t)H[0] + (uint32_t)e);  // warm-up (:288)
            const int32_t x = main ? v : w;
            int32_t o = (int32_t)((uint32_t)x << sh) >> sh;             // :309-310
            o = i == 0 ? e : o;                                         // first sample copies (:259-260)
#pragma unroll
            for (int j = M; j > 0; --j) H[j] = H[j - 1];
            H[0] = o;
            return o;
        }
    }
};

// One warp = 32 frames of one class.  kB = false: channel A (mono elements packed here, stereo elements to
// the plane); kB = true: channel B of stereo elements + un-mix + pack.
template <int M, bool kB>
__device__ __noinline__ void kf_run(const ChunkArgs &a, const uint32_t slot_in, const bool valid, uint8_t *wsm)
{
    const int lane = threadIdx.x & 31;
    const bool in_chunk = ALACGPU_CHECK(a.check, !valid || slot_in < a.n, CK_FRAME);
    const uint32_t slot = (valid && in_chunk) ? slot_in : 0u;
    const uint64_t f = a.f0 + slot;
    const FrameDesc d = a.desc[f];
    const FrameRef ref = a.refs[f];
    const TrackCfg cfg = a.cfgs[ref.track];
    constexpr int ch = kB ? 1 : 0;
    bool work = valid && in_chunk && d.status == FS_OK;             // phase B: channel A may have failed in phase A
    // (checked build) the frame's channel-A row lies inside the slot's plane
    if (work && !ALACGPU_CHECK(a.check, ((uint64_t)slot + 1u) * a.kf_row <= a.plane_bytes && d.n <= a.ns, CK_PLANE)) work = false;
    const uint32_t n = work ? (uint32_t)d.n : 0u;
    const bool is24 = cfg.sample_size == 24;
    const bool stereo = (d.flags & FF_STEREO) != 0;
    // warp-uniform copies for the branches around the output code (a warp never mixes kinds; lane 0 is
    // never a padding lane)
    const uint32_t kind_w = __shfl_sync(0xffffffffu, kf_kind(d, cfg), 0);
    const bool is24_w = (kind_w & 2u) != 0, two_ch_w = (kind_w & 1u) != 0, stereo_w = (kind_w & 4u) != 0;

    // entropy state (k1_entropy.cuh entropy_block)
    const uint32_t rssh = 32u - (uint32_t)d.rss;
    const uint32_t kmod = (uint32_t)cfg.rice_kmodifier;
    const uint32_t kmask = (1u << kmod) - 1u;
    const uint32_t kcap = kmod + 127u;
    const uint32_t kk_after_run = min(128u, kcap);
    const int32_t h0 = cfg.rice_initial_history;
    const uint32_t kk0 = min(exp_of(0x4B000000u | (uint32_t)((h0 >> 9) + 3)), kcap);
    const uint32_t start_bit = kB ? a.bstart[slot] : d.data_bit;
    BitCursor br;
    br.init(a.arena, work ? ref.off * 8ull + start_bit : 0ull, wsm + (uint32_t)lane * (uint32_t)kRingBytes, a.arena_bytes, a.check);
    const uint32_t mult = (uint32_t)((int32_t)d.rice_mod[ch] * (cfg.rice_history_mult / 4));
    uint32_t i = 0, nc = n;
    int32_t h = h0;
    uint32_t smm1 = 0xFFFFFFFFu, kk = kk0, mk = (1u << (kk0 - 127u)) - 1u, mm = mk, R = 0, W = 0;
    uint32_t cnt = 0;                                                // residuals in hand: `e`, then zeros
    int32_t e = 0;

    LaneLpc<M> lpc;
    lpc.init(a.coefs[f].c[ch], d.quant[ch], d.rss, work);

    // output: PCM, or (phase A of a stereo element) the channel-A plane row, 2 bytes per sample for
    // 16-bit tracks and 4 for 24-bit ones
    uint8_t *const plane_row = reinterpret_cast<uint8_t *>(a.planes) + (uint64_t)slot * a.kf_row;
    OutStage out;
    const uint32_t stage_addr = (uint32_t)__cvta_generic_to_shared(wsm + kRingBytes * 32u) + (uint32_t)lane * kStageStride;
    const bool to_plane = !kB && stereo_w;
    out.init(to_plane ? plane_row : a.pcm + (a.frame_off[f] - a.pcm_base), stage_addr,
             to_plane ? reinterpret_cast<const uint8_t *>(a.planes) : a.pcm, to_plane ? a.plane_bytes : a.pcm_bytes, a.check,
             to_plane ? CK_PLANE : CK_PCM);
    PlaneRing ar;
    if (kB) {
        const uint32_t astride = kf_aring_stride(a.kf_row, a.ns);
        ar.init(plane_row, (uint32_t)__cvta_generic_to_shared(wsm + kKfWarpSmemA) + (uint32_t)lane * astride,
                n * (is24 ? 4u : 2u), astride - 16u);
        if (work) ar.prime();
        cp_async_commit();
        cp_async_wait<0>();
    }
    const uint32_t aw = is24 ? 4u : 2u;                             // plane bytes per sample
    // un-mix and wasted bytes (AlacFile.cs:338-421)
    const int mw = d.mix_weight, ms = d.mix_shift & 31;
    const uint32_t ub8 = (is24 && d.ub) ? (uint32_t)d.ub * 8u : 0u;
    const uint32_t shift_mask = ub8 ? ~(0xFFFFFFFFu << ub8) : 0u;
    const uint32_t *arena32 = reinterpret_cast<const uint32_t *>(a.arena);
    const uint64_t shift_pos0 = ref.off * 8ull + d.shift_bit;
    const uint32_t shift_step = (stereo ? 2u : 1u) * ub8;
    const uint32_t ub8n = ub8 ? ub8 : 8u;                            // field width for lanes without wasted bytes (masked to nothing)
    const bool ub_w = is24_w && !to_plane && __any_sync(0xffffffffu, work && ub8 != 0u);

    uint32_t j = 0;                                                 // samples reconstructed
    // every sample takes at most four steps (value and run length, each with its raw field), so the loop ends
    // by itself; the bound only turns a logic fault into a frame status instead of a hung GPU
    const uint32_t max_periods = (4u * (uint32_t)__reduce_max_sync(0xffffffffu, n)) / kKfPeriod + 8u;
    bool stuck = false;
    for (uint32_t period = 0;; ++period) {
        br.top_up<false>();
        if (kB) ar.top_up(j * aw);
        cp_async_commit();
        cp_async_wait<1>();
        out.flush<7>();
        if (!__any_sync(0xffffffffu, cnt != 0u || i < nc)) break;
        if (period >= max_periods) { stuck = cnt != 0u || i < nc; break; }
#pragma unroll 1
        for (int u = 0; u < kKfPeriod; ++u) {
            asm volatile(
                "{\n\t"
                ".reg .pred pR, pW, pA, nA, pesc, pbig, palt, prf, pP, nP, nW, nR, pV, pU, pT, pF, pbv, pz, q0, q1, pC;\n\t"
                ".reg .b32 w, t0, fx, ex, x, s0, s1, ee, em, rice, rsh, rawv, dv, alt, tb, ta, t, sel, wa;\n\t"
                ".reg .b32 t1, t2, t3, hn, isum, tz, t4, t5, fk, kkv, t6, t7, kn, t8, t9;\n\t"
                ALACGPU_ENTROPY_STEP_F
                "}"
                : "+r"(br.cur), "+r"(br.nxt), "+r"(br.nn), "+r"(br.off), "+r"(br.wpos), "+r"(i), "+r"(nc), "+r"(h),
                  "+r"(smm1), "+r"(kk), "+r"(mk), "+r"(mm), "+r"(R), "+r"(W), "+r"(cnt), "+r"(e)
                : "r"(br.ring), "r"(mult), "r"(rssh), "r"(kcap), "r"(kmask), "r"(kk_after_run)
                : "memory");
            // the predictor runs when every lane that is still decoding has a residual in hand
            if (__all_sync(0xffffffffu, cnt != 0u || i >= nc)) {
                const uint32_t have = min(cnt, 1u);                   // 0 / 1: this lane has a sample this round
                const bool hv = cnt != 0u;
                const int32_t o = lpc.step(e, j, hv);
                if (to_plane) {
                    if (is24_w) out.put32((uint32_t)o, have); else out.put16((uint32_t)o, have);
                } else {
                    int32_t L = o, Rr = 0;
                    if (kB) {
                        const int32_t A = (int32_t)(is24_w ? ar.get32(j * 4u) : ar.get16(j * 2u));
                        const int32_t un = (int32_t)((uint32_t)A - (uint32_t)((int32_t)((uint32_t)o * (uint32_t)mw) >> ms));
                        Rr = mw != 0 ? un : o;                           // AlacFile.cs:342-355, :373-380 / :359-366, :401-404
                        L = mw != 0 ? (int32_t)((uint32_t)un + (uint32_t)o) : A;
                    }
                    if (ub_w) {                                          // :381-389, :405-413, :549-554
                        const uint64_t pos = shift_pos0 + (uint64_t)j * shift_step;
                        const uint32_t sa = arena_bits(arena32, pos, (int)ub8n) & shift_mask;
                        L = (int32_t)(((uint32_t)L << ub8) | sa);
                        if (kB) {
                            const uint32_t sb = arena_bits(arena32, pos + ub8, (int)ub8n) & shift_mask;
                            Rr = (int32_t)(((uint32_t)Rr << ub8) | sb);
                        }
                    }
                    if (!is24_w) {                                        // AlacContext.cs:231-242
                        if (two_ch_w) out.put32(((uint32_t)L & 0xffffu) | ((uint32_t)Rr << 16), have);
                        else out.put16((uint32_t)L, have);
                    } else if (two_ch_w) {                                // AlacFile.cs:390-395
                        const uint32_t l = (uint32_t)L & 0xffffffu, r = (uint32_t)Rr & 0xffffffu;
                        out.put16(l, have);
                        out.put16((l >> 16) | ((r & 0xffu) << 8), have);
                        out.put16(r >> 8, have);
                    } else {                                              // :555-557
                        out.put8((uint32_t)L, have);
                        out.put8((uint32_t)L >> 8, have);
                        out.put8((uint32_t)L >> 16, have);
                    }
                }
                j += hv ? 1u : 0u;
                cnt -= have;
                e = 0;                                                // what is left in hand are zeros of a run (:238-245)
            }
        }
    }
    cp_async_wait<0>();
    if (work) {
        out.flush<7>();
        out.finish();
        uint8_t status = FS_OK;
        if (h < 0) status = FS_HISTORY;                                    // reference: garbage k
        else if (i > (uint32_t)kMaxFrameSamples) status = FS_RUN_OVERFLOW;  // reference: IndexOutOfRange
        const uint32_t end_bit = start_bit + br.consumed();
        if (end_bit > ref.len * 8u) status = FS_OVERRUN;                    // the cursor is monotone
        if (stuck) { status = FS_INTERNAL; atomicAdd(a.faults, 1u); }       // never expected
        if (!kB && stereo) a.bstart[slot] = end_bit;
        if (status != FS_OK) a.desc[f].status = status;
    }
}

// Persistent warps with a per-SM schedule.  The loop of one class is 5-10 KB of code (the taps are unrolled:
// coefficients and history live in registers), an SM's instruction cache holds 32 KB (6 KB per sub-partition),
// and a plain grid hands consecutive blocks to DIFFERENT SMs -- so every SM would run eight classes at once
// and fetch its instructions from L2 (measured: 65 % of all issue slots lost to instruction-fetch stalls, every
// one on the first instruction of a 128-byte line).  Instead the class-sorted list is cut into one segment of
// equal estimated work per SM; the warps of an SM claim one segment and walk it front to back, so the whole
// SM executes one class (two at a boundary) at a time.  A warp that finds its segment empty moves on to the
// next segment that still has work (and stays there), which evens out the tail.  Nothing ever waits on
// another warp: the schedule is only a matter of who takes which 32 frames.
template <bool kB>
__global__ void __launch_bounds__(kKfThreads, kB ? 6 : 8)
kf_frames(const ChunkArgs a, const uint32_t nseg)
{
    extern __shared__ __align__(256) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    constexpr uint32_t l = kB ? 1u : 0u;
    uint32_t *const cnt = a.kf_count;
    const uint32_t *const bound = cnt + kSegBound + l * (kMaxSeg + 1);
    uint32_t *const next = cnt + kSegNext + l * kMaxSeg;
    uint8_t *wsm = smem + (threadIdx.x >> 5) * (kB ? kKfWarpSmemA + 32u * kf_aring_stride(a.kf_row, a.ns) : kKfWarpSmemA);
    uint32_t seg = 0;
    if (lane == 0) {                                      // the segment of this SM: the first warp to arrive claims one
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        uint32_t *slot = cnt + kSmSeg + l * kMaxSm + (smid & (kMaxSm - 1u));
        uint32_t v = atomicCAS(slot, 0u, 1u);
        if (v == 0u) {
            v = atomicAdd(cnt + kSegClaim + l, 1u) % nseg + 2u;
            atomicExch(slot, v);
        } else {
            while (v < 2u) v = *reinterpret_cast<volatile uint32_t *>(slot);     // the claimer is a few instructions away
        }
        seg = v - 2u;
    }
    for (;;) {
        uint32_t w = kNoSlot;                              // next warp-sized piece of the list
        if (lane == 0) {
            for (uint32_t k = 0; k < nseg; k++) {
                uint32_t s2 = seg + k;
                if (s2 >= nseg) s2 -= nseg;
                const uint32_t lo = bound[s2], n_here = bound[s2 + 1] - lo;
                if (*reinterpret_cast<volatile uint32_t *>(next + s2) >= n_here) continue;
                const uint32_t it = atomicAdd(next + s2, 1u);
                if (it < n_here) { w = lo + it; seg = s2; break; }
            }
        }
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w == kNoSlot) break;
        uint32_t slot = kNoSlot;
        if (ALACGPU_CHECK(a.check, w * 32u + 32u <= a.kf_cap, CK_LIST)) slot = a.kf_list[(kB ? a.kf_cap : 0u) + w * 32u + (uint32_t)lane];
        const bool valid = slot != kNoSlot;
        int order = 0;
        if (valid) order = a.desc[a.f0 + slot].order[kB ? 1 : 0];
        const int M = __shfl_sync(0xffffffffu, order, 0);                  // the warp's class (lane 0 is never padding)
#define ALACGPU_KF(MM) case MM: kf_run<MM, kB>(a, slot, valid, wsm); break
        switch (M) {
            ALACGPU_KF(0); ALACGPU_KF(1); ALACGPU_KF(2); ALACGPU_KF(3); ALACGPU_KF(4); ALACGPU_KF(5); ALACGPU_KF(6);
            ALACGPU_KF(7); ALACGPU_KF(8); ALACGPU_KF(9); ALACGPU_KF(10); ALACGPU_KF(11); ALACGPU_KF(12); ALACGPU_KF(13);
            ALACGPU_KF(14); ALACGPU_KF(15); ALACGPU_KF(16); ALACGPU_KF(17); ALACGPU_KF(18); ALACGPU_KF(19); ALACGPU_KF(20);
            ALACGPU_KF(21); ALACGPU_KF(22); ALACGPU_KF(23); ALACGPU_KF(24); ALACGPU_KF(25); ALACGPU_KF(26); ALACGPU_KF(27);
            ALACGPU_KF(28); ALACGPU_KF(29); ALACGPU_KF(30); ALACGPU_KF(31);
            default: break;
        }
#undef ALACGPU_KF
        __syncwarp();
    }
}

// ---- pack-only frames (uncompressed, or failed at the header): k3's code over the work list --------
__global__ void __launch_bounds__(kK3Threads)
kf_pack_list(const ChunkArgs a)
{
    const uint32_t count = a.kf_count[2];
    const uint32_t groups = (a.max_sf + kK3PerBlock - 1) / kK3PerBlock;
    const uint32_t *list = a.kf_list + 2u * (size_t)a.kf_cap;
    for (uint32_t task = blockIdx.x; task < count * groups; task += gridDim.x) {
        const uint32_t slot = list[task / groups], g = task % groups;
        uint32_t w[12], nbytes, cnt;
        uint8_t *dst;
        if (pack_group(a, slot, (g * kK3Threads + threadIdx.x) * kK3PerThread, w, nbytes, cnt, dst))
            store_group(dst, w, nbytes, cnt);
    }
}

// Frames that failed AFTER part of their PCM had been written (entropy faults): zero PCM of the nominal
// size.  One warp per 32 frames looks, the warp zeroes the (rare) failed ones.
__global__ void __launch_bounds__(128)
kf_fix_failed(const ChunkArgs a)
{
    const uint32_t slot0 = (blockIdx.x * 128u + threadIdx.x) & ~31u;
    const int lane = threadIdx.x & 31;
    const uint32_t slot = slot0 + (uint32_t)lane;
    bool bad = false;
    if (slot < a.n) {
        const FrameDesc d = a.desc[a.f0 + slot];
        bad = d.status != FS_OK && d.status0 == FS_OK && d.out_len != 0;
    }
    uint32_t m = __ballot_sync(0xffffffffu, bad);
    while (m) {
        const int l = __ffs((int)m) - 1;
        m &= m - 1;
        const uint64_t f = a.f0 + slot0 + (uint32_t)l;
        const uint32_t len = a.desc[f].out_len;
        uint8_t *dst = a.pcm + (a.frame_off[f] - a.pcm_base);
        for (uint32_t k = (uint32_t)lane; k < len; k += 32u) dst[k] = 0;
    }
}

// one segment of the class-sorted list per SM of the current device
static uint32_t kf_segments()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); sms = 148; }
    static const int env = getenv("ALACGPU_KF_SEGMENTS") ? atoi(getenv("ALACGPU_KF_SEGMENTS")) : 0;
    if (env > 0) sms = env;
    return (uint32_t)std::min<int>(std::max(sms, 1), (int)kMaxSeg);
}

// (attributes are per device: set on the current one every time, a cheap driver call)
cudaError_t launch_kf_sort(const ChunkArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    if (cudaError_t e = cudaMemsetAsync(a.kf_count, 0, kKfCountWords * sizeof(uint32_t), st)) return e;
    if (cudaError_t e = cudaMemsetAsync(a.kf_list, 0xFF, 2u * (size_t)a.kf_cap * sizeof(uint32_t), st)) return e;
    const uint32_t nb = (a.n + 255u) / 256u;
    kf_hist<<<nb, 256, 0, st>>>(a.desc + a.f0, a.refs + a.f0, a.cfgs, a.n, a.kf_count);
    kf_scan<<<1, 256, 0, st>>>(a.kf_count, a.kf_cap, kf_segments());
    kf_scatter<<<nb, 256, 0, st>>>(a.desc + a.f0, a.refs + a.f0, a.cfgs, a.n, a.kf_count, a.kf_list, a.kf_cap);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

// The frame-lane kernels live on shared memory (three lane-private rings per lane) and barely use L1: ask for
// the largest shared-memory carve-out so that the register file, not the carve-out, bounds the blocks per SM.
template <bool kB>
static cudaError_t kf_attributes(const ChunkArgs &a, int *blocks_per_sm, int *smem_bytes)
{
    // ALACGPU_KF_PAD_A / _B (KB): extra dynamic shared memory per block, i.e. fewer resident blocks per SM (tuning runs)
    static const int pad = getenv(kB ? "ALACGPU_KF_PAD_B" : "ALACGPU_KF_PAD_A") ? atoi(getenv(kB ? "ALACGPU_KF_PAD_B" : "ALACGPU_KF_PAD_A")) * 1024 : 0;
    const int smem = 2 * (int)(kB ? kKfWarpSmemA + 32u * kf_aring_stride(a.kf_row, a.ns) : kKfWarpSmemA) + pad;
    // per device and shared-memory size the answer never changes: ask the driver once (these calls sit on the
    // caller's critical path of every decode_all)
    static std::mutex mu;
    static int known_smem[64], known_blocks[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    {
        std::lock_guard<std::mutex> g(mu);
        if (known_smem[dev] == smem && known_blocks[dev] > 0) {
            *blocks_per_sm = known_blocks[dev];
            *smem_bytes = smem;
            return cudaSuccess;
        }
    }
    if (cudaError_t e = cudaFuncSetAttribute(kf_frames<kB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) return e;
    if (cudaError_t e = cudaFuncSetAttribute(kf_frames<kB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) return e;
    int nb = 0;
    if (cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kf_frames<kB>, kKfThreads, smem)) return e;
    *blocks_per_sm = std::max(nb, 1);
    *smem_bytes = smem;
    {
        std::lock_guard<std::mutex> g(mu);
        known_smem[dev] = smem;
        known_blocks[dev] = *blocks_per_sm;
    }
    static const bool dbg = getenv("ALACGPU_DEBUG_OCC") != nullptr;
    if (dbg) {
        cudaFuncAttributes fa{};
        cudaFuncGetAttributes(&fa, kf_frames<kB>);
        fprintf(stderr, "[alacgpu] kf_frames<%c>: %d blocks/SM of %d threads, %d registers, %d B dynamic smem, %zu B local\n",
                kB ? 'B' : 'A', nb, kKfThreads, fa.numRegs, smem, fa.localSizeBytes);
    }
    return cudaSuccess;
}

cudaError_t launch_kf_a(const ChunkArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    int per_sm = 1, smem = 0;
    if (cudaError_t e = kf_attributes<false>(a, &per_sm, &smem)) return e;
    const uint32_t nseg = kf_segments();
    // persistent warps: as many blocks as fit the machine at once (fewer for a chunk that cannot fill it)
    const uint32_t blocks = std::min<uint32_t>(nseg * (uint32_t)per_sm, a.kf_cap / kKfThreads);
    kf_frames<false><<<blocks, kKfThreads, smem, st>>>(a, nseg);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_kf_b(const ChunkArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    int per_sm = 1, smem = 0;
    if (cudaError_t e = kf_attributes<true>(a, &per_sm, &smem)) return e;
    const uint32_t nseg = kf_segments();
    const uint32_t blocks = std::min<uint32_t>(nseg * (uint32_t)per_sm, a.kf_cap / kKfThreads);
    kf_frames<true><<<blocks, kKfThreads, smem, st>>>(a, nseg);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_kf_rest(const ChunkArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    kf_pack_list<<<148 * 8, kK3Threads, 0, st>>>(a);
    kf_fix_failed<<<(a.n + 127u) / 128u, 128, 0, st>>>(a);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

}  // namespace alacgpu
