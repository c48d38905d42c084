// K1 -- adaptive Golomb-Rice entropy decode.
//
// Replaces Readbits/Readbit/Unreadbits (ALACDecoder/AlacFile.cs:101-152),
// CountLeadingZeros (:154-191), EntropyDecodeValue (:193-212) and
// EntropyRiceDecode (:214-252).
//
// Mapping.  A frame's Rice streams are one serial chain: every symbol's
// length depends on the running history, and channel B starts at the bit
// where channel A ends (AlacFile.cs:643 then :653 share one cursor).  So the
// unit of parallelism is the FRAME: one lane per frame, channel A then B.
// The loop is written in lock step over the OUTPUT index i -- a lane that is
// inside a zero run emits its zeros one per iteration instead of jumping
// ahead (AlacFile.cs:238-245 writes them in a burst) -- so all lanes of a warp
// are at the same i and the store of sample i is one coalesced 128-byte line
// of the tile-transposed residual plane.
//
// Bitstream access: each lane walks its own frame through two byte-swapped
// 32-bit words plus a prefetched third; a 32-bit window at the cursor is one
// funnel shift, the unary prefix is __clz(~window), and the k extra bits come
// from the same window (9 + 22 bits at most).  The reference's "read k bits,
// un-read one if the value is <= 1" (AlacFile.cs:205-210) becomes "consume
// k-1 bits".  CountLeadingZeros' clz(0)==40 quirk (AlacFile.cs:190) is kept.
//
// The common path (prefix < 9 ones, history >= 128) is kept free of error
// checks: the cursor only moves forward, so "some symbol ended past the
// frame's last bit" is decided once from the final cursor (policy: OVERRUN
// outranks a later HISTORY / RUN_OVERFLOW fault, exactly as the oracle's
// per-symbol check does).  The arena carries enough tail padding for a lane
// that runs past its frame until then.
#include "alacgpu_device.cuh"
#include "alacgpu_kernels.h"

namespace alacgpu {

// Per-lane bitstream ring in shared memory, filled by 16-byte cp.async (LDGSTS) copies:
// the copy engine writes shared memory directly, so no register ever waits on HBM.  The
// ring holds kRingChunks x 16 B of the lane's stream.  Once every kPeriod samples the warp
// tops every lane's ring up to kRingChunks chunks past its cursor (a warp-uniform branch
// with a short per-lane loop) and waits for the PREVIOUS period's copies; a period is
// ~1000 cycles, so that wait is normally free.  A lane consumes at most 59 bits per sample
// (9 ones + 25 raw bits, plus a zero-run symbol of 9 + 16), i.e. < 4 chunks per period, and
// reads at most one chunk ahead of its cursor: everything it touches during a period lies
// within cursor_chunk(previous top-up) + 4 + 4 + 1 < kRingChunks and was requested at
// least one period earlier.  Between top-ups the per-sample path only does one predicated
// 4-byte LDS when the cursor enters a new word.
constexpr int kRingChunks = 16;              // 256 B per lane, 32 KB per 128-thread block
constexpr int kRingWords = kRingChunks * 4;
constexpr int kPeriod = 8;
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

struct LaneReader {
    const uint8_t *base;    // 16-byte aligned address of chunk 0
    uint32_t ring;          // shared-space byte address of this lane's ring
    uint32_t widx;          // word index (from base) of `cur`
    uint32_t cur, nxt;      // byte-swapped words widx, widx+1
    uint32_t nn;            // raw word widx+2, read from the ring one word ahead
    uint32_t filled;        // chunks [0, filled) have been requested
    int off;                // cursor bit within cur, 0..31
    uint32_t start_bit;     // cursor position at init, relative to base

    __device__ __forceinline__ uint32_t word_addr(uint32_t w) const { return ring + ((w & (kRingWords - 1)) << 2); }

    // request every chunk up to kRingChunks past the cursor's chunk
    __device__ __forceinline__ void top_up()
    {
        const uint32_t want = (widx >> 2) + kRingChunks;
        while (filled < want) {
            cp_async16(ring + ((filled & (kRingChunks - 1)) << 4), base + ((uint64_t)filled << 4));
            ++filled;
        }
        cp_async_commit();
    }

    __device__ __forceinline__ void init(const uint8_t *arena, uint64_t abs_bit, uint32_t ring_addr)
    {
        const uint64_t byte = abs_bit >> 3;
        base = arena + (byte & ~15ull);
        ring = ring_addr;
        start_bit = (uint32_t)(byte & 15) * 8u + (uint32_t)(abs_bit & 7);
        widx = start_bit >> 5;
        off = (int)(start_bit & 31);
        filled = 0;
        top_up();
        cp_async_wait<0>();
        cur = bswap32(lds32(word_addr(widx)));
        nxt = bswap32(lds32(word_addr(widx + 1)));
        nn = lds32(word_addr(widx + 2));
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(nxt, cur, off); }

    // Skip 0 <= n <= 32 bits, branch free: entering a new word is a few selects plus one
    // predicated LDS.
    __device__ __forceinline__ void skip(int n)
    {
        off += n;
        const bool rf = off >= 32;
        off &= 31;
        widx += rf ? 1u : 0u;
        cur = rf ? nxt : cur;
        nxt = rf ? bswap32(nn) : nxt;
        const uint32_t rd = word_addr(widx + 2);
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.u32 p, %1, 0;\n\t"
            "@p ld.shared.u32 %0, [%2];\n\t"
            "}"
            : "+r"(nn)
            : "r"((uint32_t)rf), "r"(rd)
            : "memory");
    }
    // bits consumed since init
    __device__ __forceinline__ uint32_t consumed() const { return widx * 32u + (uint32_t)off - start_bit; }
};

// Rare path of EntropyDecodeValue: nine 1-bits, then the raw value (AlacFile.cs:198-202).
__device__ __forceinline__ uint32_t decode_escape(LaneReader &br, int raw_bits)
{
    br.skip(9);
    const uint32_t v = br.peek() >> (32 - raw_bits);
    br.skip(raw_bits);
    return v;
}

// EntropyDecodeValue (AlacFile.cs:193-212) with m = ((1 << k) - 1) & mask, kinv = 32 - k.
// k == 1 needs no special case: the generic path reads one bit that is always <= 1, gives
// it back, and multiplies by m == 1.
__device__ __forceinline__ uint32_t decode_symbol(LaneReader &br, int raw_bits, int k, int kinv, uint32_t m)
{
    const uint32_t w = br.peek();
    const int x = __clz((int)~w);                       // leading 1 bits
    if (__builtin_expect(x > 8, 0)) return decode_escape(br, raw_bits);
    const uint32_t e = (w << (x + 1)) >> kinv;          // :205
    const uint32_t em = max(e, 1u) - 1u;                // :207-208 (e > 1 ? e - 1 : 0)
    br.skip(x + k + (int)(min(e, 2u) >> 1));            // :210 Unreadbits(1) when e <= 1
    return (uint32_t)x * m + em;                        // :206
}

__global__ void __launch_bounds__(kK1Threads)
k1_entropy(const ChunkArgs a, const int lanes_log2)
{
    __shared__ __align__(16) uint32_t ring_smem[kRingWords * kK1Threads];
    const int lane = threadIdx.x & 31;
    const int S = 1 << lanes_log2;
    if (lane >= S) return;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t slot = gw * (uint32_t)S + (uint32_t)lane;
    if (slot >= a.n) return;
    const uint64_t f = a.f0 + slot;
    const FrameDesc d = a.desc[f];
    if (d.status != FS_OK || (d.flags & FF_ESCAPE)) return;   // escape frames are read directly by K3

    const FrameRef ref = a.refs[f];
    const TrackCfg cfg = a.cfgs[ref.track];
    const int n = d.n;
    const int rss = d.rss;
    const int kmod = cfg.rice_kmodifier;
    const uint32_t kmask = (1u << kmod) - 1u;                // AlacFile.cs:483,:643
    const int ech = (d.flags & FF_STEREO) ? 2 : 1;

    LaneReader br;
    const uint64_t abs_bit = ref.off * 8ull + d.data_bit;
    br.init(a.arena, abs_bit, (uint32_t)__cvta_generic_to_shared(ring_smem) + threadIdx.x * (kRingWords * 4u));

    int32_t *plane = a.planes + ((uint64_t)(slot >> 5) * 2u) * a.ns * kTile + (slot & 31);
    uint8_t status = FS_OK;

    for (int c = 0; c < ech && status == FS_OK; c++) {
        int32_t *out = plane + (uint64_t)c * a.ns * kTile;
        const int32_t mult = (int32_t)d.rice_mod[c] * (cfg.rice_history_mult / 4);   // :483
        int32_t h = cfg.rice_initial_history;                                         // :216
        uint32_t sign_mod = 0;
        uint32_t zrun = 0;
        int k = min(31 - __clz((h >> 9) + 3), kmod);                                  // :221-222
        for (int i = 0; i < n; i++) {
            if ((i & (kPeriod - 1)) == 0) {          // warp-uniform: lanes are in lock step on i
                br.top_up();
                cp_async_wait<1>();                   // everything but the group just committed
            }
            // One basic block for the common case: every lane evaluates the symbol at its
            // cursor; lanes inside a zero run (dec == false) commit nothing and emit 0.
            const bool dec = zrun == 0;
            const uint32_t w = br.peek();
            const int x = __clz((int)~w);                                             // leading 1 bits
            const uint32_t e = (w << (x + 1)) >> (32 - k);                            // :205
            uint32_t v = (uint32_t)x * ((1u << k) - 1u) + (max(e, 1u) - 1u);          // :206-208
            int adv = x + k + (int)(min(e, 2u) >> 1);                                 // :210
            if (__builtin_expect(dec && x > 8, 0)) {                                  // :198-202
                v = decode_escape(br, rss);
                adv = 0;
            }
            br.skip(dec ? adv : 0);
            const uint32_t dv = v + sign_mod;                                         // :224
            const int32_t val = dec ? ((int32_t)(dv >> 1) ^ -(int32_t)(dv & 1u)) : 0;  // :225-226
            const int32_t hn = (int32_t)((uint32_t)h + dv * (uint32_t)mult) - ((int32_t)((uint32_t)h * (uint32_t)mult) >> 9);
            const int32_t hd = dv > 0xFFFFu ? 0xFFFF : hn;                            // :229
            h = dec ? hd : h;
            sign_mod = dec ? 0u : sign_mod;
            zrun -= dec ? 0u : 1u;                                                    // :240-243, one zero per step
            if (__builtin_expect(dec && h < 128, 0)) {
                if (h < 0) { status = FS_HISTORY; break; }
                if (i + 1 < n) {                                                      // :231
                    const int kz = (h == 0 ? 40 : __clz(h)) + ((h + 16) >> 6) - 24;   // :234 (clz(0) == 40)
                    const uint32_t block = decode_symbol(br, 16, kz, 32 - kz, ((1u << kz) - 1u) & kmask);   // :236
                    if (block > 0 && (uint32_t)i + 1u + block > (uint32_t)kMaxFrameSamples) {
                        status = FS_RUN_OVERFLOW;                                     // reference: IndexOutOfRange
                        break;
                    }
                    zrun = block;
                    sign_mod = block > 0xFFFFu ? 0u : 1u;                             // :233,:246
                    h = 0;                                                            // :248
                }
            }
            k = min(31 - __clz((h >> 9) + 3), kmod);                                  // :221-222
            out[(uint32_t)i * kTile] = val;
        }
    }
    // The cursor is monotone: it ended past the frame iff some symbol did.
    if (d.data_bit + br.consumed() > ref.len * 8u) status = FS_OVERRUN;
    if (status != FS_OK) a.desc[f].status = status;
}

cudaError_t launch_k1(const ChunkArgs &a, int lanes_per_warp, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    int lg = 5;
    if (lanes_per_warp == 16) lg = 4;
    else if (lanes_per_warp == 8) lg = 3;
    else if (lanes_per_warp == 4) lg = 2;
    const uint32_t S = 1u << lg;
    const uint32_t warps = (a.n + S - 1) / S;
    const uint32_t blocks = (warps + 3) / 4;
    k1_entropy<<<blocks, kK1Threads, 0, st>>>(a, lg);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace alacgpu
