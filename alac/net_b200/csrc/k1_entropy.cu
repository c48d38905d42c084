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
#include "alacgpu_device.cuh"
#include "alacgpu_kernels.h"

namespace alacgpu {

struct LaneReader {
    const uint32_t *wp;   // word holding the cursor
    uint32_t cur, nxt, nn;
    int off;              // cursor bit within cur, 0..31
    uint32_t pos;         // bits consumed since the frame start

    __device__ __forceinline__ void init(const uint32_t *arena32, uint64_t abs_bit, uint32_t frame_pos)
    {
        wp = arena32 + (abs_bit >> 5);
        off = (int)(abs_bit & 31);
        pos = frame_pos;
        cur = bswap32(__ldg(wp));
        nxt = bswap32(__ldg(wp + 1));
        nn = bswap32(__ldg(wp + 2));
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(nxt, cur, off); }
    __device__ __forceinline__ void skip(int n)   // 0 <= n <= 32
    {
        off += n;
        pos += (uint32_t)n;
        if (off >= 32) {
            off -= 32;
            ++wp;
            cur = nxt;
            nxt = nn;
            nn = bswap32(__ldg(wp + 2));
        }
    }
};

// EntropyDecodeValue (AlacFile.cs:193-212).  m = ((1 << k) - 1) & mask.
// k == 1 needs no special case: the generic path reads one bit that is always
// <= 1 and gives it back, and multiplies by m == 1.
__device__ __forceinline__ uint32_t decode_symbol(LaneReader &br, int raw_bits, int k, uint32_t m)
{
    const uint32_t w = br.peek();
    const int x = __clz((int)~w);                  // leading 1 bits
    if (x > 8) {                                   // :198-202 nine ones: raw value follows
        br.skip(9);
        const uint32_t v = br.peek() >> (32 - raw_bits);
        br.skip(raw_bits);
        return v;
    }
    const uint32_t e = (w << (x + 1)) >> (32 - k); // :205
    uint32_t v = (uint32_t)x * m;                  // :206
    int used = x + 1 + k;
    if (e > 1) v += e - 1;                         // :207-208
    else used -= 1;                                // :210 Unreadbits(1)
    br.skip(used);
    return v;
}

__global__ void __launch_bounds__(128)
k1_entropy(const ChunkArgs a, const int lanes_log2)
{
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
    const uint32_t len_bits = ref.len * 8u;
    const int ech = (d.flags & FF_STEREO) ? 2 : 1;

    LaneReader br;
    br.init(reinterpret_cast<const uint32_t *>(a.arena), ref.off * 8ull + d.data_bit, d.data_bit);

    int32_t *plane = a.planes + ((uint64_t)(slot >> 5) * 2u) * a.ns * kTile + (slot & 31);
    uint8_t status = FS_OK;

    for (int c = 0; c < ech && status == FS_OK; c++) {
        int32_t *out = plane + (uint64_t)c * a.ns * kTile;
        const int32_t mult = (int32_t)d.rice_mod[c] * (cfg.rice_history_mult / 4);   // :483
        int32_t h = cfg.rice_initial_history;                                         // :216
        uint32_t sign_mod = 0;
        uint32_t zrun = 0;
        int k;
        {
            const int t = 31 - kmod - __clz((h >> 9) + 3);                            // :221
            k = t < 0 ? t + kmod : kmod;                                              // :222
        }
        for (int i = 0; i < n; i++) {
            int32_t val = 0;
            if (zrun > 0) {
                --zrun;                                                               // :240-243, one zero per step
            } else {
                const uint32_t dv = decode_symbol(br, rss, k, (1u << k) - 1u) + sign_mod;   // :224
                if (br.pos > len_bits) { status = FS_OVERRUN; break; }
                val = (int32_t)(dv >> 1) ^ -(int32_t)(dv & 1u);                       // :225-226
                sign_mod = 0;
                if (dv > 0xFFFFu) h = 0xFFFF;                                         // :229
                else h = (int32_t)((uint32_t)h + dv * (uint32_t)mult) - ((int32_t)((uint32_t)h * (uint32_t)mult) >> 9);
                if (h < 128) {
                    if (h < 0) { status = FS_HISTORY; break; }
                    if (i + 1 < n) {                                                  // :231
                        const int kz = (h == 0 ? 40 : __clz(h)) + ((h + 16) >> 6) - 24;   // :234 (clz(0) == 40)
                        const uint32_t block = decode_symbol(br, 16, kz, ((1u << kz) - 1u) & kmask);   // :236
                        if (br.pos > len_bits) { status = FS_OVERRUN; break; }
                        if (block > 0 && (uint32_t)i + 1u + block > (uint32_t)kMaxFrameSamples) {
                            status = FS_RUN_OVERFLOW;                                 // reference: IndexOutOfRange
                            break;
                        }
                        zrun = block;
                        sign_mod = block > 0xFFFFu ? 0u : 1u;                         // :233,:246
                        h = 0;                                                        // :248
                    }
                }
                const int t = 31 - kmod - __clz((h >> 9) + 3);
                k = t < 0 ? t + kmod : kmod;
            }
            out[(uint32_t)i * kTile] = val;
        }
    }
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
    k1_entropy<<<blocks, 128, 0, st>>>(a, lg);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace alacgpu
