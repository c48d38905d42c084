// K2 -- adaptive FIR predictor reconstruction (sign-LMS), in place on the
// residual planes.
//
// Replaces PredictorDecompressFirAdapt (ALACDecoder/AlacFile.cs:256-336).
//
// Mapping.  The recurrence is serial in the sample index but independent per
// (frame, channel), and K1 left the residuals tile-transposed, so one LANE
// owns one channel of one frame and a warp covers the 32 frames of a tile for
// one channel: sample i of all 32 channel-streams is one 128-byte line, read
// once and overwritten once.  Coefficients and the last M outputs live in
// registers (statically indexed, fully unrolled over the taps); M is the
// smallest bucket >= the largest order among the warp's lanes, chosen per warp.
// The data-dependent early exit of the coefficient update (AlacFile.cs:322)
// becomes a per-tap predicate: a tap adapts iff the running error still has
// its original sign.
//
// A lane-per-channel-stream mapping issues ~M multiply-adds + ~12M update
// instructions per sample per warp for 32 streams; the "warp per stream,
// shuffle-reduce over taps" alternative needs ~10 dependent shuffles per
// sample for ONE stream, i.e. ~25x fewer streams per issue slot (DESIGN.md
// "K2 mapping").
#include "alacgpu_device.cuh"
#include "alacgpu_kernels.h"

namespace alacgpu {

// One channel-stream, orders 1..30 (general) and 31 (first-order delta mode).
// p points at the lane's column of the plane: sample i is p[i * 32].
template <int M>
__device__ __forceinline__ void lpc_stream(int32_t *p, const int n, const int rss, const int ord,
                                           const int q, const int16_t *__restrict__ coef16)
{
    int32_t c[M];
    int32_t H[M];          // H[j] = o[i-1-j]
#pragma unroll
    for (int j = 0; j < M; j++) {
        c[j] = j < ord ? (int32_t)coef16[j] : 0;
        H[j] = 0;
    }
    const bool delta = ord == 31;                        // AlacFile.cs:268-282
    const int32_t rnd = (int32_t)(1u << ((q - 1) & 31)); // :306 (quant 0 -> 1 << 31)
    int32_t prev = p[0];                                 // first sample always copies (:259-260)
    H[0] = prev;
    int32_t e_next = n > 1 ? p[kTile] : 0;
    int32_t base_next = 0;
    for (int i = 1; i < n; i++) {
        const int32_t e = e_next;
        const int32_t base = base_next;                  // o[i-1-ord] (valid once i > ord)
        if (i + 1 < n) e_next = p[(uint32_t)(i + 1) * kTile];
        if (!delta && i >= ord) base_next = p[(uint32_t)(i - ord) * kTile];
        const bool main = !delta && i > ord;             // warm-up covers i = 1..ord (:284-293)
        int32_t o;
        if (!main) {
            o = sext((int32_t)((uint32_t)prev + (uint32_t)e), rss);   // :279, :288-291
        } else {
            int32_t d[M];
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < M; j++) {
                d[j] = (int32_t)((uint32_t)base - (uint32_t)H[j]);    // buf[b] - buf[b+order-j] (:324)
                acc += (uint32_t)c[j] * (uint32_t)d[j];               // == -(buf[..]-buf[b])*coef (:303-304)
            }
            const int32_t sum = (int32_t)(0u - acc);
            int32_t v = (int32_t)((uint32_t)rnd + (uint32_t)sum) >> q;            // :306-307
            v = (int32_t)((uint32_t)v + (uint32_t)base + (uint32_t)e);            // :308
            o = sext(v, rss);                                                     // :309-310
            if (e != 0) {                                                         // :312-332
                // sg = sign of the error; E = sg * err stays > 0 while the loop runs.
                // err -= ((d*sign) >> q) * (ord-p) with sign = sg*sgn(d):
                //   sg > 0: (|d| >> q);  sg < 0: -((-|d|) >> q) = (|d| + 2^q - 1) >> q.
                const int32_t sg = e > 0 ? 1 : -1;
                int32_t E = e > 0 ? e : (int32_t)(0u - (uint32_t)e);
                const int32_t r = e > 0 ? 0 : (int32_t)((1u << q) - 1u);
#pragma unroll
                for (int pp = M - 1; pp >= 0; --pp) {
                    const bool act = (pp < ord) && (E > 0);
                    const int32_t dj = d[pp];
                    const int32_t s = (dj > 0) - (dj < 0);
                    const int32_t a = dj < 0 ? (int32_t)(0u - (uint32_t)dj) : dj;
                    const int32_t u = (int32_t)(((uint32_t)a + (uint32_t)r) >> q);
                    if (act) {
                        c[pp] -= s * sg;                                          // :327
                        E = (int32_t)((uint32_t)E - (uint32_t)u * (uint32_t)(ord - pp));   // :329
                    }
                }
            }
        }
        p[(uint32_t)i * kTile] = o;
#pragma unroll
        for (int j = M - 1; j > 0; --j) H[j] = H[j - 1];
        H[0] = o;
        prev = o;
    }
}

template <int M>
__device__ __noinline__ void lpc_warp(int32_t *p, int n, int rss, int ord, int q, const int16_t *coef16, bool active)
{
    if (active) lpc_stream<M>(p, n, rss, ord, q, coef16);
}

__global__ void __launch_bounds__(128)
k2_lpc(const ChunkArgs a)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // warp = (tile, channel)
    const uint32_t tile = gw >> 1;
    const int ch = (int)(gw & 1);
    const uint32_t slot = tile * kTile + (uint32_t)lane;
    bool active = slot < a.n;
    int n = 0, rss = 0, ord = 0, q = 0;
    const int16_t *coef16 = nullptr;
    if (active) {
        const uint64_t f = a.f0 + slot;
        const FrameDesc d = a.desc[f];
        active = d.status == FS_OK && !(d.flags & FF_ESCAPE) && (ch == 0 || (d.flags & FF_STEREO)) &&
                 d.order[ch] != 0;                       // order 0: output == residual (:261-267)
        n = d.n; rss = d.rss; ord = d.order[ch]; q = d.quant[ch];
        coef16 = a.coefs[f].c[ch];
    }
    // taps needed by this warp: delta mode (31) needs none
    const int need = active ? (ord == 31 ? 1 : ord) : 0;
    int maxo = need;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxo = max(maxo, __shfl_xor_sync(0xffffffffu, maxo, o));
    if (maxo == 0) return;
    int32_t *p = a.planes + ((uint64_t)tile * 2u + (uint32_t)ch) * a.ns * kTile + lane;
    if (maxo <= 4) lpc_warp<4>(p, n, rss, ord, q, coef16, active);
    else if (maxo <= 8) lpc_warp<8>(p, n, rss, ord, q, coef16, active);
    else if (maxo <= 12) lpc_warp<12>(p, n, rss, ord, q, coef16, active);
    else if (maxo <= 16) lpc_warp<16>(p, n, rss, ord, q, coef16, active);
    else if (maxo <= 20) lpc_warp<20>(p, n, rss, ord, q, coef16, active);
    else if (maxo <= 24) lpc_warp<24>(p, n, rss, ord, q, coef16, active);
    else lpc_warp<30>(p, n, rss, ord, q, coef16, active);
}

cudaError_t launch_k2(const ChunkArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    const uint32_t tiles = (a.n + kTile - 1) / kTile;
    const uint32_t warps = tiles * 2;
    k2_lpc<<<(warps + 3) / 4, 128, 0, st>>>(a);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace alacgpu
