// K2 -- adaptive FIR predictor reconstruction (sign-LMS), in place on the
// residual planes.
//
// Replaces PredictorDecompressFirAdapt (ALACDecoder/AlacFile.cs:256-336).
//
// Mapping.  The recurrence is serial in the sample index but independent per
// (frame, channel), and K1 left the residuals tile-transposed, so one LANE
// owns one channel of one frame and a warp covers the 32 frames of a tile for
// one channel: sample i of all 32 channel-streams is one 128-byte line, read
// once and overwritten once.  (A "warp per stream, shuffle-reduce over taps"
// mapping issues about as many instructions per stream-sample but puts ~6
// dependent shuffles on every sample's critical path; see DESIGN.md.)
//
// The per-sample body is STRAIGHT-LINE code, identical for every lane:
//   * coefficients c[M] and the last M+1 outputs H[M+1] live in registers,
//     statically indexed, fully unrolled over the taps; M is the smallest
//     bucket >= the largest order among the warp's lanes (chosen per warp);
//   * a lane whose order is below M keeps H[j] == base for every j > order
//     (a masked shift), so its surplus taps see a zero difference and drop
//     out of the dot product AND of the adaptation without any predicate;
//   * the data-dependent early exit of the coefficient update
//     (AlacFile.cs:322) is the predicate "running error still positive" on a
//     sign-normalised error E = sign(err) * err;
//   * warm-up samples, delta mode (order 31) and zero residuals run the same
//     code with E = 0 and a select on the output.
#include "alacgpu_device.cuh"
#include "alacgpu_kernels.h"

namespace alacgpu {

// One tap of one sample: dot-product term + sign-LMS step, branch free.  Written in PTX so
// the update stays two predicated instructions (nvcc otherwise turns `if (E > 0)` into a
// branch per tap and sinks the operand computation into it).
//   dp  = sign * (buf[b] - buf[b+order-p])              (AlacFile.cs:324, :328)
//   acc += coef[p] * dp                                   (:303-304, sign folded out)
//   if (E > 0) { coef[p] -= sgn(dp); E -= ((|dp| + r) >> q) * (order - p); }   (:322-330)
__device__ __forceinline__ void lpc_tap(int32_t &c, int32_t &E, uint32_t &acc, const int32_t h, const int32_t nsg,
                                        const int32_t sgbase, const uint32_t r, const uint32_t q, const int32_t negm)
{
    asm("{\n\t"
        ".reg .s32 dp, a, u, t;\n\t"
        ".reg .pred act;\n\t"
        "mad.lo.s32 dp, %3, %4, %5;\n\t"
        "setp.gt.s32 act, %1, 0;\n\t"
        "mad.lo.s32 %2, %0, dp, %2;\n\t"
        "abs.s32 a, dp;\n\t"
        "max.s32 t, dp, -1;\n\t"
        "add.s32 a, a, %6;\n\t"
        "min.s32 t, t, 1;\n\t"
        "shr.u32 u, a, %7;\n\t"
        "@act sub.s32 %0, %0, t;\n\t"
        "@act mad.lo.s32 %1, u, %8, %1;\n\t"
        "}"
        : "+r"(c), "+r"(E), "+r"(acc)
        : "r"(h), "r"(nsg), "r"(sgbase), "r"(r), "r"(q), "r"(negm));
}

constexpr int kK2Threads = 128;

// All 32 lanes run this; `active` gates memory traffic only.
//   p      : lane's column of the plane (sample i at p[i * 32])
//   n      : samples of this lane's stream (0 if inactive), nmax: warp maximum
//   ord    : 1..30 general, 31 delta mode (AlacFile.cs:268-282); inactive lanes pass 31
template <int M>
__device__ __noinline__ void lpc_warp(int32_t *p, const int n, const int nmax, const int rss, const int ord,
                                      const int q, const int16_t *__restrict__ coef16, const bool active,
                                      int32_t *hist /* this thread's column of a [32][blockDim] shared ring */)
{
    const bool delta = ord == 31;
    const int ordm = delta ? 0 : ord;              // taps this lane really has
    int32_t c[M], negm[M], H[M + 1];
    uint32_t msk[M + 1];
#pragma unroll
    for (int j = 0; j < M; j++) {
        c[j] = (active && j < ordm) ? (int32_t)coef16[j] : 0;
        negm[j] = j - ordm;                        // -(order - p), AlacFile.cs:329
        asm volatile("" : "+r"(negm[j]));          // keep as a register operand of the IMAD
    }
#pragma unroll
    for (int j = 0; j <= M; j++) {
        msk[j] = (uint32_t)((ordm - j) >> 31);     // all ones iff j > order
        asm volatile("" : "+r"(msk[j]));           // keep as data: the shift below is one LOP3 per tap
        H[j] = 0;
    }

    const int32_t rnd = (int32_t)(1u << ((q - 1) & 31));            // :306 (quant 0 -> 1 << 31)
    // sign * ((val*sign) >> quant) = (|val| + r) >> quant with r = 0 for a positive error and
    // r = 2^quant - 1 for a negative one (arithmetic shift of the negated magnitude, :328-329)
    const uint32_t rneg = (1u << q) - 1u;
    const int sh = (32 - rss) & 31;

    H[0] = active ? p[0] : 0;                                       // first sample always copies (:259-260)
    hist[0] = H[0];
    // residuals are fetched two samples ahead (HBM latency ~ one sample's worth of taps)
    int32_t e_next = (active && n > 1) ? p[kTile] : 0;
    int32_t e_next2 = (active && n > 2) ? p[2 * kTile] : 0;
    for (int i = 1; i < nmax; i++) {
        const bool live = active && i < n;
        const int32_t e = e_next;
        e_next = e_next2;
        if (active && i + 2 < n) e_next2 = p[(uint32_t)(i + 2) * kTile];
        // base of the NEXT sample, o[i - ord] (ord >= 1): from the lane's ring of its last 32
        // outputs in shared memory ([slot][thread]: bank == lane, conflict free)
        const int32_t nb = hist[((uint32_t)(i - ord) & 31u) * kK2Threads];
        const bool main = !delta && i > ord;                        // warm-up covers i = 1..ord (:284-293)
        const int32_t base = H[M];                                  // o[i-1-ord]
        const int32_t nsg = e < 0 ? 1 : -1;                         // -sign(err)
        const int32_t sgbase = e < 0 ? (int32_t)(0u - (uint32_t)base) : base;
        int32_t E = main ? (e < 0 ? (int32_t)(0u - (uint32_t)e) : e) : 0;   // sign(err) * err
        uint32_t r = e < 0 ? rneg : 0u;
        uint32_t acc = 0;
#pragma unroll
        for (int pp = M - 1; pp >= 0; --pp)
            lpc_tap(c[pp], E, acc, H[pp], nsg, sgbase, r, (uint32_t)q, negm[pp]);
        const int32_t sum = (int32_t)(acc * (uint32_t)nsg);         // sum of (buf[b+order-j]-buf[b])*coef[j]
        int32_t v = (int32_t)((uint32_t)rnd + (uint32_t)sum) >> q;  // :306-307
        v = (int32_t)((uint32_t)v + (uint32_t)base + (uint32_t)e);  // :308
        const int32_t w = (int32_t)((uint32_t)H[0] + (uint32_t)e);  // warm-up / delta (:279, :288)
        const int32_t x = main ? v : w;
        const int32_t o = (int32_t)((uint32_t)x << sh) >> sh;       // :309-310
        if (live) p[(uint32_t)i * kTile] = o;
        hist[((uint32_t)i & 31u) * kK2Threads] = o;
        // masked shift: true history up to the lane's order, the new base beyond it
#pragma unroll
        for (int j = M; j > 0; --j) H[j] = (int32_t)(((uint32_t)nb & msk[j]) | ((uint32_t)H[j - 1] & ~msk[j]));
        H[0] = o;
    }
}

__global__ void __launch_bounds__(kK2Threads)
k2_lpc(const ChunkArgs a)
{
    __shared__ int32_t hist_smem[32 * kK2Threads];
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // warp = (tile, channel)
    const uint32_t tile = gw >> 1;
    const int ch = (int)(gw & 1);
    const uint32_t slot = tile * kTile + (uint32_t)lane;
    bool active = slot < a.n;
    int n = 0, rss = 32, ord = 31, q = 0;
    const int16_t *coef16 = nullptr;
    if (active) {
        const uint64_t f = a.f0 + slot;
        const FrameDesc d = a.desc[f];
        active = d.status == FS_OK && !(d.flags & FF_ESCAPE) && (ch == 0 || (d.flags & FF_STEREO)) &&
                 d.order[ch] != 0 && d.n > 1;            // order 0: output == residual (:261-267)
        if (active) {
            n = d.n; rss = d.rss; ord = d.order[ch]; q = d.quant[ch];
            coef16 = a.coefs[f].c[ch];
        }
    }
    if (!active) { n = 0; ord = 31; }
    // taps needed by this warp: delta mode (31) needs none
    const int need = active ? (ord == 31 ? 1 : ord) : 0;
    const int maxo = __reduce_max_sync(0xffffffffu, need);
    if (maxo == 0) return;
    const int nmax = __reduce_max_sync(0xffffffffu, n);
    int32_t *p = a.planes + ((uint64_t)tile * 2u + (uint32_t)ch) * a.ns * kTile + lane;
    if (maxo <= 4) lpc_warp<4>(p, n, nmax, rss, ord, q, coef16, active, hist_smem + threadIdx.x);
    else if (maxo <= 8) lpc_warp<8>(p, n, nmax, rss, ord, q, coef16, active, hist_smem + threadIdx.x);
    else if (maxo <= 12) lpc_warp<12>(p, n, nmax, rss, ord, q, coef16, active, hist_smem + threadIdx.x);
    else if (maxo <= 16) lpc_warp<16>(p, n, nmax, rss, ord, q, coef16, active, hist_smem + threadIdx.x);
    else if (maxo <= 20) lpc_warp<20>(p, n, nmax, rss, ord, q, coef16, active, hist_smem + threadIdx.x);
    else if (maxo <= 24) lpc_warp<24>(p, n, nmax, rss, ord, q, coef16, active, hist_smem + threadIdx.x);
    else lpc_warp<30>(p, n, nmax, rss, ord, q, coef16, active, hist_smem + threadIdx.x);
}

cudaError_t launch_k2(const ChunkArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    const uint32_t tiles = (a.n + kTile - 1) / kTile;
    const uint32_t warps = tiles * 2;
    k2_lpc<<<(warps + 3) / 4, kK2Threads, 0, st>>>(a);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace alacgpu
