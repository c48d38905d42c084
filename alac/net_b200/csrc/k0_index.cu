// K0 -- frame-header pre-pass.
//
// Replaces the header section of AlacFile.DecodeFrame (ALACDecoder/AlacFile.cs
// :435-475 mono, :584-641 stereo) for every frame at once: element type,
// sample count, wasted bytes, escape flag, mix parameters, per-channel
// predictor / Rice parameters and coefficients, and the bit offsets where the
// wasted-byte block and the Rice (or raw) data start.  The PCM byte count of
// each frame follows from tag + hassize + N alone; the host runtime derives
// the same number from the same seven header bytes when it lays out the output
// (runtime.cu frame_pcm_bytes), and K0 counts any disagreement.
//
// One thread per frame.  Header bits are read through a byte-safe reader that
// returns 0 for bytes at or past the frame's stsz length, which is how the
// oracle defines truncated frames (the reference would see stale bytes of
// older frames in its 80 KiB scratch, AlacContext.cs:64,195).
#include "alacgpu_device.cuh"
#include "alacgpu_kernels.h"

namespace alacgpu {

struct SafeReader {
    const uint8_t *p;
    uint32_t len;
    uint64_t pos;   // bits
    __device__ __forceinline__ uint32_t byte(uint64_t i) const { return i < len ? (uint32_t)p[i] : 0u; }
    __device__ uint32_t get(int n)   // 0..32 bits, MSB first (Readbits, AlacFile.cs:125-129)
    {
        if (n == 0) return 0;
        const uint64_t b = pos >> 3;
        uint64_t acc = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) acc = (acc << 8) | byte(b + k);
        const int sh = 40 - (int)(pos & 7) - n;
        pos += n;
        return (uint32_t)((acc >> sh) & (n == 32 ? 0xffffffffull : ((1ull << n) - 1ull)));
    }
};

// Parses frame f and returns the number of PCM bytes it emits.
__device__ __forceinline__ uint32_t
parse_one(const uint32_t f, const uint8_t *__restrict__ arena, const FrameRef *__restrict__ refs,
          const TrackCfg *__restrict__ cfgs, FrameDesc *__restrict__ desc, FrameCoefs *__restrict__ coefs)
{
    const FrameRef ref = refs[f];
    const TrackCfg cfg = cfgs[ref.track];
    const int ss = cfg.sample_size;
    const int nch = cfg.num_channels;
    const uint32_t bytes_per_sf = (uint32_t)(ss / 8) * (uint32_t)nch;   // AlacFile.cs:19
    const uint64_t len_bits = (uint64_t)ref.len * 8;

    SafeReader br{arena + ref.off, ref.len, 0};
    FrameDesc d;
    d.data_bit = 0; d.shift_bit = 0; d.out_len = 0; d.n = 0; d.flags = 0; d.ub = 0;
    d.status = FS_OK; d.rss = 0; d.mix_shift = 0; d.mix_weight = 0;
    d.order[0] = d.order[1] = 0; d.quant[0] = d.quant[1] = 0; d.rice_mod[0] = d.rice_mod[1] = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) d.pad[k] = 0;
    FrameCoefs fc;
#pragma unroll
    for (int k = 0; k < 32; k++) { fc.c[0][k] = 0; fc.c[1][k] = 0; }

    uint32_t n = (uint32_t)cfg.max_samples_per_frame;               // :430
    const uint32_t tag = br.get(3);                                  // :435
    uint8_t status = FS_OK;
    if (tag > 1) {                                                   // :437,:577 -> :718
        status = FS_BAD_TAG;
        d.out_len = n * bytes_per_sf;
        n = 0;
    } else {
        const bool stereo = tag == 1;
        br.get(4); br.get(12);                                       // :442-443 / :584-585
        const uint32_t hassize = br.get(1);
        uint32_t ub = br.get(2);
        const uint32_t escape = br.get(1);
        if (hassize) n = br.get(32);                                 // :447-453 / :589-595
        if (n > (uint32_t)kMaxFrameSamples || (uint64_t)n * bytes_per_sf > (uint64_t)kMaxFramePcmBytes) {
            status = FS_TOO_MANY;
            n = 0;
        } else {
            d.out_len = n * bytes_per_sf;
            d.flags = (stereo ? FF_STEREO : 0) | (escape ? FF_ESCAPE : 0);
            const int ech = stereo ? 2 : 1;
            const int rss = ss - (int)ub * 8 + (stereo ? 1 : 0);     // :454 / :596
            if (!escape) {
                if (rss < 1) {
                    status = FS_BAD_RSS;
                } else {
                    d.rss = (uint8_t)rss;
                    d.ub = (uint8_t)ub;
                    const uint32_t ms = br.get(8), mw = br.get(8);   // :459-460 / :599-600
                    if (stereo) { d.mix_shift = (uint8_t)ms; d.mix_weight = (uint8_t)mw; }
                    uint32_t pred_type[2] = {0, 0};
                    for (int c = 0; c < ech; c++) {                  // :461-475 / :602-632
                        pred_type[c] = br.get(4);
                        d.quant[c] = (uint8_t)br.get(4);
                        d.rice_mod[c] = (uint8_t)br.get(3);
                        const uint32_t order = br.get(5);
                        d.order[c] = (uint8_t)order;
                        for (uint32_t j = 0; j < order; j++) fc.c[c][j] = (int16_t)br.get(16);
                    }
                    d.shift_bit = (uint32_t)br.pos;
                    br.pos += (uint64_t)n * ech * ub * 8;            // :476-482 / :634-641
                    d.data_bit = (uint32_t)br.pos;
                    if (pred_type[0] != 0 || pred_type[1] != 0) status = FS_PRED_TYPE;
                    else if ((d.order[0] == 0 || (stereo && d.order[1] == 0)) && n > 4096) status = FS_ORDER0_LONG;
                    else if (br.pos > len_bits) status = FS_OVERRUN;
                }
            } else {                                                 // :498-526 / :663-700
                d.data_bit = (uint32_t)br.pos;
                d.rss = (uint8_t)ss;
                if (br.pos + (uint64_t)n * ech * ss > len_bits) status = FS_OVERRUN;
            }
        }
    }
    d.n = (uint16_t)n;
    d.status = status;
    d.status0 = status;
    desc[f] = d;
    coefs[f] = fc;
    return d.out_len;
}

__global__ void __launch_bounds__(128)
k0_parse_headers(const uint8_t *__restrict__ arena, const FrameRef *__restrict__ refs,
                 const TrackCfg *__restrict__ cfgs, uint32_t n_frames,
                 FrameDesc *__restrict__ desc, FrameCoefs *__restrict__ coefs,
                 const uint32_t *__restrict__ expect_len, uint32_t *__restrict__ mismatch,
                 const uint64_t arena_bytes, uint32_t *__restrict__ check)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
#ifdef ALACGPU_CHECKED
    if (!ALACGPU_CHECK(check, refs[f].off + refs[f].len <= arena_bytes, CK_ARENA)) return;
#endif
    const uint32_t len = parse_one(f, arena, refs, cfgs, desc, coefs);
    if (len != expect_len[f]) atomicAdd(mismatch, 1u);   // host layout rule out of sync (never expected)
}

// ---- host launcher -----------------------------------------------------------
cudaError_t launch_k0(const K0Args &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n == 0) return cudaSuccess;
    const uint32_t nb = (a.n + 127) / 128;
    k0_parse_headers<<<nb, 128, 0, st>>>(a.arena, a.refs + a.f0, a.cfgs, a.n, a.desc + a.f0, a.coefs + a.f0,
                                         a.expect_len + a.f0, a.mismatch, a.arena_bytes, a.check);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace alacgpu
