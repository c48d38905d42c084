// K0 -- frame-header pre-pass + PCM offset scan.
//
// Replaces the header section of AlacFile.DecodeFrame (ALACDecoder/AlacFile.cs
// :435-475 mono, :584-641 stereo) for every frame at once and turns the
// demuxer's per-frame sizes into a device-resident index with PCM output
// offsets (the reference learns each frame's sample count only while decoding
// it, AlacFile.cs:447-453; AlacContext.cs:199 adds stts durations afterwards).
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

// Parses frame f and returns the number of sample-frames of PCM it emits.
__device__ __forceinline__ uint32_t
parse_one(const uint32_t f, const uint8_t *__restrict__ arena, const FrameRef *__restrict__ refs,
          const TrackCfg *__restrict__ cfgs, FrameDesc *__restrict__ desc, FrameCoefs *__restrict__ coefs,
          uint32_t *__restrict__ out_len)
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
    for (int k = 0; k < 6; k++) d.pad[k] = 0;
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
    desc[f] = d;
    coefs[f] = fc;
    out_len[f] = d.out_len;
    return bytes_per_sf ? d.out_len / bytes_per_sf : 0;
}

__global__ void __launch_bounds__(128)
k0_parse_headers(const uint8_t *__restrict__ arena, const FrameRef *__restrict__ refs,
                 const TrackCfg *__restrict__ cfgs, uint32_t n_frames,
                 FrameDesc *__restrict__ desc, FrameCoefs *__restrict__ coefs,
                 uint32_t *__restrict__ out_len, uint32_t *__restrict__ max_samples)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n_eff = 0;
    if (f < n_frames) n_eff = parse_one(f, arena, refs, cfgs, desc, coefs, out_len);
    // plane stride / K3 grid: the longest run of sample-frames any frame emits
    const uint32_t m = __reduce_max_sync(0xffffffffu, n_eff);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(max_samples, m);
}

// ---- exclusive scan of out_len (uint32) into uint64 offsets ----------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanBlock = kScanThreads * kScanItems;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t *total)
{
    __shared__ uint64_t warp_sums[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint64_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        const uint64_t s = warp_sums[w];
        if (w < warp) base += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

__global__ void __launch_bounds__(kScanThreads)
k0_scan_block_sums(const uint32_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ block_sums)
{
    const uint64_t base = (uint64_t)blockIdx.x * kScanBlock + (uint64_t)threadIdx.x * kScanItems;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) if (base + k < n) s += in[base + k];
    uint64_t tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads)
k0_scan_sums(uint64_t *__restrict__ block_sums, uint32_t n_blocks, uint64_t *__restrict__ grand_total)
{
    uint64_t carry = 0;
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += kScanThreads) {
        const uint32_t i = b0 + threadIdx.x;
        const uint64_t v = i < n_blocks ? block_sums[i] : 0;
        uint64_t tot;
        const uint64_t ex = block_exclusive_scan(v, &tot);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(kScanThreads)
k0_scan_apply(const uint32_t *__restrict__ in, uint64_t n, const uint64_t *__restrict__ block_sums,
              uint64_t *__restrict__ out)
{
    const uint64_t base = (uint64_t)blockIdx.x * kScanBlock + (uint64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { v[k] = base + k < n ? in[base + k] : 0; s += v[k]; }
    uint64_t tot;
    uint64_t ex = block_exclusive_scan(s, &tot) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
}

// unpadded PCM offset of each track's first frame (tracks with no frames get
// the offset of the next frame, or the grand total)
__global__ void k0_gather_track_starts(const uint64_t *__restrict__ frame_off, const uint64_t *__restrict__ grand_total,
                                       const uint64_t *__restrict__ track_first_frame, uint32_t n_tracks,
                                       uint64_t n_frames, uint64_t *__restrict__ track_start)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tracks) return;
    const uint64_t f = t < n_tracks ? track_first_frame[t] : n_frames;
    track_start[t] = f < n_frames ? frame_off[f] : *grand_total;
}

// ---- host launchers ---------------------------------------------------------
cudaError_t launch_k0(const K0Args &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n_frames == 0) return cudaSuccess;
    const uint32_t nb = (uint32_t)((a.n_frames + 127) / 128);
    k0_parse_headers<<<nb, 128, 0, st>>>(a.arena, a.refs, a.cfgs, (uint32_t)a.n_frames, a.desc, a.coefs, a.out_len, a.max_samples);
    const uint32_t sb = (uint32_t)((a.n_frames + kScanBlock - 1) / kScanBlock);
    k0_scan_block_sums<<<sb, kScanThreads, 0, st>>>(a.out_len, a.n_frames, a.block_sums);
    k0_scan_sums<<<1, kScanThreads, 0, st>>>(a.block_sums, sb, a.grand_total);
    k0_scan_apply<<<sb, kScanThreads, 0, st>>>(a.out_len, a.n_frames, a.block_sums, a.frame_off);
    k0_gather_track_starts<<<(a.n_tracks + 1 + 127) / 128, 128, 0, st>>>(a.frame_off, a.grand_total, a.track_first_frame,
                                                                        a.n_tracks, a.n_frames, a.track_start);
    if (launches) *launches += 5;
    return cudaGetLastError();
}

uint32_t k0_scan_blocks(uint64_t n_frames) { return (uint32_t)((n_frames + kScanBlock - 1) / kScanBlock); }

}  // namespace alacgpu
