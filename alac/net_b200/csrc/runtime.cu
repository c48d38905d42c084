// runtime.cu -- host runtime and C ABI of libalacgpu.so.
//
// Owns: the host-side frame index (the demuxer's stsz table turned into
// {arena offset, length, PCM offset, PCM length} per frame; ALACDecoder/
// DemuxResT.cs:28, AlacContext.cs:194-195), the per-device mdat arena, the
// multi-GPU frame partition, the chunked H2D -> K0 -> K1 -> K2 -> K3 -> D2H
// pipeline, and the per-frame pull that stands behind AlacContext.Read
// (ALACDecoder/AlacContext.cs:163-204).  No CPU decode path exists here: every
// sample is produced by the kernels.  The only bitstream bytes the host looks
// at are the first seven of each frame (tag / hassize / sample count), to know
// how many PCM bytes the frame will produce (AlacFile.cs:435-453 / :584-595) so
// that the output layout and every copy size are known before any kernel runs.
//
// Pipeline.  A frame is a serial chain of ~8k symbols and ~8k predictor steps,
// so a chunk of frames takes about the same time whether it holds 500 or
// 15,000 frames; throughput comes from having many frames in flight.  Chunks
// therefore run CONCURRENTLY on kSlots compute streams (each with its own
// plane buffer), fed by one H2D stream and drained by one D2H stream:
//   h2d:   copy(c0) copy(c1) copy(c2) ...
//   slot0:          K0..K3(c0)            K0..K3(c8) ...
//   slot1:                   K0..K3(c1)   ...
//   d2h:                        pcm(c0) pcm(c1) ...
#include "../../../include/alacgpu.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "alacgpu_device.cuh"
#include "alacgpu_kernels.h"
#include "host_staging.h"

using namespace alacgpu;

namespace {

constexpr uint64_t kTrackAlign = 256;     // PCM start alignment of each track in the global layout
// zero padding after the last staged byte: a lane whose frame is truncated keeps reading
// (at most 2 channels x 16384 symbols x 59 bits) until K1 flags the overrun at the end
constexpr uint64_t kArenaTail = 256 * 1024 + 256;
constexpr uint32_t kMaxChunkFrames = 32768;
constexpr int kSlots = 16;
constexpr uint32_t kFullFusionMaxFrames = 20480;   // largest chunk decoded by the fully fused launch
constexpr uint32_t kSmallBatchFrames = 10240;      // resident chunk up to here: multi-lane LPC on both channels from order 5 up (issue_chunk)
constexpr uint32_t kWideLpcMaxFrames = 1536;       // ... mono tracks, up to here: eight lanes per stream instead of four
// Frame-lane path (kf_frame.cu): one lane per frame and channel from bitstream to PCM.  A lane's task is 32
// frames x 4096 samples (~5 ms), so the path needs MANY tasks per SM before its tails stop mattering: measured
// on B200 (16-bit stereo, resident inputs) 46.6 vs 62.9 Gsamples/s for the stream-lane kernels at 174 k frames
// (an earlier r2 build), 66.1 vs 66.1 at 678 k, 77.6 vs 66.4 at 2.71 M -- with 1.7x instead of 4.2x the
// algorithmic bytes in DRAM traffic.  A device whose share of the batch holds at least this many frames takes it.
constexpr uint64_t kFrameLaneMinFrames = 650000;
constexpr uint32_t kFrameLaneMaxChunk = 8u << 20;
// bytes of channel-A planes over all slots in flight.  Resident inputs: one chunk as big as this allows (a frame
// lane's task is 32 frames x 4096 samples, ~5 ms: the fewer launches, the less of the machine idles in their
// tails); inputs streamed in: chunks of 1/16 of the frames, as many slots as fit
constexpr uint64_t kFrameLanePlaneBudget = 96ull << 30;
constexpr uint64_t kReadWindow = 8ull << 20;   // host-side cache window of alacgpu_read_frame
constexpr uint64_t kRingSlotBytes = 16ull << 20;   // one slot of the page-locked staging rings (host_staging.h)

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;   // elements
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    // exactly n elements (the multi-gigabyte planes of the frame-lane path: no growth slack)
    cudaError_t reserve_exact(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = n;
        return cudaSuccess;
    }
    // grow, keeping the first `keep` elements (the arena of a context whose earlier tracks stay resident)
    cudaError_t reserve_keep(size_t n, size_t keep)
    {
        if (n <= cap) return cudaSuccess;
        if (!p || keep == 0) return reserve(n);
        T *q = nullptr;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc(&q, want * sizeof(T));
        if (e != cudaSuccess) return e;
        e = cudaMemcpy(q, p, std::min(keep, cap) * sizeof(T), cudaMemcpyDeviceToDevice);
        if (e != cudaSuccess) { cudaFree(q); return e; }
        cudaFree(p);
        p = q;
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct HostTrack {
    alacgpu_track_cfg cfg;
    const uint8_t *mdat;          // borrowed until the track's bytes are in HBM (the first prepare() / decode_all()
                                  // after the add); nulled then -- the library never goes back to host memory
    bool staged = false;          // bytes are resident in the arena (single-device contexts keep them across later adds)
    uint64_t mdat_len;
    uint64_t first_frame_offset;
    uint64_t first_frame;         // index into the global frame list
    uint32_t n_frames;
    uint64_t pcm_off, pcm_len;    // global (padded) layout
    std::vector<uint64_t> offs;   // explicit byte offset of every frame (empty: frames are back to back)
};

struct HostCopy { const uint8_t *src; uint64_t dst, len; bool pinned; };   // pinned: the caller's memory is page-locked

struct Chunk {
    uint64_t f0;                  // device-local first frame
    uint32_t n;
    uint32_t copy_lo, copy_hi;    // range in Device::copies
    uint64_t pcm_lo, pcm_hi;      // global PCM byte range produced by this chunk
};

struct Slot {
    cudaStream_t st = nullptr;
    DevBuf<int32_t> planes;
    DevBuf<uint32_t> perm;        // K2 work list (2 entries per frame) + 1 count word at the end
    DevBuf<uint32_t> progress;    // fused launch: [0,2cf) entropy->LPC words, [2cf,4cf) LPC->pack words, 4 words of
                                  // pack task counter, then 2cf bytes of LPC work-list flags (cf = chunk frames)
    DevBuf<uint32_t> kf;          // frame-lane path: work lists, sort counters, channel-B start bits
};

struct Device {
    int id = 0;
    cudaStream_t st_h2d = nullptr, st_d2h = nullptr;
    Slot slots[kSlots];
    // shard = global frames [f_lo, f_hi)
    uint64_t f_lo = 0, f_hi = 0;
    DevBuf<uint8_t> arena;
    uint64_t arena_used = 0;
    uint64_t arena_staged = 0;        // bytes [0, arena_staged) hold tracks whose host memory is no longer borrowed
    DevBuf<FrameRef> refs;
    DevBuf<TrackCfg> cfgs;
    DevBuf<FrameDesc> desc;
    DevBuf<FrameCoefs> coefs;
    DevBuf<uint32_t> expect_len;
    DevBuf<uint64_t> frame_off;
    DevBuf<uint64_t> scalars;         // [0] K0 size mismatches (u32), [2] checksum
    DevBuf<uint8_t> pcm;
    uint64_t pcm_lo = 0, pcm_hi = 0;  // global byte range held by `pcm` (pcm_lo 256-aligned)
    uint64_t pcm_first = 0;           // global offset of the first byte this shard produces
    uint32_t ns = 64;                 // plane row stride (samples): multiple of 32, plus 32 so rows are not 16 KiB apart
    uint32_t kf_row = 256;            // frame-lane path: bytes per frame of the channel-A plane (ns x 2 without 24-bit tracks)
    std::vector<FrameRef> h_refs;
    std::vector<HostCopy> track_copies;   // one per (track, device): host bytes -> arena
    std::vector<HostCopy> copies;         // the same bytes split at chunk boundaries
    std::vector<Chunk> chunks;
    uint32_t chunk_frames = 0;
    bool chunk_taper = false;
    PinnedRing ring_in, ring_out;     // page-locked staging for pageable callers (host_staging.h)
    std::unique_ptr<Progress> h2d_prog, k_prog;   // stager -> issuer -> drainer hand-over of recorded events
    // chunk size and slot count of the frame-lane path per mode ([0] resident inputs, [1] streamed in), decided once
    // per plan: the decision asks the driver for free memory (cudaMemGetInfo), which was seen to take 30-100 ms
    // now and then -- not something to do in every decode_all
    mutable uint32_t cf_cache[2] = {0, 0};
    int slots_cache[2] = {0, 0};
    bool frame_lanes = false;         // this pipeline run decodes with the frame-lane kernels
    int slots_n = kSlots;             // slot streams in use by this pipeline run
    std::vector<cudaEvent_t> events;
    bool resident = false;            // arena bytes + K0 results are on the device
    bool decoded = false;             // kernels ran: per-frame status is valid
    bool pcm_resident = false;        // ... and the PCM is in `pcm` (not after a zero-copy decode)
};

}  // namespace

struct alacgpu_ctx {
    std::vector<Device> devs;
    std::vector<HostTrack> tracks;
    std::vector<uint32_t> sizes;          // stsz of every frame, track-major
    std::vector<uint32_t> out_len;        // PCM bytes of every frame (host rule == K0 rule)
    std::vector<uint64_t> frame_off;      // global padded PCM offset of every frame
    alacgpu_opts opts{};
    std::unique_ptr<CopyPool> pool;       // host threads that move bytes between caller memory and the staging rings
    std::mutex pool_mutex;
    bool planned = false;
    uint64_t total_pcm = 0;
    uint64_t compressed_bytes = 0;
    uint32_t max_sf = 0;                  // most sample-frames any frame emits
    bool mono_only = true;                // every track's container has one channel
    uint32_t index_launches = 0;
    alacgpu_timing timing{};
    // stage timings of the last pipeline are read back from its CUDA events on demand (alacgpu_get_timing):
    // ~80 event queries are not on the caller's critical path
    uint32_t internal_faults = 0;         // frames flagged FS_INTERNAL by the last pipeline (a hand-off that timed out)
    bool index_stale = false;             // alacgpu_reindex: K0 runs again inside the next decode_all
    bool timing_pending = false;
    bool tp_stage = false, tp_index = false, tp_decode = false, tp_d2h = false, tp_zc = false;
    std::string err;
    std::vector<int32_t> h_status;
    bool have_status = false;
    uint8_t *window = nullptr;            // pinned
    uint64_t win_lo = 0, win_hi = 0;
};

namespace {

int32_t fail(alacgpu_ctx *c, int32_t code, const char *what, cudaError_t e = cudaSuccess)
{
    if (c) {
        char buf[512];
        if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
        else snprintf(buf, sizeof buf, "%s", what);
        c->err = buf;
    }
    return code;
}

#define CU(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? ALACGPU_ERR_OUT_OF_MEMORY \
                                                              : ALACGPU_ERR_CUDA,        \
                        #call, e_);                                                      \
    } while (0)

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

cudaEvent_t get_event(Device &d, size_t i)
{
    while (d.events.size() <= i) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        d.events.push_back(e);
    }
    return d.events[i];
}

// PCM bytes frame `p[0..len)` decodes to -- the same rule K0 applies on the device
// (k0_index.cu parse_one): element tag, hassize and the 32-bit sample count are the first
// 23 (+32) bits of the frame (AlacFile.cs:435-453 / :584-595); bytes past `len` read as 0.
uint32_t frame_pcm_bytes(const alacgpu_track_cfg &cfg, const uint8_t *p, uint64_t len)
{
    uint8_t b[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 7 && (uint64_t)i < len; i++) b[i] = p[i];
    const uint32_t bpsf = (uint32_t)(cfg.sample_size / 8) * (uint32_t)cfg.num_channels;
    const uint32_t tag = b[0] >> 5;
    uint32_t n = (uint32_t)cfg.max_samples_per_frame;
    if (tag > 1) return n * bpsf;                                   // AlacFile.cs:436-437,:577,:718
    const uint32_t hassize = (b[2] >> 4) & 1u;                      // bit 19
    if (hassize)                                                    // bits 23..54
        n = ((uint32_t)(b[2] & 1u) << 31) | ((uint32_t)b[3] << 23) | ((uint32_t)b[4] << 15) |
            ((uint32_t)b[5] << 7) | ((uint32_t)b[6] >> 1);
    if (n > (uint32_t)kMaxFrameSamples || (uint64_t)n * bpsf > (uint64_t)kMaxFramePcmBytes) return 0;
    return n * bpsf;
}

// which track holds global frame g (tracks sorted by first_frame)
size_t track_of(const alacgpu_ctx *ctx, uint64_t g)
{
    size_t lo = 0, hi = ctx->tracks.size();
    while (hi - lo > 1) {
        size_t mid = (lo + hi) / 2;
        if (ctx->tracks[mid].first_frame <= g) lo = mid; else hi = mid;
    }
    while (lo + 1 < ctx->tracks.size() && ctx->tracks[lo].n_frames == 0) lo++;
    return lo;
}

int lanes_for(const alacgpu_ctx *ctx)
{
    const uint32_t o = ctx->opts.entropy_lanes;
    if (o == 4 || o == 8 || o == 16 || o == 32) return (int)o;
    return 32;
}

void invalidate(alacgpu_ctx *ctx)
{
    ctx->planned = false;
    ctx->have_status = false;
    ctx->win_lo = ctx->win_hi = 0;
    for (Device &d : ctx->devs) {
        d.resident = false; d.decoded = false; d.pcm_resident = false; d.chunk_frames = 0; d.chunks.clear();
        d.cf_cache[0] = d.cf_cache[1] = 0; d.slots_cache[0] = d.slots_cache[1] = 0;
    }
}

// Is `p` page-locked (cudaHostAlloc / cudaHostRegister / alacgpu_host_alloc) memory?
bool is_pinned(const void *p)
{
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost) return true;
    cudaGetLastError();
    return false;
}

// Partition + per-device frame index.  Pure host work plus allocations and the small table
// uploads (on the H2D stream, ahead of any mdat copy).
int32_t build_plan(alacgpu_ctx *ctx)
{
    if (ctx->planned) return ALACGPU_OK;
    const uint64_t n_frames = ctx->sizes.size();
    const uint32_t n_tracks = (uint32_t)ctx->tracks.size();
    const int n_dev = (int)ctx->devs.size();
    std::vector<uint64_t> cut(n_dev + 1);
    alacgpu_plan_partition(ctx->sizes.data(), n_frames, n_dev, cut.data());
    uint64_t compressed = 0;
    std::vector<TrackCfg> cfgs(std::max<uint32_t>(n_tracks, 1));
    for (uint32_t t = 0; t < n_tracks; t++) {
        const alacgpu_track_cfg &h = ctx->tracks[t].cfg;
        TrackCfg c{};
        c.sample_size = h.sample_size; c.num_channels = h.num_channels;
        c.max_samples_per_frame = h.max_samples_per_frame;
        c.rice_history_mult = h.rice_history_mult; c.rice_initial_history = h.rice_initial_history;
        c.rice_kmodifier = h.rice_kmodifier;
        cfgs[t] = c;
    }
    for (int g = 0; g < n_dev; g++) {
        Device &d = ctx->devs[g];
        d.f_lo = cut[g];
        d.f_hi = cut[g + 1];
        d.resident = d.decoded = d.pcm_resident = false;
        d.chunks.clear();
        d.chunk_frames = 0;
        d.cf_cache[0] = d.cf_cache[1] = 0; d.slots_cache[0] = d.slots_cache[1] = 0;
        const uint64_t n_local = d.f_hi - d.f_lo;
        CU(cudaSetDevice(d.id));
        d.h_refs.assign(n_local, FrameRef{});
        d.track_copies.clear();
        uint64_t used = 0;
        for (uint32_t t = 0; t < n_tracks; t++) {
            const HostTrack &ht = ctx->tracks[t];
            const uint64_t a = std::max<uint64_t>(ht.first_frame, d.f_lo);
            const uint64_t b = std::min<uint64_t>(ht.first_frame + ht.n_frames, d.f_hi);
            if (a >= b) continue;
            // byte offset of every frame of [a, b) within the track's buffer: back to back from
            // first_frame_offset (AlacContext.cs:194-195), or the caller's explicit table
            std::vector<uint64_t> foff(b - a);
            if (ht.offs.empty()) {
                uint64_t off = ht.first_frame_offset;
                for (uint64_t f = ht.first_frame; f < a; f++) off += ctx->sizes[f];
                for (uint64_t f = a; f < b; f++) { foff[f - a] = off; off += ctx->sizes[f]; }
            } else {
                for (uint64_t f = a; f < b; f++) foff[f - a] = ht.offs[f - ht.first_frame];
            }
            uint64_t src_lo = ht.mdat_len, src_hi = 0;     // staged span: covers every frame (and any gap between them)
            for (uint64_t f = a; f < b; f++) {
                const uint64_t o = std::min(foff[f - a], ht.mdat_len);
                const uint64_t e = std::min(foff[f - a] + ctx->sizes[f], ht.mdat_len);
                if (e > o) { src_lo = std::min(src_lo, o); src_hi = std::max(src_hi, e); }
            }
            if (src_hi < src_lo) src_lo = src_hi = 0;
            const uint64_t base = align_up(used, 16);
            for (uint64_t f = a; f < b; f++) {
                FrameRef r;
                const uint64_t cur = foff[f - a];
                const uint64_t avail = cur < ht.mdat_len ? ht.mdat_len - cur : 0;
                r.len = (uint32_t)std::min<uint64_t>(ctx->sizes[f], avail);   // short read (MyStream.cs:47-52)
                r.off = r.len ? base + (cur - src_lo) : base;
                r.track = t;
                d.h_refs[f - d.f_lo] = r;
                compressed += r.len;
            }
            if (src_hi > src_lo && !ht.staged) d.track_copies.push_back({ht.mdat + src_lo, base, src_hi - src_lo, is_pinned(ht.mdat)});
            used = base + (src_hi - src_lo);
        }
        d.arena_used = used;
        d.ns = std::max<uint32_t>((ctx->max_sf + 31u) & ~31u, 32u) + 32u;
        {
            bool any24 = false;
            for (uint32_t t = 0; t < n_tracks; t++) {
                const HostTrack &ht = ctx->tracks[t];
                if (ht.first_frame < d.f_hi && ht.first_frame + ht.n_frames > d.f_lo && ht.cfg.sample_size == 24) any24 = true;
            }
            d.kf_row = d.ns * (any24 ? 4u : 2u);
        }
        // PCM byte range of the shard (contiguous in the global layout)
        if (n_local) {
            d.pcm_first = ctx->frame_off[d.f_lo];
            d.pcm_lo = d.pcm_first / kTrackAlign * kTrackAlign;
            d.pcm_hi = ctx->frame_off[d.f_hi - 1] + ctx->out_len[d.f_hi - 1];
        } else {
            d.pcm_first = d.pcm_lo = d.pcm_hi = 0;
        }
        CU(d.arena.reserve_keep(used + kArenaTail, d.arena_staged));
        CU(d.refs.reserve(std::max<uint64_t>(n_local, 1)));
        CU(d.cfgs.reserve(cfgs.size()));
        CU(d.desc.reserve(std::max<uint64_t>(n_local, 1)));
        CU(d.coefs.reserve(std::max<uint64_t>(n_local, 1)));
        CU(d.expect_len.reserve(std::max<uint64_t>(n_local, 1)));
        CU(d.frame_off.reserve(std::max<uint64_t>(n_local, 1)));
        CU(d.scalars.reserve(4));
        CU(d.pcm.reserve(d.pcm_hi - d.pcm_lo + 64));
        // small tables first, on the same stream the mdat copies will use
        CU(cudaMemsetAsync(d.arena.p + used, 0, kArenaTail, d.st_h2d));
        if (n_local) {
            CU(cudaMemcpyAsync(d.refs.p, d.h_refs.data(), n_local * sizeof(FrameRef), cudaMemcpyHostToDevice, d.st_h2d));
            CU(cudaMemcpyAsync(d.expect_len.p, ctx->out_len.data() + d.f_lo, n_local * sizeof(uint32_t), cudaMemcpyHostToDevice, d.st_h2d));
            CU(cudaMemcpyAsync(d.frame_off.p, ctx->frame_off.data() + d.f_lo, n_local * sizeof(uint64_t), cudaMemcpyHostToDevice, d.st_h2d));
        }
        CU(cudaMemcpyAsync(d.cfgs.p, cfgs.data(), cfgs.size() * sizeof(TrackCfg), cudaMemcpyHostToDevice, d.st_h2d));
        CU(cudaMemsetAsync(d.scalars.p, 0, 4 * sizeof(uint64_t), d.st_h2d));
        // alignment gaps between tracks (and below the shard's first byte) are zero bytes
        if (n_local) {
            const size_t t0 = track_of(ctx, d.f_lo), t1 = track_of(ctx, d.f_hi - 1);
            if (d.pcm_first > d.pcm_lo) CU(cudaMemsetAsync(d.pcm.p, 0, d.pcm_first - d.pcm_lo, d.st_h2d));
            for (size_t t = t0; t < t1; t++) {
                const uint64_t end = ctx->tracks[t].pcm_off + ctx->tracks[t].pcm_len, nxt = ctx->tracks[t + 1].pcm_off;
                if (nxt > end && end >= d.pcm_lo && nxt <= d.pcm_hi)
                    CU(cudaMemsetAsync(d.pcm.p + (end - d.pcm_lo), 0, nxt - end, d.st_h2d));
            }
        }
    }
    // `cfgs` is a local and the tables must precede every kernel: finish the uploads now
    for (Device &d : ctx->devs) {
        CU(cudaSetDevice(d.id));
        CU(cudaStreamSynchronize(d.st_h2d));
    }
    ctx->compressed_bytes = compressed;
    ctx->planned = true;
    return ALACGPU_OK;
}

// Split the device's frames into chunks of `cf` frames and split the host->arena copies at
// the chunk boundaries, so chunk c never waits for bytes of chunk c+1.
// `taper`: the first chunks are smaller (cf/4, cf/2, 3cf/4): while chunks stream in from the host the
// call ends at (first PCM ready) + (D2H of all PCM), so the first chunk should be on the GPU early.
void build_chunks(alacgpu_ctx *ctx, Device &d, uint32_t cf, bool taper)
{
    if (d.chunk_frames == cf && d.chunk_taper == taper && !d.chunks.empty()) return;
    d.chunks.clear();
    d.copies.clear();
    d.chunk_taper = taper;
    const uint64_t n_local = d.f_hi - d.f_lo;
    uint32_t step = cf;
    for (uint64_t f0 = 0; f0 < n_local; f0 += step) {
        step = cf;
        if (taper && d.chunks.size() < 3) step = std::max<uint32_t>(64u, ((cf * (uint32_t)(d.chunks.size() + 1) / 4u) + 31u) & ~31u);
        Chunk c{};
        c.f0 = f0;
        c.n = (uint32_t)std::min<uint64_t>(step, n_local - f0);
        c.copy_lo = (uint32_t)d.copies.size();
        uint64_t f = f0;
        while (f < f0 + c.n) {                      // runs of frames of one track inside the chunk
            const uint32_t t = d.h_refs[f].track;
            uint64_t e = f;
            uint64_t lo = ~0ull, hi = 0;            // arena bytes of the run (empty frames carry none)
            while (e < f0 + c.n && d.h_refs[e].track == t) {
                const FrameRef &r = d.h_refs[e];
                if (r.len) { lo = std::min<uint64_t>(lo, r.off); hi = std::max<uint64_t>(hi, r.off + r.len); }
                e++;
            }
            if (hi > lo)
                for (const HostCopy &hc : d.track_copies)       // one staged range per (track, device)
                    if (lo >= hc.dst && lo < hc.dst + hc.len) {
                        d.copies.push_back({hc.src + (lo - hc.dst), lo, std::min(hi, hc.dst + hc.len) - lo, hc.pinned});
                        break;
                    }
            f = e;
        }
        c.copy_hi = (uint32_t)d.copies.size();
        c.pcm_lo = ctx->frame_off[d.f_lo + f0];
        c.pcm_hi = ctx->frame_off[d.f_lo + f0 + c.n - 1] + ctx->out_len[d.f_lo + f0 + c.n - 1];
        d.chunks.push_back(c);
    }
    d.chunk_frames = cf;
}

constexpr size_t kEvPerChunk = 6;   // start, after K0, K1, K2, K3, h2d-done
constexpr size_t kEvBase = 4;       // [0] pipeline start, [1] pipeline end, [2] d2h start, [3] d2h end

// Issue one chunk's kernels on its slot stream.
// (errors go to *err, not to ctx->err: several devices issue from their own threads)
#define CUI(call)                                                                                \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            if (err) { *err = #call; *err += ": "; *err += cudaGetErrorString(e_); }             \
            return e_ == cudaErrorMemoryAllocation ? ALACGPU_ERR_OUT_OF_MEMORY : ALACGPU_ERR_CUDA; \
        }                                                                                        \
    } while (0)

int32_t issue_chunk(alacgpu_ctx *ctx, Device &d, const Chunk &c, Slot &s, bool with_k0, bool with_decode,
                    size_t ev, uint32_t *launches, uint8_t *pcm_override, bool streaming, bool early, std::string *err)
{
    ChunkArgs ca{};
    ca.arena = d.arena.p; ca.refs = d.refs.p; ca.cfgs = d.cfgs.p; ca.desc = d.desc.p; ca.coefs = d.coefs.p;
    ca.frame_off = d.frame_off.p; ca.planes = s.planes.p; ca.pcm = d.pcm.p; ca.pcm_base = d.pcm_lo;
    ca.ns = d.ns; ca.f0 = c.f0; ca.n = c.n; ca.max_sf = ctx->max_sf; ca.kf_row = d.kf_row;
    ca.perm = s.perm.p; ca.perm_count = s.perm.p + 4u * (size_t)d.chunk_frames + 1024u;
    ca.faults = reinterpret_cast<uint32_t *>(d.scalars.p) + 1;
    ca.check = reinterpret_cast<uint32_t *>(d.scalars.p) + 2;          // ALACGPU_CHECKED build: extents of every buffer
    ca.arena_bytes = d.arena_used + kArenaTail;
    ca.plane_bytes = (uint64_t)s.planes.cap * sizeof(int32_t);
    ca.pcm_bytes = pcm_override ? ctx->total_pcm : d.pcm_hi - d.pcm_lo + 64;
    {
        // multi-lane LPC (k2_lpc.cuh) for small, latency-bound chunks.  A mid-size resident batch (configs[1]):
        // the last channel's streams from order 17 up get four lanes (final r1 kernels: 2.65 ms; from 25 up 2.61,
        // from 21 up 3.24, from 13 up 3.00, none 2.86, both channels from 17 up 3.1 ms -- the extra warps slow the
        // entropy lanes down, and which blocks end up sharing an SM matters as much as the threshold).  A SMALL
        // resident batch (configs[0], [2]; up to kSmallBatchFrames) leaves schedulers idle and is bound by
        // ONE stream's chain -- a lone warp's time per sample is its instruction count -- so both channels get
        // four lanes from order 5 up (r2 sweep, 16-bit stereo, 646 / 1,292 / 2,584 / 3,876 / 5,168 / 7,752 / 10,336
        // frames: 1.27 / 1.29 / 1.31 / 1.48 / 1.74 / 2.09 / 2.32 ms against 1.71 / 1.67 / 1.71 / ~1.72 / 1.74 / 2.42 /
        // 2.42 with the mid-size rule; 15,504 / 19,380 frames: 3.50 / 3.96 against 2.97 / 3.14; configs[1], 14,063
        // frames of 24-bit material: 3.21 against 2.73), and eight lanes when the tracks are mono and the batch is
        // tiny (646 / 1,292 frames: 0.67 / 0.67 ms against 0.71 / 0.72 with four).  While chunks stream in from
        // the host the GPU has slack and the first PCM should leave as early as possible: both channels from 17
        // up (end to end 9.9 -> 9.65 ms).
        static const int q_last = getenv("ALACGPU_QUAD_MIN_LAST") ? atoi(getenv("ALACGPU_QUAD_MIN_LAST")) : -1;
        static const int q_first = getenv("ALACGPU_QUAD_MIN_FIRST") ? atoi(getenv("ALACGPU_QUAD_MIN_FIRST")) : -1;
        static const int q_early = getenv("ALACGPU_QUAD_MIN_EARLY") ? atoi(getenv("ALACGPU_QUAD_MIN_EARLY")) : 0;
        static const int q_wide = getenv("ALACGPU_LPC_WIDE") ? atoi(getenv("ALACGPU_LPC_WIDE")) : -1;
        static const uint32_t small_max = getenv("ALACGPU_SMALL_BATCH_FRAMES") ? (uint32_t)atoi(getenv("ALACGPU_SMALL_BATCH_FRAMES")) : kSmallBatchFrames;
        int ql = 17, qf = streaming ? 17 : 0, wide = 0;
        if (!streaming && c.n <= small_max) {
            ql = qf = 5;
            wide = ctx->mono_only && c.n <= kWideLpcMaxFrames;
        }
        if (q_last >= 0) ql = q_last;
        if (q_first >= 0) qf = q_first;
        if (q_wide >= 0) wide = q_wide;
        if (early && q_early > 0) ql = qf = q_early;          // the first chunks of a streamed call run on an idle GPU
        ca.use_quads = (c.n <= kFullFusionMaxFrames && !(ctx->opts.flags & ALACGPU_FLAG_NO_QUAD_LPC))
                           ? ((ql & 255) | ((qf & 255) << 8) | (wide ? 1 << 16 : 0)) : 0;
    }
    const size_t cf2 = 2u * (size_t)d.chunk_frames;
    ca.progress = s.progress.p;
    ca.lpc_done = s.progress.p + cf2;
    ca.pack_next = s.progress.p + 2u * cf2;
    ca.lpc_flag = reinterpret_cast<uint8_t *>(s.progress.p + 2u * cf2 + 4u);
    if (d.frame_lanes) {
        ca.kf_cap = kf_list_cap(d.chunk_frames);
        ca.kf_list = s.kf.p;
        ca.kf_count = s.kf.p + kf_list_words(d.chunk_frames);
        ca.bstart = ca.kf_count + kKfCountWords;
    }
    CUI(cudaEventRecord(get_event(d, ev), s.st));
    if (with_k0) {
        K0Args ka{};
        ka.arena = d.arena.p; ka.refs = d.refs.p; ka.cfgs = d.cfgs.p; ka.desc = d.desc.p; ka.coefs = d.coefs.p;
        ka.expect_len = d.expect_len.p; ka.mismatch = reinterpret_cast<uint32_t *>(d.scalars.p);
        ka.f0 = c.f0; ka.n = c.n;
        ka.arena_bytes = ca.arena_bytes; ka.check = ca.check;
        CUI(launch_k0(ka, s.st, launches));
    }
    CUI(cudaEventRecord(get_event(d, ev + 1), s.st));
    // fusion level: 2 = entropy + LPC + pack in one launch, 1 = entropy + LPC with K3 apart, 0 = three
    // kernels.  Measured on configs[1] (one B200): level 1 is the fastest way to PCM in HBM (3.54 ms;
    // the pack blocks of level 2 only get SM slots once producers retire: 4.23 ms) and while chunks are
    // still being streamed in from the host (11.7 vs 12.1 ms end to end).  Level 2 pays when the PCM
    // can stream straight into a page-locked destination while the batch is still decoding (resident
    // inputs -> host PCM in 8.4 ms instead of kernels + 6.3 ms of D2H), so it is used exactly then.
    int fused = (ctx->opts.flags & ALACGPU_FLAG_NO_FUSION) ? 0 : (ctx->opts.flags & ALACGPU_FLAG_NO_PACK_FUSION) ? 1 : 2;
    if (fused == 2 && !pcm_override && !(ctx->opts.flags & ALACGPU_FLAG_FORCE_PACK_FUSION)) fused = 1;
    if (pcm_override) { ca.pcm = pcm_override; ca.pcm_base = 0; }      // PCM straight into host-mapped memory
    if (d.frame_lanes) {
        // frame-lane kernels: sort + phase A | phase B | pack-only frames + fix-up (timing slots: entropy, lpc, stereo)
        if (with_decode) { CUI(launch_kf_sort(ca, s.st, launches)); CUI(launch_kf_a(ca, s.st, launches)); }
        CUI(cudaEventRecord(get_event(d, ev + 2), s.st));
        if (with_decode) CUI(launch_kf_b(ca, s.st, launches));
        CUI(cudaEventRecord(get_event(d, ev + 3), s.st));
        if (with_decode) CUI(launch_kf_rest(ca, s.st, launches));
        CUI(cudaEventRecord(get_event(d, ev + 4), s.st));
        return ALACGPU_OK;
    }
    if (with_decode) {
        CUI(launch_sort(ca, s.st, launches));             // after K0: the work list needs only the headers
        if (fused) CUI(cudaMemsetAsync(s.progress.p, 0, (2u * cf2 + 4u) * sizeof(uint32_t), s.st));
        if (fused == 2) CUI(launch_k123(ca, lanes_for(ctx), s.st, launches));
        else if (fused == 1) CUI(launch_k12(ca, lanes_for(ctx), s.st, launches));
        else CUI(launch_k1(ca, lanes_for(ctx), s.st, launches));
    }
    CUI(cudaEventRecord(get_event(d, ev + 2), s.st));
    if (with_decode && fused == 0) CUI(launch_k2(ca, s.st, launches));
    CUI(cudaEventRecord(get_event(d, ev + 3), s.st));
    if (with_decode) {
        if (fused == 2) CUI(launch_fix(ca, s.st, launches));
        else CUI(launch_k3(ca, s.st, launches));
    }
    CUI(cudaEventRecord(get_event(d, ev + 4), s.st));
    return ALACGPU_OK;
}

// Stage timings of the last pipeline, from the CUDA events its streams recorded (max over devices).
void finish_timing(alacgpu_ctx *ctx)
{
    if (!ctx->timing_pending) return;
    ctx->timing_pending = false;
    const bool stage = ctx->tp_stage, index = ctx->tp_index, decode = ctx->tp_decode, to_host = ctx->tp_d2h && !ctx->tp_zc;
    float k0 = 0, k1 = 0, k2 = 0, k3 = 0, kall = 0, d2h = 0, h2d = 0;
    for (Device &d : ctx->devs) {
        if (d.f_hi == d.f_lo || d.chunks.empty()) continue;
        cudaSetDevice(d.id);
        float s0 = 0, s1 = 0, s2 = 0, s3 = 0, ms = 0;
        for (size_t ci = 0; ci < d.chunks.size(); ci++) {
            const size_t ev = kEvBase + ci * kEvPerChunk;
            cudaEventElapsedTime(&ms, d.events[ev], d.events[ev + 1]); s0 += ms;
            cudaEventElapsedTime(&ms, d.events[ev + 1], d.events[ev + 2]); s1 += ms;
            cudaEventElapsedTime(&ms, d.events[ev + 2], d.events[ev + 3]); s2 += ms;
            cudaEventElapsedTime(&ms, d.events[ev + 3], d.events[ev + 4]); s3 += ms;
        }
        static const bool dump = getenv("ALACGPU_HOST_TIMING") != nullptr;
        if (dump) {            // debug: the pipeline's schedule, ms from its start
            for (size_t ci = 0; ci < d.chunks.size(); ci++) {
                const size_t ev = kEvBase + ci * kEvPerChunk;
                float a = 0, b = 0, c2 = 0, e2 = 0, h = 0;
                if (stage) cudaEventElapsedTime(&h, d.events[0], d.events[ev + 5]);
                cudaEventElapsedTime(&a, d.events[0], d.events[ev]);
                cudaEventElapsedTime(&b, d.events[0], d.events[ev + 1]);
                cudaEventElapsedTime(&c2, d.events[0], d.events[ev + 2]);
                cudaEventElapsedTime(&e2, d.events[0], d.events[ev + 4]);
                fprintf(stderr, "[alacgpu] chunk %2zu n %5u h2d-done %6.2f start %6.2f k0-done %6.2f k12-done %6.2f k3-done %6.2f\n",
                        ci, d.chunks[ci].n, h, a, b, c2, e2);
            }
            if (to_host) { float x = 0; cudaEventElapsedTime(&x, d.events[0], d.events[3]); fprintf(stderr, "[alacgpu] d2h-done %6.2f\n", x); }
        }
        cudaEventElapsedTime(&ms, d.events[0], d.events[1]);
        k0 = std::max(k0, s0); k1 = std::max(k1, s1); k2 = std::max(k2, s2); k3 = std::max(k3, s3);
        kall = std::max(kall, ms);
        if (to_host) { cudaEventElapsedTime(&ms, d.events[2], d.events[3]); d2h = std::max(d2h, ms); }
        if (stage) {
            cudaEventElapsedTime(&ms, d.events[0], d.events[kEvBase + (d.chunks.size() - 1) * kEvPerChunk + 5]);
            h2d = std::max(h2d, ms);
        }
    }
    cudaGetLastError();
    alacgpu_timing &tm = ctx->timing;
    if (index) tm.index_ms = k0;
    if (decode) { tm.entropy_ms = k1; tm.lpc_ms = k2; tm.stereo_ms = k3; }
    tm.kernels_ms = kall;
    if (stage) tm.h2d_ms = h2d;
    if (ctx->tp_d2h) tm.d2h_ms = to_host ? d2h : 0.f;
}

struct PipeArgs {
    bool stage, index, decode;
    uint8_t *pcm_dst;       // caller's host buffer (or null)
    uint8_t *zc;            // device alias of pcm_dst when the pack roles write into it directly
    bool dst_pinned;
    uint32_t cf_override;   // 0 = automatic
};

struct DevRun {             // what one device's pipeline run reports back
    int32_t rc = ALACGPU_OK;
    std::string err;
    uint32_t launches = 0, chunks = 0, faults = 0;
    std::mutex m;
    void set_error(int32_t code, const char *what, cudaError_t e)
    {
        std::lock_guard<std::mutex> g(m);
        if (rc != ALACGPU_OK) return;
        rc = code;
        err = what;
        if (e != cudaSuccess) { err += ": "; err += cudaGetErrorString(e); }
    }
};

// CUDA call inside a per-device worker: record the first error and leave
#define CUD(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            res.set_error(e_ == cudaErrorMemoryAllocation ? ALACGPU_ERR_OUT_OF_MEMORY : ALACGPU_ERR_CUDA, #call, e_); \
            return false;                                                                                \
        }                                                                                                \
    } while (0)

uint64_t frame_lane_min_frames()
{
    const char *e = getenv("ALACGPU_KF_MIN");          // read per call: tests and tuning runs move the threshold
    return e ? strtoull(e, nullptr, 10) : kFrameLaneMinFrames;
}

bool frame_lanes_for(const alacgpu_ctx *ctx, const Device &d)
{
    if (ctx->opts.flags & (ALACGPU_FLAG_NO_FRAME_LANES | ALACGPU_FLAG_NO_FUSION)) return false;
    if (ctx->opts.flags & ALACGPU_FLAG_FORCE_FRAME_LANES) return true;
    return d.f_hi - d.f_lo >= frame_lane_min_frames();
}

// chunk size: as much as possible in flight at once when the bytes are already resident; ~kSlots chunks when
// streaming from the host so copies and kernels overlap
// bytes the channel-A planes of the frame-lane path may take on this device: the budget, or what is left of
// HBM (planes already allocated for earlier calls count as available)
uint64_t plane_budget(const Device &d)
{
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return kFrameLanePlaneBudget; }
    uint64_t have = 0;
    for (const Slot &s : d.slots) have += (uint64_t)s.planes.cap * sizeof(int32_t);
    const uint64_t avail = (uint64_t)free_b + have;
    const uint64_t keep = 6ull << 30;                       // work lists, status words, the caller's own allocations
    return std::min<uint64_t>(kFrameLanePlaneBudget, avail > keep ? avail - keep : 0);
}

uint32_t chunk_frames_for(const alacgpu_ctx *ctx, const Device &d, bool stage)
{
    const uint64_t n_local = d.f_hi - d.f_lo;
    const bool kf = frame_lanes_for(ctx, d);
    const uint32_t cap = kf ? kFrameLaneMaxChunk : kMaxChunkFrames;
    uint32_t cf = ctx->opts.chunk_frames;
    if (!cf && kf && d.cf_cache[stage ? 1 : 0]) return d.cf_cache[stage ? 1 : 0];
    if (!cf) {
        if (stage) {
            cf = (uint32_t)std::max<uint64_t>(kf ? 32768 : 256, (n_local + kSlots - 1) / kSlots);
            if (kf) cf = (uint32_t)std::min<uint64_t>(cf, std::max<uint64_t>(32768, plane_budget(d) / 2 / d.kf_row));
        } else if (kf) {
            // the whole shard in one chunk if its plane fits, else as few chunks as two alternating slots allow
            const uint64_t budget = std::max<uint64_t>(plane_budget(d), 2ull * 32768 * d.kf_row);
            uint64_t chunks = 1;
            if (n_local * d.kf_row > budget) chunks = (n_local * d.kf_row + budget / 2 - 1) / (budget / 2);
            cf = (uint32_t)std::min<uint64_t>((n_local + chunks - 1) / chunks, cap);
        } else {
            cf = (uint32_t)std::min<uint64_t>(n_local, cap);
        }
        cf = std::min<uint32_t>((cf + 31u) & ~31u, cap);
        if (kf) d.cf_cache[stage ? 1 : 0] = cf;
        return cf;
    }
    return std::min<uint32_t>((cf + 31u) & ~31u, cap);
}

// One device's share of a pipeline run: issue every chunk (H2D of its mdat -> kernels -> D2H of its PCM) and
// wait for it.  Runs on the caller's thread for a single-device context and on one thread per device
// otherwise.  When the caller's memory is pageable, two helper threads keep the three stages independent:
//   stager : caller bytes -> page-locked ring (copy pool) -> HBM, records "chunk c is on its way" events
//   issuer : (this thread) kernels of chunk c once the stager has recorded c's event
//   drainer: HBM -> page-locked ring -> caller's buffer (copy pool), chunk by chunk behind the kernels
bool run_device(alacgpu_ctx *ctx, Device &d, const PipeArgs &pa, DevRun &res)
{
    const uint64_t n_local = d.f_hi - d.f_lo;
    if (!n_local) return true;
    CUD(cudaSetDevice(d.id));
    const bool stage = pa.stage, index = pa.index, decode = pa.decode;
    uint8_t *const pcm_dst = pa.pcm_dst;
    uint8_t *const zc = pa.zc;
    d.frame_lanes = frame_lanes_for(ctx, d);
    static const bool no_taper = getenv("ALACGPU_NO_TAPER") != nullptr;
    uint32_t cf = chunk_frames_for(ctx, d, stage);
    size_t n_chunks = 0;
    // Chunks, slots and their buffers.  The frame-lane planes are sized by what cudaMemGetInfo reports as free; if
    // the allocation fails all the same (fragmentation, another tenant of the GPU), halve the chunk and try again
    // rather than fail a decode that fits.
    for (int attempt = 0;; attempt++) {
        build_chunks(ctx, d, cf, stage && !ctx->opts.chunk_frames && !no_taper && !d.frame_lanes);
        n_chunks = d.chunks.size();
        // slots in flight: a frame-lane chunk fills the machine on its own, and its channel-A plane is big
        d.slots_n = kSlots;
        const size_t kf_plane_elems = ((size_t)cf * d.kf_row + 3u) / 4u + 64u;
        if (d.frame_lanes && d.slots_cache[stage ? 1 : 0] && !ctx->opts.chunk_frames) {
            d.slots_n = d.slots_cache[stage ? 1 : 0];
        } else if (d.frame_lanes) {
            // As many slots as HBM has room for: two when the inputs are resident (the persistent kernels of
            // consecutive chunks run one after the other anyway), four while chunks stream in.  A slot whose plane is
            // already big enough costs nothing; growing one frees its old buffer first.
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
            int64_t room = (int64_t)free_b - (int64_t)(2ull << 30);
            int fit = 0;
            for (int s = 0; s < (stage ? 4 : 2); s++) {
                const size_t have = d.slots[s].planes.cap;
                if (have < kf_plane_elems) {
                    const int64_t cost = (int64_t)(kf_plane_elems - have) * (int64_t)sizeof(int32_t);
                    if (room < cost) break;
                    room -= cost;
                }
                fit++;
            }
            d.slots_n = std::max(1, fit);
            d.slots_cache[stage ? 1 : 0] = d.slots_n;
        }
        const int slots_used = (int)std::min<size_t>((size_t)d.slots_n, n_chunks);
        cudaError_t oom = cudaSuccess;
        if (decode)
            for (int s = 0; s < slots_used && oom == cudaSuccess; s++) {
                if (d.frame_lanes) {
                    oom = d.slots[s].planes.reserve_exact(kf_plane_elems);          // one row per frame (channel A)
                    if (oom == cudaSuccess) CUD(d.slots[s].kf.reserve(kf_list_words(cf) + kKfCountWords + cf));
                } else {
                    CUD(d.slots[s].planes.reserve((size_t)cf * 2u * d.ns));
                    CUD(d.slots[s].perm.reserve((size_t)cf * 4u + 1024u + 4u));
                    CUD(d.slots[s].progress.reserve((size_t)cf * 4u + 8u + ((size_t)cf * 2u + 3u) / 4u));
                }
            }
        if (oom == cudaSuccess) break;
        cudaGetLastError();
        if (oom != cudaErrorMemoryAllocation || attempt >= 4 || cf <= 65536 || ctx->opts.chunk_frames) CUD(oom);
        cf = ((cf / 2u) + 31u) & ~31u;
        d.cf_cache[stage ? 1 : 0] = cf;
        d.slots_cache[stage ? 1 : 0] = 0;
    }
    const int slots_used = (int)std::min<size_t>((size_t)d.slots_n, n_chunks);
    const bool to_host = pcm_dst && !zc && decode;
    bool ring_in = false;
    if (stage)
        for (const HostCopy &hc : d.copies) ring_in = ring_in || !hc.pinned;
    const bool ring_out = to_host && !pa.dst_pinned;
    static const bool no_ring = getenv("ALACGPU_NO_STAGING_RING") != nullptr;   // A/B: plain cudaMemcpyAsync on pageable memory
    if (no_ring) ring_in = false;
    const bool threaded = ring_in || (ring_out && !no_ring);
    if (threaded) {
        if (!ctx->pool) {
            const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
            static const int env_threads = getenv("ALACGPU_COPY_THREADS") ? atoi(getenv("ALACGPU_COPY_THREADS")) : 0;
            std::lock_guard<std::mutex> g(ctx->pool_mutex);
            if (!ctx->pool) ctx->pool.reset(new CopyPool(env_threads > 0 ? env_threads : (int)std::min(16u, std::max(2u, hw - 2))));   // measured on a 16-core box: 11.6 / 13.8 / 16.4 Gsamples/s end to end with 6 / 8 / 14 threads
        }
        if (ring_in) CUD(d.ring_in.ensure(kRingSlotBytes));
        if (ring_out) CUD(d.ring_out.ensure(kRingSlotBytes));
    }
    get_event(d, kEvBase + n_chunks * kEvPerChunk);          // create every event up front
    CUD(cudaEventRecord(d.events[0], d.slots[0].st));
    for (int s = 1; s < slots_used; s++) CUD(cudaStreamWaitEvent(d.slots[s].st, d.events[0], 0));
    if (to_host) {
        CUD(cudaStreamWaitEvent(d.st_d2h, d.events[0], 0));
        CUD(cudaEventRecord(d.events[2], d.st_d2h));
    }

    // ---- the three stages of one chunk ------------------------------------------------------------
    auto stage_chunk = [&](size_t ci) -> bool {          // H2D of the chunk's mdat bytes, then its "bytes are coming" event
        const Chunk &c = d.chunks[ci];
        for (uint32_t k = c.copy_lo; k < c.copy_hi; k++) {
            const HostCopy &hc = d.copies[k];
            if (hc.pinned || !ring_in) {
                CUD(cudaMemcpyAsync(d.arena.p + hc.dst, hc.src, hc.len, cudaMemcpyHostToDevice, d.st_h2d));
                continue;
            }
            PinnedRing &r = d.ring_in;
            for (uint64_t off = 0; off < hc.len; off += r.slot_bytes) {
                const uint64_t n = std::min<uint64_t>(r.slot_bytes, hc.len - off);
                const int j = r.next;
                if (r.busy[j]) CUD(cudaEventSynchronize(r.ev[j]));
                ctx->pool->copy(r.buf[j], hc.src + off, n);
                CUD(cudaMemcpyAsync(d.arena.p + hc.dst + off, r.buf[j], n, cudaMemcpyHostToDevice, d.st_h2d));
                CUD(cudaEventRecord(r.ev[j], d.st_h2d));
                r.busy[j] = true;
                r.next = (j + 1) % PinnedRing::kSlots;
            }
        }
        CUD(cudaEventRecord(d.events[kEvBase + ci * kEvPerChunk + 5], d.st_h2d));
        return true;
    };
    auto launch_chunk = [&](size_t ci) -> bool {         // the chunk's kernels on its slot stream
        const Chunk &c = d.chunks[ci];
        Slot &s = d.slots[ci % (size_t)d.slots_n];
        const size_t ev = kEvBase + ci * kEvPerChunk;
        if (stage) CUD(cudaStreamWaitEvent(s.st, d.events[ev + 5], 0));
        std::string err;
        uint32_t l = 0;
        const int32_t r = issue_chunk(ctx, d, c, s, index, decode, ev, &l, zc, stage, stage && ci < 3, &err);
        res.launches += l;
        res.chunks++;
        if (r) { res.set_error(r, err.c_str(), cudaSuccess); return false; }
        return true;
    };
    struct Flying { int slot; uint8_t *dst; uint64_t len; };
    std::deque<Flying> flying;                            // D2H pieces in the ring, oldest first
    auto land_one = [&]() -> bool {                       // oldest piece: wait for its DMA, copy it out to the caller
        const Flying f = flying.front();
        flying.pop_front();
        CUD(cudaEventSynchronize(d.ring_out.ev[f.slot]));
        ctx->pool->copy(f.dst, d.ring_out.buf[f.slot], f.len, /*streaming=*/true);
        d.ring_out.busy[f.slot] = false;
        return true;
    };
    auto drain_chunk = [&](size_t ci) -> bool {           // D2H of the chunk's PCM behind its last kernel
        const Chunk &c = d.chunks[ci];
        if (!(to_host && c.pcm_hi > c.pcm_lo)) return true;
        const size_t ev = kEvBase + ci * kEvPerChunk;
        CUD(cudaStreamWaitEvent(d.st_d2h, d.events[ev + 4], 0));
        if (!ring_out || no_ring) {
            CUD(cudaMemcpyAsync(pcm_dst + c.pcm_lo, d.pcm.p + (c.pcm_lo - d.pcm_lo), c.pcm_hi - c.pcm_lo,
                                cudaMemcpyDeviceToHost, d.st_d2h));
            return true;
        }
        PinnedRing &r = d.ring_out;
        for (uint64_t off = c.pcm_lo; off < c.pcm_hi; off += r.slot_bytes) {
            const uint64_t n = std::min<uint64_t>(r.slot_bytes, c.pcm_hi - off);
            const int j = r.next;
            while (r.busy[j]) if (!land_one()) return false;
            CUD(cudaMemcpyAsync(r.buf[j], d.pcm.p + (off - d.pcm_lo), n, cudaMemcpyDeviceToHost, d.st_d2h));
            CUD(cudaEventRecord(r.ev[j], d.st_d2h));
            r.busy[j] = true;
            r.next = (j + 1) % PinnedRing::kSlots;
            flying.push_back({j, pcm_dst + off, n});
        }
        return true;
    };

    bool ok = true;
    if (!threaded) {
        for (size_t ci = 0; ci < n_chunks && ok; ci++) {
            if (stage) ok = stage_chunk(ci);
            ok = ok && launch_chunk(ci);
            ok = ok && drain_chunk(ci);
        }
    } else {
        d.h2d_prog->reset();
        d.k_prog->reset();
        std::thread stager, drainer;
        if (stage)
            stager = std::thread([&] {
                if (cudaSetDevice(d.id) != cudaSuccess) { d.h2d_prog->fail(); return; }
                for (size_t ci = 0; ci < n_chunks; ci++) {
                    if (!stage_chunk(ci)) { d.h2d_prog->fail(); return; }
                    d.h2d_prog->set(ci + 1);
                }
            });
        if (to_host)
            drainer = std::thread([&] {
                bool good = cudaSetDevice(d.id) == cudaSuccess;
                for (size_t ci = 0; ci < n_chunks && good; ci++) {
                    if (!d.k_prog->wait_above(ci)) { good = false; break; }
                    good = drain_chunk(ci);
                }
                while (good && !flying.empty()) good = land_one();
                if (!good) res.set_error(ALACGPU_ERR_CUDA, "PCM drain failed", cudaSuccess);
            });
        for (size_t ci = 0; ci < n_chunks && ok; ci++) {
            if (stage && !d.h2d_prog->wait_above(ci)) { ok = false; break; }
            ok = launch_chunk(ci);
            if (ok) d.k_prog->set(ci + 1);
        }
        if (!ok) d.k_prog->fail();
        if (stager.joinable()) stager.join();
        if (drainer.joinable()) drainer.join();
        for (int j = 0; j < PinnedRing::kSlots; j++) d.ring_in.busy[j] = false;    // everything is waited for below
        ok = ok && res.rc == ALACGPU_OK;
    }
    if (!ok) {
        cudaDeviceSynchronize();
        if (res.rc == ALACGPU_OK) res.set_error(ALACGPU_ERR_CUDA, "pipeline issue failed", cudaSuccess);
        return false;
    }
    // join: slot 0 waits for the last chunk of every other slot, then stamps the end
    for (size_t ci = n_chunks > (size_t)d.slots_n ? n_chunks - (size_t)d.slots_n : 0; ci < n_chunks; ci++)
        if (ci % (size_t)d.slots_n != 0) CUD(cudaStreamWaitEvent(d.slots[0].st, d.events[kEvBase + ci * kEvPerChunk + 4], 0));
    CUD(cudaEventRecord(d.events[1], d.slots[0].st));
    if (to_host) CUD(cudaEventRecord(d.events[3], d.st_d2h));
    // ---- wait ---------------------------------------------------------------------------------------
    CUD(cudaStreamSynchronize(d.slots[0].st));
    if (to_host) CUD(cudaStreamSynchronize(d.st_d2h));
    if (stage) CUD(cudaStreamSynchronize(d.st_h2d));
    if (stage || index) d.resident = true;
    if (stage) d.arena_staged = d.arena_used;
    if (decode) { d.decoded = true; d.pcm_resident = !zc; }     // zero-copy output leaves no PCM in HBM
    if (index || decode) {
        uint32_t sc[3] = {0, 0, 0};                                // [0] K0 size mismatches, [1] FS_INTERNAL frames, [2] bounds checks
        CUD(cudaMemcpy(sc, d.scalars.p, sizeof sc, cudaMemcpyDeviceToHost));
        if (sc[0]) { res.set_error(ALACGPU_ERR_STATE, "internal: host and device disagree on a frame's PCM size", cudaSuccess); return false; }
        if (sc[2]) {                                               // only the ALACGPU_CHECKED build ever sets these
            char msg[160];
            snprintf(msg, sizeof msg, "checked build: bounds assertion failed, mask 0x%x (bit 0 arena, 1 plane, 2 list, 3 progress, 4 pcm, 5 ring, 6 frame)", sc[2]);
            cudaMemset(reinterpret_cast<uint32_t *>(d.scalars.p) + 2, 0, sizeof(uint32_t));
            res.set_error(ALACGPU_ERR_STATE, msg, cudaSuccess);
            return false;
        }
        if (sc[1]) {
            res.faults = sc[1];
            CUD(cudaMemset(reinterpret_cast<uint32_t *>(d.scalars.p) + 1, 0, sizeof(uint32_t)));
        }
    }
    return true;
}

// The whole pipeline on every device.  stage: copy mdat to the arena (chunk by chunk); index: run K0;
// decode: run the decode kernels; pcm_dst: copy PCM back chunk by chunk.
int32_t run_pipeline(alacgpu_ctx *ctx, bool stage, bool index, bool decode, uint8_t *pcm_dst, uint32_t *launches_out)
{
    finish_timing(ctx);          // the events are about to be re-recorded
    const double t_enter = now_ms();
    const int n_dev = (int)ctx->devs.size();
    // Zero-copy output: when the caller's buffer is page-locked and mapped (alacgpu_host_alloc, or any
    // cudaHostAlloc / cudaHostRegister'ed range) and the pack stage is fused, the pack warps write the
    // PCM straight into it over PCIe while the frames are still being decoded; there is no device PCM
    // copy and no D2H stage.
    bool all_fully_fused = !stage;
    for (const Device &d : ctx->devs)
        if (d.f_hi > d.f_lo && (chunk_frames_for(ctx, d, stage) > kFullFusionMaxFrames || frame_lanes_for(ctx, d))) all_fully_fused = false;
    PipeArgs pa{stage, index, decode, pcm_dst, nullptr, false, 0};
    if (pcm_dst) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, pcm_dst) == cudaSuccess && at.type == cudaMemoryTypeHost) {
            pa.dst_pinned = true;
            if (decode && all_fully_fused && at.devicePointer &&
                !(ctx->opts.flags & (ALACGPU_FLAG_NO_FUSION | ALACGPU_FLAG_NO_PACK_FUSION | ALACGPU_FLAG_NO_ZERO_COPY)))
                pa.zc = static_cast<uint8_t *>(at.devicePointer);
        } else {
            cudaGetLastError();
        }
    }
    if (pcm_dst && decode) {       // alignment gaps between tracks are zero bytes in the layout (copies skip them)
        uint64_t end = 0;
        for (const HostTrack &ht : ctx->tracks) {
            if (ht.pcm_off > end) memset(pcm_dst + end, 0, ht.pcm_off - end);
            end = ht.pcm_off + ht.pcm_len;
        }
    }
    // one issuing thread per device (a single device runs on the caller's thread)
    std::vector<DevRun> runs(n_dev);
    if (n_dev == 1) {
        run_device(ctx, ctx->devs[0], pa, runs[0]);
    } else {
        std::vector<std::thread> th;
        for (int g = 0; g < n_dev; g++) th.emplace_back([&, g] { run_device(ctx, ctx->devs[g], pa, runs[g]); });
        for (std::thread &t : th) t.join();
    }
    uint32_t launches = 0, chunks_total = 0, faults = 0;
    for (int g = 0; g < n_dev; g++) {
        if (runs[g].rc != ALACGPU_OK) return fail(ctx, runs[g].rc, runs[g].err.c_str());
        launches += runs[g].launches;
        chunks_total += runs[g].chunks;
        faults += runs[g].faults;
    }
    ctx->internal_faults = faults;
    if (stage)                        // every byte is in HBM now: the host memory is no longer borrowed
        for (HostTrack &ht : ctx->tracks) { ht.staged = true; ht.mdat = nullptr; }
    if (getenv("ALACGPU_HOST_TIMING"))
        fprintf(stderr, "[alacgpu] run_pipeline: %.3f ms\n", now_ms() - t_enter);
    ctx->timing_pending = true;
    ctx->tp_stage = stage; ctx->tp_index = index; ctx->tp_decode = decode; ctx->tp_d2h = pcm_dst != nullptr; ctx->tp_zc = pa.zc != nullptr;
    ctx->timing.chunks = chunks_total;
    if (launches_out) *launches_out = launches;
    ctx->have_status = false;
    return ALACGPU_OK;
}

void fill_totals(alacgpu_ctx *ctx)
{
    uint64_t samples = 0, pcm = 0;
    for (const HostTrack &ht : ctx->tracks) {
        samples += ht.pcm_len / (uint64_t)(ht.cfg.sample_size / 8);
        pcm += ht.pcm_len;
    }
    ctx->timing.compressed_bytes = ctx->compressed_bytes;
    ctx->timing.pcm_bytes = pcm;
    ctx->timing.samples = samples;
}

bool all_resident(alacgpu_ctx *ctx)
{
    if (!ctx->planned) return false;
    for (Device &d : ctx->devs) if (d.f_hi > d.f_lo && !d.resident) return false;
    return true;
}

bool all_decoded(alacgpu_ctx *ctx)
{
    if (!ctx->planned) return false;
    for (Device &d : ctx->devs) if (d.f_hi > d.f_lo && !d.decoded) return false;
    return true;
}

bool all_pcm_resident(alacgpu_ctx *ctx)
{
    if (!ctx->planned) return false;
    for (Device &d : ctx->devs) if (d.f_hi > d.f_lo && !d.pcm_resident) return false;
    return true;
}

}  // namespace

// ---------------------------------------------------------------------------
extern "C" {

const char *alacgpu_strerror(int32_t s)
{
    switch (s) {
    case ALACGPU_OK: return "ok";
    case ALACGPU_ERR_INVALID_ARG: return "invalid argument";
    case ALACGPU_ERR_NO_DEVICE: return "no usable CUDA device (libalacgpu has no CPU fallback)";
    case ALACGPU_ERR_CUDA: return "CUDA runtime error";
    case ALACGPU_ERR_OUT_OF_MEMORY: return "out of device or pinned memory";
    case ALACGPU_ERR_UNSUPPORTED: return "unsupported stream parameters (sample size must be 16 or 24, channels 1 or 2)";
    case ALACGPU_ERR_CAPACITY: return "destination buffer too small";
    case ALACGPU_ERR_STATE: return "call made in the wrong state";
    case ALACGPU_ERR_RANGE: return "track or frame index out of range";
    default: return "unknown status";
    }
}

const char *alacgpu_last_error(alacgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }
int32_t alacgpu_abi_version(void) { return ALACGPU_ABI_VERSION; }

int32_t alacgpu_device_count(int32_t *n)
{
    if (!n) return ALACGPU_ERR_INVALID_ARG;
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); c = 0; }
    *n = c;
    return ALACGPU_OK;
}

int32_t alacgpu_create(const int32_t *device_ids, int32_t n_devices, const alacgpu_opts *opts, alacgpu_ctx **out)
{
    if (!out) return ALACGPU_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); return ALACGPU_ERR_NO_DEVICE; }
    alacgpu_ctx *ctx = new alacgpu_ctx();
    if (opts) {
        size_t n = std::min<size_t>(opts->struct_size ? opts->struct_size : sizeof(alacgpu_opts), sizeof(alacgpu_opts));
        memcpy(&ctx->opts, opts, n);
    }
    std::vector<int> ids;
    if (!device_ids || n_devices <= 0) ids.push_back(0);
    else ids.assign(device_ids, device_ids + n_devices);
    for (int id : ids) {
        if (id < 0 || id >= count) { delete ctx; return ALACGPU_ERR_NO_DEVICE; }
        ctx->devs.emplace_back();
        ctx->devs.back().id = id;
        ctx->devs.back().h2d_prog.reset(new Progress());
        ctx->devs.back().k_prog.reset(new Progress());
    }
    for (Device &d : ctx->devs) {
        bool ok = cudaSetDevice(d.id) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&d.st_h2d, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&d.st_d2h, cudaStreamNonBlocking) == cudaSuccess;
        for (int s = 0; ok && s < kSlots; s++)
            ok = cudaStreamCreateWithFlags(&d.slots[s].st, cudaStreamNonBlocking) == cudaSuccess;
        if (!ok) {
            cudaGetLastError();
            alacgpu_destroy(ctx);
            return ALACGPU_ERR_NO_DEVICE;
        }
    }
    *out = ctx;
    return ALACGPU_OK;
}

int32_t alacgpu_destroy(alacgpu_ctx *ctx)
{
    if (!ctx) return ALACGPU_OK;
    for (Device &d : ctx->devs) {
        cudaSetDevice(d.id);
        cudaDeviceSynchronize();
        d.arena.release(); d.refs.release(); d.cfgs.release(); d.desc.release(); d.coefs.release();
        d.expect_len.release(); d.frame_off.release(); d.scalars.release(); d.pcm.release();
        for (Slot &s : d.slots) { s.planes.release(); s.perm.release(); s.progress.release(); s.kf.release(); if (s.st) cudaStreamDestroy(s.st); }
        for (cudaEvent_t e : d.events) cudaEventDestroy(e);
        d.ring_in.release();
        d.ring_out.release();
        if (d.st_h2d) cudaStreamDestroy(d.st_h2d);
        if (d.st_d2h) cudaStreamDestroy(d.st_d2h);
    }
    if (ctx->window) cudaFreeHost(ctx->window);
    delete ctx;
    return ALACGPU_OK;
}

static int32_t add_track_impl(alacgpu_ctx *ctx, const alacgpu_track_cfg *cfg, const uint8_t *mdat, uint64_t mdat_len,
                              uint64_t first_frame_offset, const uint64_t *frame_offsets, const uint32_t *frame_sizes,
                              uint32_t n_frames, int32_t *track_id)
{
    if (!ctx || !cfg || (!mdat && mdat_len) || (!frame_sizes && n_frames)) return ALACGPU_ERR_INVALID_ARG;
    // "FIXME: unimplemented sample size" (AlacFile.cs:570-574, :713-715)
    if (cfg->sample_size != 16 && cfg->sample_size != 24) return fail(ctx, ALACGPU_ERR_UNSUPPORTED, "sample size must be 16 or 24");
    if (cfg->num_channels != 1 && cfg->num_channels != 2) return fail(ctx, ALACGPU_ERR_UNSUPPORTED, "container channels must be 1 or 2");
    const int64_t bpsf = (int64_t)(cfg->sample_size / 8) * cfg->num_channels;
    if (cfg->max_samples_per_frame < 1 || cfg->max_samples_per_frame > kMaxFrameSamples ||
        cfg->max_samples_per_frame * bpsf > kMaxFramePcmBytes)
        return fail(ctx, ALACGPU_ERR_UNSUPPORTED, "max_samples_per_frame outside the reference's buffers (AlacFile.cs:28, AlacContext.cs:218)");
    if (cfg->rice_kmodifier < 0 || cfg->rice_kmodifier > 31 || cfg->rice_history_mult < 0 || cfg->rice_history_mult > 255 ||
        cfg->rice_initial_history < 0 || cfg->rice_initial_history > 255)
        return fail(ctx, ALACGPU_ERR_UNSUPPORTED, "rice parameters outside one cookie byte / kmodifier 0..31");
    if (ctx->devs.size() > 1)
        for (const HostTrack &ht : ctx->tracks)
            if (ht.staged)
                return fail(ctx, ALACGPU_ERR_STATE, "a multi-device context cannot take more tracks once its tracks are staged "
                                                    "(the partition would move frames whose host bytes are no longer borrowed): "
                                                    "call alacgpu_clear_tracks first");
    HostTrack t{};
    t.cfg = *cfg;
    if (cfg->num_channels != 1) ctx->mono_only = false;
    t.mdat = mdat;
    t.mdat_len = mdat_len;
    t.first_frame_offset = first_frame_offset;
    t.first_frame = ctx->sizes.size();
    t.n_frames = n_frames;
    // output layout: every frame's PCM size follows from its first seven bytes
    t.pcm_off = align_up(ctx->total_pcm, kTrackAlign);
    uint64_t off = first_frame_offset, pcm = t.pcm_off;
    ctx->sizes.insert(ctx->sizes.end(), frame_sizes, frame_sizes + n_frames);
    ctx->out_len.reserve(ctx->out_len.size() + n_frames);
    ctx->frame_off.reserve(ctx->frame_off.size() + n_frames);
    if (frame_offsets) t.offs.assign(frame_offsets, frame_offsets + n_frames);
    for (uint32_t f = 0; f < n_frames; f++) {
        if (frame_offsets) off = frame_offsets[f];
        const uint64_t avail = off < mdat_len ? mdat_len - off : 0;
        const uint64_t len = std::min<uint64_t>(frame_sizes[f], avail);
        const uint32_t bytes = frame_pcm_bytes(*cfg, mdat + std::min(off, mdat_len), len);
        ctx->out_len.push_back(bytes);
        ctx->frame_off.push_back(pcm);
        ctx->max_sf = std::max<uint32_t>(ctx->max_sf, bytes / (uint32_t)bpsf);
        pcm += bytes;
        off += frame_sizes[f];
    }
    t.pcm_len = pcm - t.pcm_off;
    ctx->total_pcm = pcm;
    ctx->tracks.push_back(t);
    invalidate(ctx);
    if (track_id) *track_id = (int32_t)ctx->tracks.size() - 1;
    return ALACGPU_OK;
}

int32_t alacgpu_add_track(alacgpu_ctx *ctx, const alacgpu_track_cfg *cfg, const uint8_t *mdat, uint64_t mdat_len,
                          uint64_t first_frame_offset, const uint32_t *frame_sizes, uint32_t n_frames,
                          int32_t *track_id)
{
    return add_track_impl(ctx, cfg, mdat, mdat_len, first_frame_offset, nullptr, frame_sizes, n_frames, track_id);
}

int32_t alacgpu_add_track_offsets(alacgpu_ctx *ctx, const alacgpu_track_cfg *cfg, const uint8_t *file, uint64_t file_len,
                                  const uint64_t *frame_offsets, const uint32_t *frame_sizes, uint32_t n_frames,
                                  int32_t *track_id)
{
    if (!frame_offsets && n_frames) return ALACGPU_ERR_INVALID_ARG;
    return add_track_impl(ctx, cfg, file, file_len, 0, frame_offsets, frame_sizes, n_frames, track_id);
}

int32_t alacgpu_clear_tracks(alacgpu_ctx *ctx)
{
    if (!ctx) return ALACGPU_ERR_INVALID_ARG;
    ctx->tracks.clear();
    ctx->sizes.clear();
    ctx->out_len.clear();
    ctx->frame_off.clear();
    ctx->total_pcm = 0;
    ctx->max_sf = 0;
    ctx->mono_only = true;
    invalidate(ctx);
    for (Device &d : ctx->devs) { d.f_lo = d.f_hi = 0; d.arena_staged = 0; }
    return ALACGPU_OK;
}

int32_t alacgpu_plan_partition(const uint32_t *frame_sizes, uint64_t n_frames, int32_t n_parts, uint64_t *cut)
{
    if (n_parts <= 0 || !cut || (!frame_sizes && n_frames)) return ALACGPU_ERR_INVALID_ARG;
    uint64_t total = 0;
    for (uint64_t i = 0; i < n_frames; i++) total += frame_sizes[i];
    cut[0] = 0;
    uint64_t acc = 0, f = 0;
    for (int32_t p = 1; p < n_parts; p++) {
        // cut after the frame whose midpoint crosses total * p / n_parts
        const long double target = (long double)total * p / n_parts;
        while (f < n_frames && (long double)acc + frame_sizes[f] / 2.0L <= target) acc += frame_sizes[f++];
        cut[p] = f;
    }
    cut[n_parts] = n_frames;
    return ALACGPU_OK;
}

int32_t alacgpu_total_pcm_bytes(alacgpu_ctx *ctx, uint64_t *total_pcm_bytes)
{
    if (!ctx || !total_pcm_bytes) return ALACGPU_ERR_INVALID_ARG;
    *total_pcm_bytes = ctx->total_pcm;
    return ALACGPU_OK;
}

int32_t alacgpu_prepare(alacgpu_ctx *ctx, uint64_t *total_pcm_bytes)
{
    if (!ctx) return ALACGPU_ERR_INVALID_ARG;
    if (!all_resident(ctx)) {
        const double t0 = now_ms();
        int32_t r = build_plan(ctx);
        if (r) return r;
        ctx->timing = alacgpu_timing{};
        r = run_pipeline(ctx, /*stage=*/true, /*index=*/true, /*decode=*/false, nullptr, &ctx->index_launches);
        if (r) return r;
        fill_totals(ctx);
        ctx->timing.kernel_launches = ctx->index_launches;
        ctx->timing.total_ms = (float)(now_ms() - t0);
    }
    if (total_pcm_bytes) *total_pcm_bytes = ctx->total_pcm;
    return ALACGPU_OK;
}

int32_t alacgpu_reindex(alacgpu_ctx *ctx)
{
    if (!ctx) return ALACGPU_ERR_INVALID_ARG;
    int32_t r = alacgpu_prepare(ctx, nullptr);
    if (r) return r;
    // lazy: the next decode_all runs K0 over the resident bytes inside its own pipeline (one launch
    // sequence, one synchronisation per step instead of three)
    ctx->index_stale = true;
    for (Device &d : ctx->devs) { d.decoded = false; d.pcm_resident = false; }
    return ALACGPU_OK;
}

int32_t alacgpu_decode_all(alacgpu_ctx *ctx, uint8_t *pcm_dst, uint64_t cap, uint64_t *track_pcm_off,
                           uint64_t *track_pcm_len, int32_t *frame_status)
{
    if (!ctx) return ALACGPU_ERR_INVALID_ARG;
    if (pcm_dst && cap < ctx->total_pcm) return fail(ctx, ALACGPU_ERR_CAPACITY, "pcm_dst smaller than the total PCM size");
    const double t0 = now_ms();
    int32_t r = build_plan(ctx);
    if (r) return r;
    if (getenv("ALACGPU_HOST_TIMING")) fprintf(stderr, "[alacgpu] build_plan %.3f ms\n", now_ms() - t0);
    const bool resident = all_resident(ctx);
    const bool reindex = !resident || ctx->index_stale;
    if (reindex) { finish_timing(ctx); ctx->timing = alacgpu_timing{}; ctx->index_launches = 0; }
    // not yet staged: stream mdat in, index and decode chunk by chunk; else decode only (after
    // alacgpu_reindex: header pre-pass + decode over the resident bytes)
    uint32_t launches = 0;
    r = run_pipeline(ctx, !resident, reindex, true, pcm_dst, &launches);
    if (r) return r;
    uint32_t retries = 0;
    static const bool inject = getenv("ALACGPU_TEST_INJECT_INTERNAL") != nullptr;   // tests: exercise the retry below
    if (inject) ctx->internal_faults = 1;
    if (ctx->internal_faults) {
        // A consumer of a fused launch gave up waiting for its producer (never expected: only a GPU that is
        // time-sliced or single-stepped can starve a resident block that long).  The frames it flagged hold
        // zeros; decode the batch again with three plain kernels that do not wait on each other, so the caller
        // never gets a silently wrong frame.
        const uint32_t saved = ctx->opts.flags;
        ctx->opts.flags |= ALACGPU_FLAG_NO_FUSION | ALACGPU_FLAG_NO_FRAME_LANES;
        uint32_t more = 0;
        r = run_pipeline(ctx, false, true, true, pcm_dst, &more);
        ctx->opts.flags = saved;
        if (r) return r;
        launches += more;
        retries = 1;
        if (ctx->internal_faults && !inject) return fail(ctx, ALACGPU_ERR_STATE, "internal: frames still flagged after the unfused retry");
    }
    ctx->index_stale = false;
    fill_totals(ctx);
    ctx->timing.kernel_launches = ctx->index_launches + launches;
    ctx->timing.internal_retries = retries;
    ctx->timing.total_ms = (float)(now_ms() - t0);
    if (getenv("ALACGPU_HOST_TIMING")) fprintf(stderr, "[alacgpu] decode_all %.3f ms\n", now_ms() - t0);
    for (size_t t = 0; t < ctx->tracks.size(); t++) {
        if (track_pcm_off) track_pcm_off[t] = ctx->tracks[t].pcm_off;
        if (track_pcm_len) track_pcm_len[t] = ctx->tracks[t].pcm_len;
    }
    if (frame_status) {
        for (Device &d : ctx->devs) {
            const uint64_t n_local = d.f_hi - d.f_lo;
            if (!n_local) continue;
            CU(cudaSetDevice(d.id));
            std::vector<FrameDesc> h(n_local);
            CU(cudaMemcpy(h.data(), d.desc.p, n_local * sizeof(FrameDesc), cudaMemcpyDeviceToHost));
            for (uint64_t i = 0; i < n_local; i++) frame_status[d.f_lo + i] = h[i].status;
        }
    }
    return ALACGPU_OK;
}

// ---------------------------------------------------------------------------
static int32_t ensure_status(alacgpu_ctx *ctx)
{
    if (ctx->have_status) return ALACGPU_OK;
    ctx->h_status.assign(ctx->sizes.size(), 0);
    for (Device &d : ctx->devs) {
        const uint64_t n_local = d.f_hi - d.f_lo;
        if (!n_local) continue;
        CU(cudaSetDevice(d.id));
        std::vector<FrameDesc> h(n_local);
        CU(cudaMemcpy(h.data(), d.desc.p, n_local * sizeof(FrameDesc), cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < n_local; i++) ctx->h_status[d.f_lo + i] = h[i].status;
    }
    ctx->have_status = true;
    return ALACGPU_OK;
}

static int32_t check_track(alacgpu_ctx *ctx, int32_t track)
{
    if (!ctx) return ALACGPU_ERR_INVALID_ARG;
    if (track < 0 || (size_t)track >= ctx->tracks.size()) return fail(ctx, ALACGPU_ERR_RANGE, "track index out of range");
    return ALACGPU_OK;
}

int32_t alacgpu_read_frame(alacgpu_ctx *ctx, int32_t track, uint32_t frame_idx, uint8_t *dst, uint32_t cap, uint32_t *bytes_out)
{
    int32_t r = check_track(ctx, track);
    if (r) return r;
    if (!bytes_out) return ALACGPU_ERR_INVALID_ARG;
    *bytes_out = 0;
    const HostTrack &ht = ctx->tracks[track];
    if (frame_idx >= ht.n_frames) return ALACGPU_OK;          // AlacContext.cs:182-186: return 0
    if (!all_pcm_resident(ctx)) {
        r = alacgpu_decode_all(ctx, nullptr, 0, nullptr, nullptr, nullptr);
        if (r) return r;
    }
    const uint64_t gidx = ht.first_frame + frame_idx;
    const uint64_t off = ctx->frame_off[gidx];
    const uint32_t len = ctx->out_len[gidx];
    if (len == 0) return ALACGPU_OK;
    if (!dst || cap < len) return fail(ctx, ALACGPU_ERR_CAPACITY, "frame does not fit the destination");
    if (!(off >= ctx->win_lo && off + len <= ctx->win_hi)) {
        // refill the pinned window from the device that owns this frame
        if (!ctx->window) CU(cudaMallocHost(&ctx->window, kReadWindow));
        Device *own = nullptr;
        for (Device &d : ctx->devs) if (gidx >= d.f_lo && gidx < d.f_hi) own = &d;
        if (!own) return fail(ctx, ALACGPU_ERR_STATE, "frame has no owning device");
        CU(cudaSetDevice(own->id));
        const uint64_t hi = std::min<uint64_t>(own->pcm_hi, off + kReadWindow);
        CU(cudaMemcpy(ctx->window, own->pcm.p + (off - own->pcm_lo), hi - off, cudaMemcpyDeviceToHost));
        ctx->win_lo = off;
        ctx->win_hi = hi;
    }
    memcpy(dst, ctx->window + (off - ctx->win_lo), len);
    *bytes_out = len;
    return ALACGPU_OK;
}

int32_t alacgpu_track_count(alacgpu_ctx *ctx, int32_t *n)
{
    if (!ctx || !n) return ALACGPU_ERR_INVALID_ARG;
    *n = (int32_t)ctx->tracks.size();
    return ALACGPU_OK;
}

int32_t alacgpu_frame_count(alacgpu_ctx *ctx, int32_t track, uint32_t *n_frames)
{
    int32_t r = check_track(ctx, track);
    if (r) return r;
    if (!n_frames) return ALACGPU_ERR_INVALID_ARG;
    *n_frames = ctx->tracks[track].n_frames;
    return ALACGPU_OK;
}

int32_t alacgpu_frame_samples(alacgpu_ctx *ctx, int32_t track, uint32_t frame_idx, uint32_t *n_samples)
{
    int32_t r = check_track(ctx, track);
    if (r) return r;
    if (!n_samples) return ALACGPU_ERR_INVALID_ARG;
    const HostTrack &ht = ctx->tracks[track];
    if (frame_idx >= ht.n_frames) return fail(ctx, ALACGPU_ERR_RANGE, "frame index out of range");
    *n_samples = ctx->out_len[ht.first_frame + frame_idx] / (uint32_t)((ht.cfg.sample_size / 8) * ht.cfg.num_channels);
    return ALACGPU_OK;
}

int32_t alacgpu_frame_status(alacgpu_ctx *ctx, int32_t track, uint32_t frame_idx, int32_t *status)
{
    int32_t r = check_track(ctx, track);
    if (r) return r;
    if (!status) return ALACGPU_ERR_INVALID_ARG;
    const HostTrack &ht = ctx->tracks[track];
    if (frame_idx >= ht.n_frames) return fail(ctx, ALACGPU_ERR_RANGE, "frame index out of range");
    if (!all_decoded(ctx)) return fail(ctx, ALACGPU_ERR_STATE, "frames have not been decoded yet");
    r = ensure_status(ctx);
    if (r) return r;
    *status = ctx->h_status[ht.first_frame + frame_idx];
    return ALACGPU_OK;
}

int32_t alacgpu_track_pcm_bytes(alacgpu_ctx *ctx, int32_t track, uint64_t *off, uint64_t *len)
{
    int32_t r = check_track(ctx, track);
    if (r) return r;
    if (off) *off = ctx->tracks[track].pcm_off;
    if (len) *len = ctx->tracks[track].pcm_len;
    return ALACGPU_OK;
}

int32_t alacgpu_get_timing(alacgpu_ctx *ctx, alacgpu_timing *out)
{
    if (!ctx || !out) return ALACGPU_ERR_INVALID_ARG;
    finish_timing(ctx);
    *out = ctx->timing;
    return ALACGPU_OK;
}

int32_t alacgpu_device_pcm(alacgpu_ctx *ctx, int32_t dev_slot, void **dptr, uint64_t *shard_off, uint64_t *shard_len)
{
    if (!ctx || dev_slot < 0 || (size_t)dev_slot >= ctx->devs.size()) return ALACGPU_ERR_INVALID_ARG;
    Device &d = ctx->devs[dev_slot];
    if (!d.pcm_resident && d.f_hi > d.f_lo) return fail(ctx, ALACGPU_ERR_STATE, "no decoded PCM on this device");
    if (dptr) *dptr = d.pcm.p;
    if (shard_off) *shard_off = d.pcm_lo;
    if (shard_len) *shard_len = d.pcm_hi - d.pcm_lo;
    return ALACGPU_OK;
}

int32_t alacgpu_pcm_checksum(alacgpu_ctx *ctx, uint64_t off, uint64_t len, uint64_t *sum)
{
    if (!ctx || !sum || (off & 7)) return ALACGPU_ERR_INVALID_ARG;
    uint64_t total = 0;
    for (Device &d : ctx->devs) {
        if (d.f_hi == d.f_lo) continue;
        if (!d.pcm_resident) return fail(ctx, ALACGPU_ERR_STATE, "no decoded PCM on this device (decode with pcm_dst == NULL first)");
        // this shard's bytes inside [off, off+len); shards start on 256-byte boundaries of the
        // global layout except where a track is split between devices (then on a frame boundary)
        uint64_t lo = std::max(off, d.pcm_first), hi = std::min(off + len, d.pcm_hi);
        if (hi <= lo) continue;
        if (lo & 7) return fail(ctx, ALACGPU_ERR_INVALID_ARG, "checksum range must start on an 8-byte boundary of each shard");
        CU(cudaSetDevice(d.id));
        cudaStream_t st = d.slots[0].st;
        CU(cudaMemsetAsync(d.scalars.p + 2, 0, sizeof(uint64_t), st));
        CU(launch_checksum(d.pcm.p + (lo - d.pcm_lo), lo, hi - lo, d.scalars.p + 2, st));
        uint64_t part = 0;
        CU(cudaMemcpyAsync(&part, d.scalars.p + 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        total += part;
    }
    *sum = total;
    return ALACGPU_OK;
}

int32_t alacgpu_host_alloc(uint64_t bytes, void **ptr)
{
    if (!ptr) return ALACGPU_ERR_INVALID_ARG;
    *ptr = nullptr;
    cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorMemoryAllocation ? ALACGPU_ERR_OUT_OF_MEMORY : ALACGPU_ERR_NO_DEVICE; }
    return ALACGPU_OK;
}

int32_t alacgpu_host_free(void *ptr)
{
    if (ptr) cudaFreeHost(ptr);
    return ALACGPU_OK;
}

}  // extern "C"
