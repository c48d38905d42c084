// runtime.cu -- host runtime and C ABI of libalacgpu.so.
//
// Owns: the per-device mdat arena and frame index (the device-resident form of
// the demuxer's stsz table, ALACDecoder/DemuxResT.cs:28), the multi-GPU frame
// partition, the chunked K1->K2->K3 pipeline, PCM staging back to the caller
// and the per-frame pull that stands behind AlacContext.Read
// (ALACDecoder/AlacContext.cs:163-204).  No CPU decode path exists here: every
// sample is produced by the kernels.
#include "../../../include/alacgpu.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "alacgpu_device.cuh"
#include "alacgpu_kernels.h"

using namespace alacgpu;

namespace {

constexpr uint64_t kTrackAlign = 256;     // PCM start alignment of each track in the global layout
// zero padding after the last staged byte: a lane whose frame is truncated keeps reading
// (at most 2 channels x 16384 symbols x 59 bits) until K1 flags the overrun at the end
constexpr uint64_t kArenaTail = 256 * 1024 + 256;
constexpr uint32_t kDefaultChunkFrames = 32768;
constexpr uint64_t kReadWindow = 8ull << 20;   // host-side cache window of alacgpu_read_frame

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
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct HostTrack {
    alacgpu_track_cfg cfg;
    const uint8_t *mdat;          // borrowed until prepare()
    uint64_t mdat_len;
    uint64_t first_frame_offset;
    uint64_t first_frame;         // index into the global frame list
    uint32_t n_frames;
    uint64_t pcm_off, pcm_len;    // global (padded) layout, valid after prepare
};

struct Device {
    int id = 0;
    cudaStream_t st = nullptr;        // compute
    cudaStream_t st_copy = nullptr;   // H2D / D2H
    // shard = global frames [f_lo, f_hi)
    uint64_t f_lo = 0, f_hi = 0;
    DevBuf<uint8_t> arena;
    uint64_t arena_used = 0;
    DevBuf<FrameRef> refs;
    DevBuf<TrackCfg> cfgs;
    DevBuf<FrameDesc> desc;
    DevBuf<FrameCoefs> coefs;
    DevBuf<uint32_t> out_len;
    DevBuf<uint64_t> frame_off;
    DevBuf<uint64_t> block_sums;
    DevBuf<uint64_t> track_first;     // per track: first local frame index (or n_local)
    DevBuf<uint64_t> track_start;     // n_tracks + 1
    DevBuf<uint64_t> track_shift;     // per track
    DevBuf<uint64_t> scalars;         // [0] grand total, [1] (u32) max samples, [2] checksum
    DevBuf<int32_t> planes;
    DevBuf<uint8_t> pcm;
    DevBuf<int32_t> status32;
    uint64_t pcm_lo = 0, pcm_hi = 0;  // global byte range of this shard's PCM buffer (pcm_lo 256-aligned)
    uint64_t pcm_first = 0;           // global offset of the first byte this shard produces
    uint64_t total_unpadded = 0;
    uint32_t max_samples = 0;
    std::vector<uint64_t> h_track_start;
    std::vector<cudaEvent_t> events;
    bool decoded = false;
};

}  // namespace

struct alacgpu_ctx {
    std::vector<Device> devs;
    std::vector<HostTrack> tracks;
    std::vector<uint32_t> sizes;          // stsz of every frame, track-major
    alacgpu_opts opts{};
    bool prepared = false;
    uint64_t total_pcm = 0;
    uint64_t compressed_bytes = 0;
    alacgpu_timing timing{};
    std::string err;
    // read_frame support (host mirrors, filled lazily)
    std::vector<uint64_t> h_frame_off;    // global padded offset of every frame
    std::vector<uint32_t> h_frame_len;
    std::vector<int32_t> h_status;
    bool have_frame_tables = false;
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

// which track holds global frame g (tracks sorted by first_frame)
size_t track_of(const alacgpu_ctx *ctx, uint64_t g)
{
    size_t lo = 0, hi = ctx->tracks.size();
    while (hi - lo > 1) {
        size_t mid = (lo + hi) / 2;
        if (ctx->tracks[mid].first_frame <= g) lo = mid; else hi = mid;
    }
    // skip empty tracks that share the same first_frame
    while (lo + 1 < ctx->tracks.size() && ctx->tracks[lo].n_frames == 0) lo++;
    return lo;
}

int lanes_for(const alacgpu_ctx *ctx, uint32_t n_frames)
{
    const uint32_t o = ctx->opts.entropy_lanes;
    if (o == 4 || o == 8 || o == 16 || o == 32) return (int)o;
    // enough warps to cover 148 SMs x 4 schedulers a few times over, else
    // trade idle lanes for more resident warps (the stage is latency bound)
    if (n_frames / 32 >= 148u * 12u) return 32;
    if (n_frames / 16 >= 148u * 12u) return 16;
    return 8;
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
    if (ctx->opts.chunk_frames == 0) ctx->opts.chunk_frames = kDefaultChunkFrames;
    ctx->opts.chunk_frames = (ctx->opts.chunk_frames + 31u) & ~31u;
    std::vector<int> ids;
    if (!device_ids || n_devices <= 0) ids.push_back(0);
    else ids.assign(device_ids, device_ids + n_devices);
    for (int id : ids) {
        if (id < 0 || id >= count) { delete ctx; return ALACGPU_ERR_NO_DEVICE; }
        Device d;
        d.id = id;
        ctx->devs.push_back(std::move(d));
    }
    for (Device &d : ctx->devs) {
        if (cudaSetDevice(d.id) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.st_copy, cudaStreamNonBlocking) != cudaSuccess) {
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
        if (d.st) cudaStreamSynchronize(d.st);
        if (d.st_copy) cudaStreamSynchronize(d.st_copy);
        d.arena.release(); d.refs.release(); d.cfgs.release(); d.desc.release(); d.coefs.release();
        d.out_len.release(); d.frame_off.release(); d.block_sums.release(); d.track_first.release();
        d.track_start.release(); d.track_shift.release(); d.scalars.release(); d.planes.release();
        d.pcm.release(); d.status32.release();
        for (cudaEvent_t e : d.events) cudaEventDestroy(e);
        if (d.st) cudaStreamDestroy(d.st);
        if (d.st_copy) cudaStreamDestroy(d.st_copy);
    }
    if (ctx->window) cudaFreeHost(ctx->window);
    delete ctx;
    return ALACGPU_OK;
}

int32_t alacgpu_add_track(alacgpu_ctx *ctx, const alacgpu_track_cfg *cfg, const uint8_t *mdat, uint64_t mdat_len,
                          uint64_t first_frame_offset, const uint32_t *frame_sizes, uint32_t n_frames,
                          int32_t *track_id)
{
    if (!ctx || !cfg || (!mdat && mdat_len) || (!frame_sizes && n_frames)) return ALACGPU_ERR_INVALID_ARG;
    // "FIXME: unimplemented sample size" (AlacFile.cs:570-574, :713-715)
    if (cfg->sample_size != 16 && cfg->sample_size != 24) return fail(ctx, ALACGPU_ERR_UNSUPPORTED, "sample size must be 16 or 24");
    if (cfg->num_channels != 1 && cfg->num_channels != 2) return fail(ctx, ALACGPU_ERR_UNSUPPORTED, "container channels must be 1 or 2");
    const int64_t bpsf = (int64_t)(cfg->sample_size / 8) * cfg->num_channels;
    if (cfg->max_samples_per_frame < 1 || cfg->max_samples_per_frame > kMaxFrameSamples ||
        cfg->max_samples_per_frame * bpsf > kMaxFramePcmBytes)
        return fail(ctx, ALACGPU_ERR_UNSUPPORTED, "max_samples_per_frame outside the reference's buffers (AlacFile.cs:28, AlacContext.cs:218)");
    if (cfg->rice_kmodifier < 1 || cfg->rice_kmodifier > 31 || cfg->rice_history_mult < 0 || cfg->rice_history_mult > 255 ||
        cfg->rice_initial_history < 0 || cfg->rice_initial_history > 255)
        return fail(ctx, ALACGPU_ERR_UNSUPPORTED, "rice parameters outside one cookie byte / kmodifier 1..31");
    HostTrack t{};
    t.cfg = *cfg;
    t.mdat = mdat;
    t.mdat_len = mdat_len;
    t.first_frame_offset = first_frame_offset;
    t.first_frame = ctx->sizes.size();
    t.n_frames = n_frames;
    ctx->sizes.insert(ctx->sizes.end(), frame_sizes, frame_sizes + n_frames);
    ctx->tracks.push_back(t);
    ctx->prepared = false;
    ctx->have_frame_tables = false;
    for (Device &d : ctx->devs) d.decoded = false;
    if (track_id) *track_id = (int32_t)ctx->tracks.size() - 1;
    return ALACGPU_OK;
}

int32_t alacgpu_clear_tracks(alacgpu_ctx *ctx)
{
    if (!ctx) return ALACGPU_ERR_INVALID_ARG;
    ctx->tracks.clear();
    ctx->sizes.clear();
    ctx->prepared = false;
    ctx->have_frame_tables = false;
    ctx->total_pcm = 0;
    ctx->win_lo = ctx->win_hi = 0;
    for (Device &d : ctx->devs) { d.decoded = false; d.f_lo = d.f_hi = 0; }
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
        // smallest f with prefix(f) >= total * p / n_parts (ties: fewer frames on the left)
        const long double target = (long double)total * p / n_parts;
        while (f < n_frames && (long double)acc + frame_sizes[f] / 2.0L <= target) acc += frame_sizes[f++];
        cut[p] = f;
    }
    cut[n_parts] = n_frames;
    return ALACGPU_OK;
}

// ---------------------------------------------------------------------------
static int32_t prepare_impl(alacgpu_ctx *ctx, uint64_t *total_pcm_bytes, const bool restage)
{
    const double t_begin = now_ms();
    const uint64_t n_frames = ctx->sizes.size();
    const uint32_t n_tracks = (uint32_t)ctx->tracks.size();
    const int n_dev = (int)ctx->devs.size();
    ctx->timing = alacgpu_timing{};

    std::vector<uint64_t> cut(n_dev + 1);
    alacgpu_plan_partition(ctx->sizes.data(), n_frames, n_dev, cut.data());

    uint64_t compressed = 0;
    // ---- per device: stage its byte ranges, upload the index, run K0 --------
    struct Stage {
        std::vector<FrameRef> refs;
        std::vector<TrackCfg> cfgs;
        std::vector<uint64_t> tfirst;
        uint64_t sc[2] = {0, 0};
    };
    std::vector<Stage> stage(n_dev);
    for (int g = 0; g < n_dev; g++) {
        Device &d = ctx->devs[g];
        Stage &sg = stage[g];
        d.f_lo = cut[g];
        d.f_hi = cut[g + 1];
        d.decoded = false;
        const uint64_t n_local = d.f_hi - d.f_lo;
        CU(cudaSetDevice(d.id));
        if (!restage) {
            // index only: the arena, FrameRef[] and TrackCfg[] staged earlier are still resident
            cudaEvent_t e0 = get_event(d, 0), e1 = get_event(d, 1), e2 = get_event(d, 2);
            CU(cudaEventRecord(e0, d.st));
            CU(cudaMemsetAsync(d.scalars.p, 0, 4 * sizeof(uint64_t), d.st));
            CU(cudaEventRecord(e1, d.st));
            K0Args ka{};
            ka.arena = d.arena.p; ka.refs = d.refs.p; ka.cfgs = d.cfgs.p; ka.n_frames = n_local; ka.n_tracks = n_tracks;
            ka.track_first_frame = d.track_first.p; ka.desc = d.desc.p; ka.coefs = d.coefs.p; ka.out_len = d.out_len.p;
            ka.block_sums = d.block_sums.p; ka.grand_total = d.scalars.p; ka.frame_off = d.frame_off.p;
            ka.track_start = d.track_start.p; ka.max_samples = reinterpret_cast<uint32_t *>(d.scalars.p + 1);
            if (n_local) CU(launch_k0(ka, d.st, &ctx->timing.kernel_launches));
            CU(cudaEventRecord(e2, d.st));
            CU(cudaMemcpyAsync(d.h_track_start.data(), d.track_start.p, (n_tracks + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, d.st));
            CU(cudaMemcpyAsync(sg.sc, d.scalars.p, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, d.st));
            continue;
        }
        sg.refs.resize(n_local);
        sg.cfgs.resize(std::max<uint32_t>(n_tracks, 1));
        sg.tfirst.resize(std::max<uint32_t>(n_tracks, 1));
        struct Copy { const uint8_t *src; uint64_t dst, len; };
        std::vector<Copy> copies;
        uint64_t used = 0;
        for (uint32_t t = 0; t < n_tracks; t++) {
            const HostTrack &ht = ctx->tracks[t];
            TrackCfg c{};
            c.sample_size = ht.cfg.sample_size; c.num_channels = ht.cfg.num_channels;
            c.max_samples_per_frame = ht.cfg.max_samples_per_frame;
            c.rice_history_mult = ht.cfg.rice_history_mult;
            c.rice_initial_history = ht.cfg.rice_initial_history;
            c.rice_kmodifier = ht.cfg.rice_kmodifier;
            sg.cfgs[t] = c;
            const uint64_t a = std::max<uint64_t>(ht.first_frame, d.f_lo);
            const uint64_t b = std::min<uint64_t>(ht.first_frame + ht.n_frames, d.f_hi);
            // local index of the track's first frame if this device owns it, else n_local
            sg.tfirst[t] = (ht.first_frame >= d.f_lo && ht.first_frame < d.f_hi) ? ht.first_frame - d.f_lo : n_local;
            if (a >= b) continue;
            uint64_t off = ht.first_frame_offset;       // byte offset of frame a within the track's mdat
            for (uint64_t f = ht.first_frame; f < a; f++) off += ctx->sizes[f];
            const uint64_t base = align_up(used, 16);
            const uint64_t src_lo = std::min(off, ht.mdat_len);
            uint64_t cur = off;
            for (uint64_t f = a; f < b; f++) {
                FrameRef r;
                const uint64_t avail = cur < ht.mdat_len ? ht.mdat_len - cur : 0;
                r.len = (uint32_t)std::min<uint64_t>(ctx->sizes[f], avail);   // short read (MyStream.cs:47-52)
                r.off = r.len ? base + (cur - src_lo) : base;
                r.track = t;
                sg.refs[f - d.f_lo] = r;
                compressed += r.len;
                cur += ctx->sizes[f];
            }
            const uint64_t src_hi = std::min(cur, ht.mdat_len);
            if (src_hi > src_lo) copies.push_back({ht.mdat + src_lo, base, src_hi - src_lo});
            used = base + (src_hi - src_lo);
        }
        d.arena_used = used;
        CU(d.arena.reserve(used + kArenaTail));
        CU(d.refs.reserve(std::max<uint64_t>(n_local, 1)));
        CU(d.cfgs.reserve(sg.cfgs.size()));
        CU(d.desc.reserve(std::max<uint64_t>(n_local, 1)));
        CU(d.coefs.reserve(std::max<uint64_t>(n_local, 1)));
        CU(d.out_len.reserve(std::max<uint64_t>(n_local, 1)));
        CU(d.frame_off.reserve(std::max<uint64_t>(n_local, 1)));
        CU(d.block_sums.reserve(std::max<uint32_t>(k0_scan_blocks(n_local), 1)));
        CU(d.track_first.reserve(sg.cfgs.size()));
        CU(d.track_start.reserve(n_tracks + 1));
        CU(d.track_shift.reserve(sg.cfgs.size()));
        CU(d.scalars.reserve(4));
        d.h_track_start.assign(n_tracks + 1, 0);

        cudaEvent_t e0 = get_event(d, 0), e1 = get_event(d, 1), e2 = get_event(d, 2);
        CU(cudaEventRecord(e0, d.st));
        for (const Copy &c : copies)
            CU(cudaMemcpyAsync(d.arena.p + c.dst, c.src, c.len, cudaMemcpyHostToDevice, d.st));
        CU(cudaMemsetAsync(d.arena.p + used, 0, kArenaTail, d.st));
        if (n_local) CU(cudaMemcpyAsync(d.refs.p, sg.refs.data(), n_local * sizeof(FrameRef), cudaMemcpyHostToDevice, d.st));
        CU(cudaMemcpyAsync(d.cfgs.p, sg.cfgs.data(), sg.cfgs.size() * sizeof(TrackCfg), cudaMemcpyHostToDevice, d.st));
        CU(cudaMemcpyAsync(d.track_first.p, sg.tfirst.data(), sg.tfirst.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, d.st));
        CU(cudaMemsetAsync(d.scalars.p, 0, 4 * sizeof(uint64_t), d.st));
        CU(cudaEventRecord(e1, d.st));
        K0Args ka{};
        ka.arena = d.arena.p; ka.refs = d.refs.p; ka.cfgs = d.cfgs.p; ka.n_frames = n_local; ka.n_tracks = n_tracks;
        ka.track_first_frame = d.track_first.p; ka.desc = d.desc.p; ka.coefs = d.coefs.p; ka.out_len = d.out_len.p;
        ka.block_sums = d.block_sums.p; ka.grand_total = d.scalars.p; ka.frame_off = d.frame_off.p;
        ka.track_start = d.track_start.p; ka.max_samples = reinterpret_cast<uint32_t *>(d.scalars.p + 1);
        if (n_local) {
            CU(launch_k0(ka, d.st, &ctx->timing.kernel_launches));
        } else {
            CU(cudaMemsetAsync(d.track_start.p, 0, (n_tracks + 1) * sizeof(uint64_t), d.st));
        }
        CU(cudaEventRecord(e2, d.st));
        CU(cudaMemcpyAsync(d.h_track_start.data(), d.track_start.p, (n_tracks + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, d.st));
        CU(cudaMemcpyAsync(sg.sc, d.scalars.p, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, d.st));
    }
    for (int g = 0; g < n_dev; g++) {
        Device &d = ctx->devs[g];
        CU(cudaSetDevice(d.id));
        CU(cudaStreamSynchronize(d.st));
        d.total_unpadded = stage[g].sc[0];
        d.max_samples = (uint32_t)(stage[g].sc[1] & 0xffffffffu);
        float ms = 0;
        cudaEventElapsedTime(&ms, d.events[0], d.events[1]); ctx->timing.h2d_ms = std::max(ctx->timing.h2d_ms, ms);
        cudaEventElapsedTime(&ms, d.events[1], d.events[2]); ctx->timing.index_ms = std::max(ctx->timing.index_ms, ms);
    }

    // ---- global layout: tracks at 256-byte aligned offsets -------------------
    std::vector<uint64_t> dev_base(n_dev + 1, 0);
    for (int g = 0; g < n_dev; g++) dev_base[g + 1] = dev_base[g] + ctx->devs[g].total_unpadded;
    // unpadded global start of every track (+ grand total at the end)
    std::vector<uint64_t> ts(n_tracks + 1, dev_base[n_dev]);
    for (uint32_t t = 0; t < n_tracks; t++) {
        const HostTrack &ht = ctx->tracks[t];
        // device that owns the track's first frame
        int g = 0;
        while (g + 1 < n_dev && ht.first_frame >= ctx->devs[g].f_hi) g++;
        if (ht.first_frame >= n_frames) ts[t] = dev_base[n_dev];
        else ts[t] = dev_base[g] + ctx->devs[g].h_track_start[t];
    }
    uint64_t pos = 0;
    std::vector<uint64_t> shift(std::max<uint32_t>(n_tracks, 1), 0);
    for (uint32_t t = 0; t < n_tracks; t++) {
        HostTrack &ht = ctx->tracks[t];
        pos = align_up(pos, kTrackAlign);
        ht.pcm_off = pos;
        ht.pcm_len = ts[t + 1] - ts[t];
        shift[t] = pos - ts[t];          // mod 2^64
        pos += ht.pcm_len;
    }
    ctx->total_pcm = pos;
    uint64_t samples = 0;
    for (const HostTrack &ht : ctx->tracks) samples += ht.pcm_len / (uint64_t)(ht.cfg.sample_size / 8);
    for (int g = 0; g < n_dev; g++) {
        Device &d = ctx->devs[g];
        CU(cudaSetDevice(d.id));
        // device-local frame_off is relative to the device's first frame: fold dev_base into the shift
        std::vector<uint64_t> sh(shift);
        for (uint64_t &v : sh) v += dev_base[g];
        CU(cudaMemcpyAsync(d.track_shift.p, sh.data(), sh.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, d.st));
        CU(cudaStreamSynchronize(d.st));
        // PCM byte range of the shard
        if (d.f_hi > d.f_lo) {
            const size_t t0 = track_of(ctx, d.f_lo);
            d.pcm_first = dev_base[g] + shift[t0];                   // offset of frame f_lo
            d.pcm_lo = d.pcm_first / kTrackAlign * kTrackAlign;
            const size_t t1 = track_of(ctx, d.f_hi - 1);
            d.pcm_hi = dev_base[g + 1] + shift[t1];
        } else {
            d.pcm_lo = d.pcm_hi = d.pcm_first = 0;
        }
    }
    if (restage) ctx->compressed_bytes = compressed;
    ctx->timing.compressed_bytes = ctx->compressed_bytes;
    ctx->timing.pcm_bytes = dev_base[n_dev];
    ctx->timing.samples = samples;
    ctx->timing.total_ms = (float)(now_ms() - t_begin);
    ctx->prepared = true;
    ctx->have_frame_tables = false;
    ctx->win_lo = ctx->win_hi = 0;
    if (total_pcm_bytes) *total_pcm_bytes = ctx->total_pcm;
    return ALACGPU_OK;
}

int32_t alacgpu_prepare(alacgpu_ctx *ctx, uint64_t *total_pcm_bytes)
{
    if (!ctx) return ALACGPU_ERR_INVALID_ARG;
    if (ctx->prepared) { if (total_pcm_bytes) *total_pcm_bytes = ctx->total_pcm; return ALACGPU_OK; }
    return prepare_impl(ctx, total_pcm_bytes, true);
}

int32_t alacgpu_reindex(alacgpu_ctx *ctx)
{
    if (!ctx) return ALACGPU_ERR_INVALID_ARG;
    if (!ctx->prepared) return prepare_impl(ctx, nullptr, true);
    return prepare_impl(ctx, nullptr, false);
}

// ---------------------------------------------------------------------------
int32_t alacgpu_decode_all(alacgpu_ctx *ctx, uint8_t *pcm_dst, uint64_t cap, uint64_t *track_pcm_off,
                           uint64_t *track_pcm_len, int32_t *frame_status)
{
    if (!ctx) return ALACGPU_ERR_INVALID_ARG;
    if (!ctx->prepared) {
        int32_t r = alacgpu_prepare(ctx, nullptr);
        if (r != ALACGPU_OK) return r;
    }
    if (pcm_dst && cap < ctx->total_pcm) return fail(ctx, ALACGPU_ERR_CAPACITY, "pcm_dst smaller than alacgpu_prepare's total");
    const double t_begin = now_ms();
    const int n_dev = (int)ctx->devs.size();
    const uint32_t chunk_frames = ctx->opts.chunk_frames;
    uint32_t launches = 0, chunks_total = 0;

    // ---- issue: all devices, all chunks (asynchronous) -----------------------
    for (int g = 0; g < n_dev; g++) {
        Device &d = ctx->devs[g];
        const uint64_t n_local = d.f_hi - d.f_lo;
        if (n_local == 0) continue;
        CU(cudaSetDevice(d.id));
        const uint32_t ns = std::max<uint32_t>((d.max_samples + 31u) & ~31u, 32u);
        const uint32_t cf = (uint32_t)std::min<uint64_t>(chunk_frames, (n_local + 31) & ~31ull);
        CU(d.planes.reserve((size_t)cf * 2u * ns));
        CU(d.pcm.reserve(d.pcm_hi - d.pcm_lo + 64));
        // gaps between tracks (alignment padding) are defined as zero
        ChunkArgs ca{};
        ca.arena = d.arena.p; ca.refs = d.refs.p; ca.cfgs = d.cfgs.p; ca.desc = d.desc.p; ca.coefs = d.coefs.p;
        ca.frame_off = d.frame_off.p; ca.track_shift = d.track_shift.p; ca.planes = d.planes.p;
        ca.pcm = d.pcm.p; ca.pcm_base = d.pcm_lo; ca.ns = ns;
        size_t ev = 4;
        CU(cudaEventRecord(get_event(d, 3), d.st));
        for (uint64_t f0 = 0; f0 < n_local; f0 += cf) {
            ca.f0 = f0;
            ca.n = (uint32_t)std::min<uint64_t>(cf, n_local - f0);
            cudaEvent_t a0 = get_event(d, ev), a1 = get_event(d, ev + 1), a2 = get_event(d, ev + 2), a3 = get_event(d, ev + 3);
            ev += 4;
            CU(cudaEventRecord(a0, d.st));
            CU(launch_k1(ca, lanes_for(ctx, ca.n), d.st, &launches));
            CU(cudaEventRecord(a1, d.st));
            CU(launch_k2(ca, d.st, &launches));
            CU(cudaEventRecord(a2, d.st));
            CU(launch_k3(ca, d.st, &launches));
            CU(cudaEventRecord(a3, d.st));
            chunks_total++;
        }
        CU(cudaEventRecord(get_event(d, ev), d.st));
        // alignment gaps between tracks are defined as zero bytes
        {
            std::vector<uint64_t> gaps;   // (offset within d.pcm, length) pairs
            const size_t t0 = track_of(ctx, d.f_lo), t1 = track_of(ctx, d.f_hi - 1);
            for (size_t t = t0; t < t1; t++) {
                const uint64_t end = ctx->tracks[t].pcm_off + ctx->tracks[t].pcm_len, nxt = ctx->tracks[t + 1].pcm_off;
                if (nxt > end && end >= d.pcm_lo) { gaps.push_back(end - d.pcm_lo); gaps.push_back(nxt - end); }
            }
            if (d.pcm_first > d.pcm_lo) { gaps.push_back(0); gaps.push_back(d.pcm_first - d.pcm_lo); }
            for (size_t k = 0; k < gaps.size(); k += 2)
                CU(cudaMemsetAsync(d.pcm.p + gaps[k], 0, gaps[k + 1], d.st));
        }
        // ---- PCM to the caller: the shard is contiguous in the global layout -------
        if (pcm_dst) {
            CU(cudaEventRecord(get_event(d, ev + 1), d.st));
            if (d.pcm_hi > d.pcm_first)
                CU(cudaMemcpyAsync(pcm_dst + d.pcm_first, d.pcm.p + (d.pcm_first - d.pcm_lo), d.pcm_hi - d.pcm_first,
                                   cudaMemcpyDeviceToHost, d.st));
            CU(cudaEventRecord(get_event(d, ev + 2), d.st));
        }
        d.decoded = true;
    }
    // ---- wait + timings --------------------------------------------------------
    float k1 = 0, k2 = 0, k3 = 0, kall = 0, d2h = 0;
    for (int g = 0; g < n_dev; g++) {
        Device &d = ctx->devs[g];
        const uint64_t n_local = d.f_hi - d.f_lo;
        if (n_local == 0) continue;
        CU(cudaSetDevice(d.id));
        CU(cudaStreamSynchronize(d.st));
        const uint32_t ns = std::max<uint32_t>((d.max_samples + 31u) & ~31u, 32u);
        (void)ns;
        const uint32_t cf = (uint32_t)std::min<uint64_t>(chunk_frames, (n_local + 31) & ~31ull);
        size_t ev = 4;
        float s1 = 0, s2 = 0, s3 = 0, ms = 0;
        for (uint64_t f0 = 0; f0 < n_local; f0 += cf) {
            cudaEventElapsedTime(&ms, d.events[ev], d.events[ev + 1]); s1 += ms;
            cudaEventElapsedTime(&ms, d.events[ev + 1], d.events[ev + 2]); s2 += ms;
            cudaEventElapsedTime(&ms, d.events[ev + 2], d.events[ev + 3]); s3 += ms;
            ev += 4;
        }
        cudaEventElapsedTime(&ms, d.events[3], d.events[ev]);
        k1 = std::max(k1, s1); k2 = std::max(k2, s2); k3 = std::max(k3, s3); kall = std::max(kall, ms);
        if (pcm_dst) { cudaEventElapsedTime(&ms, d.events[ev + 1], d.events[ev + 2]); d2h = std::max(d2h, ms); }
    }
    ctx->timing.entropy_ms = k1; ctx->timing.lpc_ms = k2; ctx->timing.stereo_ms = k3;
    ctx->timing.kernels_ms = kall; ctx->timing.d2h_ms = d2h;
    ctx->timing.kernel_launches += launches;
    ctx->timing.chunks = chunks_total;
    ctx->timing.total_ms = (float)(now_ms() - t_begin);
    for (size_t t = 0; t < ctx->tracks.size(); t++) {
        if (track_pcm_off) track_pcm_off[t] = ctx->tracks[t].pcm_off;
        if (track_pcm_len) track_pcm_len[t] = ctx->tracks[t].pcm_len;
    }
    if (frame_status) {
        for (int g = 0; g < n_dev; g++) {
            Device &d = ctx->devs[g];
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
static int32_t ensure_frame_tables(alacgpu_ctx *ctx)
{
    if (ctx->have_frame_tables) return ALACGPU_OK;
    const uint64_t n_frames = ctx->sizes.size();
    ctx->h_frame_off.assign(n_frames, 0);
    ctx->h_frame_len.assign(n_frames, 0);
    ctx->h_status.assign(n_frames, 0);
    for (Device &d : ctx->devs) {
        const uint64_t n_local = d.f_hi - d.f_lo;
        if (!n_local) continue;
        CU(cudaSetDevice(d.id));
        std::vector<uint64_t> sh(ctx->tracks.size());
        CU(cudaMemcpy(ctx->h_frame_off.data() + d.f_lo, d.frame_off.p, n_local * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(ctx->h_frame_len.data() + d.f_lo, d.out_len.p, n_local * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(sh.data(), d.track_shift.p, sh.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        std::vector<FrameDesc> h(n_local);
        CU(cudaMemcpy(h.data(), d.desc.p, n_local * sizeof(FrameDesc), cudaMemcpyDeviceToHost));
        size_t t = track_of(ctx, d.f_lo);
        for (uint64_t i = 0; i < n_local; i++) {
            const uint64_t gidx = d.f_lo + i;
            while (t + 1 < ctx->tracks.size() && gidx >= ctx->tracks[t].first_frame + ctx->tracks[t].n_frames) t++;
            ctx->h_frame_off[gidx] += sh[t];
            ctx->h_status[gidx] = h[i].status;
        }
    }
    ctx->have_frame_tables = true;
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
    bool decoded = ctx->prepared;
    for (Device &d : ctx->devs) if (d.f_hi > d.f_lo && !d.decoded) decoded = false;
    if (!decoded) {
        r = alacgpu_decode_all(ctx, nullptr, 0, nullptr, nullptr, nullptr);
        if (r) return r;
    }
    r = ensure_frame_tables(ctx);
    if (r) return r;
    const uint64_t gidx = ht.first_frame + frame_idx;
    const uint64_t off = ctx->h_frame_off[gidx];
    const uint32_t len = ctx->h_frame_len[gidx];
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
    if (!ctx->prepared) return fail(ctx, ALACGPU_ERR_STATE, "alacgpu_prepare has not run");
    const HostTrack &ht = ctx->tracks[track];
    if (frame_idx >= ht.n_frames) return fail(ctx, ALACGPU_ERR_RANGE, "frame index out of range");
    r = ensure_frame_tables(ctx);
    if (r) return r;
    *n_samples = ctx->h_frame_len[ht.first_frame + frame_idx] / (uint32_t)((ht.cfg.sample_size / 8) * ht.cfg.num_channels);
    return ALACGPU_OK;
}

int32_t alacgpu_frame_status(alacgpu_ctx *ctx, int32_t track, uint32_t frame_idx, int32_t *status)
{
    int32_t r = check_track(ctx, track);
    if (r) return r;
    if (!status) return ALACGPU_ERR_INVALID_ARG;
    if (!ctx->prepared) return fail(ctx, ALACGPU_ERR_STATE, "alacgpu_prepare has not run");
    const HostTrack &ht = ctx->tracks[track];
    if (frame_idx >= ht.n_frames) return fail(ctx, ALACGPU_ERR_RANGE, "frame index out of range");
    ctx->have_frame_tables = false;      // statuses may have changed since the last decode
    r = ensure_frame_tables(ctx);
    if (r) return r;
    *status = ctx->h_status[ht.first_frame + frame_idx];
    return ALACGPU_OK;
}

int32_t alacgpu_track_pcm_bytes(alacgpu_ctx *ctx, int32_t track, uint64_t *off, uint64_t *len)
{
    int32_t r = check_track(ctx, track);
    if (r) return r;
    if (!ctx->prepared) return fail(ctx, ALACGPU_ERR_STATE, "alacgpu_prepare has not run");
    if (off) *off = ctx->tracks[track].pcm_off;
    if (len) *len = ctx->tracks[track].pcm_len;
    return ALACGPU_OK;
}

int32_t alacgpu_get_timing(alacgpu_ctx *ctx, alacgpu_timing *out)
{
    if (!ctx || !out) return ALACGPU_ERR_INVALID_ARG;
    *out = ctx->timing;
    return ALACGPU_OK;
}

int32_t alacgpu_device_pcm(alacgpu_ctx *ctx, int32_t dev_slot, void **dptr, uint64_t *shard_off, uint64_t *shard_len)
{
    if (!ctx || dev_slot < 0 || (size_t)dev_slot >= ctx->devs.size()) return ALACGPU_ERR_INVALID_ARG;
    Device &d = ctx->devs[dev_slot];
    if (!d.decoded && d.f_hi > d.f_lo) return fail(ctx, ALACGPU_ERR_STATE, "no decoded PCM on this device");
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
        if (!d.decoded) return fail(ctx, ALACGPU_ERR_STATE, "no decoded PCM on this device");
        // intersection of [off, off+len) with this shard, on 8-byte word boundaries of the global layout
        uint64_t lo = std::max(off, d.pcm_lo), hi = std::min(off + len, d.pcm_hi);
        if (hi <= lo) continue;
        if (lo & 7) return fail(ctx, ALACGPU_ERR_INVALID_ARG, "checksum range must start on an 8-byte boundary of each shard");
        CU(cudaSetDevice(d.id));
        CU(cudaMemsetAsync(d.scalars.p + 2, 0, sizeof(uint64_t), d.st));
        CU(launch_checksum(d.pcm.p + (lo - d.pcm_lo), lo, hi - lo, d.scalars.p + 2, d.st));
        uint64_t part = 0;
        CU(cudaMemcpyAsync(&part, d.scalars.p + 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, d.st));
        CU(cudaStreamSynchronize(d.st));
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
