// Host-visible launch interface of the four kernels (K0 index, K1 entropy,
// K2 LPC, K3 stereo/pack).  Plain structs so runtime.cu stays free of kernel
// details.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace alacgpu {

struct FrameRef;
struct TrackCfg;
struct FrameDesc;
struct FrameCoefs;

struct K0Args {
    const uint8_t *arena;
    const FrameRef *refs;
    const TrackCfg *cfgs;
    FrameDesc *desc;
    FrameCoefs *coefs;
    const uint32_t *expect_len;   // PCM bytes per frame as laid out by the host
    uint32_t *mismatch;           // incremented when K0's size differs from expect_len
    uint64_t f0;                  // first frame of this launch (device-local index)
    uint32_t n;
    uint64_t arena_bytes;         // (ALACGPU_CHECKED) staged bytes + tail padding
    uint32_t *check;              // (ALACGPU_CHECKED) violation word
};
cudaError_t launch_k0(const K0Args &a, cudaStream_t st, uint32_t *launches);

// One pipeline chunk = frames [f0, f0 + n) of the device's frame list; planes
// are stream-major: row (slot * 2 + ch) holds the ns int32 residual / predicted
// samples of channel ch of the chunk's frame `slot` (rows 128-byte aligned).
struct ChunkArgs {
    const uint8_t *arena;
    const FrameRef *refs;
    const TrackCfg *cfgs;
    FrameDesc *desc;
    const FrameCoefs *coefs;
    const uint64_t *frame_off;     // PCM byte offset of every frame in the global layout
    int32_t *planes;
    uint8_t *pcm;                  // device PCM base (global layout offset `pcm_base` maps to pcm[0])
    uint64_t pcm_base;
    uint64_t f0;
    uint32_t n;
    uint32_t ns;                   // plane row stride in samples (multiple of 32)
    uint32_t max_sf;               // most sample-frames any frame of the batch emits
    uint32_t *perm;                // K2 work lists, heaviest first: [0, 2n) one-lane streams, [2n, 4n) multi-lane streams
    uint32_t *perm_count;          // [0] = one-lane streams, [1] = multi-lane streams
    int use_quads;                 // small (latency-bound) chunk: streams that get four lanes each (thresholds, k2_lpc.cuh lpc_quad; bit 16: eight lanes); 0 = none
    uint32_t *progress;            // fused launch: residuals published per stream by the entropy lanes (2n entries, zeroed before the launch)
    uint32_t *lpc_done;            // fused launch: predicted samples published per stream by the LPC lanes (2n entries, zeroed)
    uint32_t *pack_next;           // fused launch: next pack task (zeroed)
    uint8_t *lpc_flag;             // per stream: 1 if the stream is on the LPC work list (written by the sort)
    uint32_t *faults;              // device counter of frames flagged FS_INTERNAL (the runtime then re-decodes unfused)
    // extents for the ALACGPU_CHECKED build (alacgpu_device.cuh); `check` is its violation word
    uint32_t *check;
    uint64_t arena_bytes;          // staged bytes + tail padding
    uint64_t plane_bytes;          // bytes of this slot's planes
    uint64_t pcm_bytes;            // bytes of the device PCM buffer (or of the mapped destination)
    // frame-lane path (kf_frame.cu): work lists of chunk-local frame slots, every class padded to whole warps
    uint32_t *kf_list;             // [0, kf_cap) phase A, [kf_cap, 2 kf_cap) phase B, [2 kf_cap, 2 kf_cap + n) pack-only frames
    uint32_t *kf_count;            // [0] phase A entries (padded), [1] phase B entries (padded), [2] pack-only frames;
                                   // [8 ..] scratch of the sort (class histograms and cursors)
    uint32_t kf_cap;               // entries per padded list: n + kKfClasses * 31 rounded up to 32
    uint32_t *bstart;              // per frame slot: bit offset (from the frame start) where channel B's Rice stream starts
    uint32_t kf_row;               // bytes per frame of the channel-A plane: ns x 2 (only 16-bit tracks on the device) or ns x 4
};
cudaError_t launch_k1(const ChunkArgs &a, int lanes_per_warp, cudaStream_t st, uint32_t *launches);
cudaError_t launch_sort(const ChunkArgs &a, cudaStream_t st, uint32_t *launches);   // K2 / K12 work list
cudaError_t launch_k2(const ChunkArgs &a, cudaStream_t st, uint32_t *launches);
// fused entropy + LPC (one launch, LPC warps consume residuals as the entropy lanes publish them)
cudaError_t launch_k12(const ChunkArgs &a, int lanes_per_warp, cudaStream_t st, uint32_t *launches);
// fused entropy + LPC + pack (pack warps stream PCM out while the frames are still being decoded),
// followed by the fix-up of frames whose decode failed after part of their PCM had been written
cudaError_t launch_k123(const ChunkArgs &a, int lanes_per_warp, cudaStream_t st, uint32_t *launches);
cudaError_t launch_fix(const ChunkArgs &a, cudaStream_t st, uint32_t *launches);
cudaError_t launch_k3(const ChunkArgs &a, cudaStream_t st, uint32_t *launches);
// frame-lane path for machine-filling chunks: class sort -> phase A (entropy + LPC of channel A in one lane per
// frame; mono frames packed straight to PCM, stereo frames leave channel A in a half-width plane) -> phase B
// (entropy + LPC of channel B, un-mix with the plane, PCM) -> pack-only frames (escape / failed) -> fix-up
constexpr uint32_t kKfClasses = 256;
inline uint32_t kf_list_cap(uint32_t n) { return (n + kKfClasses * 31u + 31u) & ~31u; }
inline size_t kf_list_words(uint32_t n) { return 2u * (size_t)kf_list_cap(n) + n; }
constexpr size_t kKfCountWords = 3200;   // counts, class histograms / cursors and the per-SM schedule (kf_frame.cu)
cudaError_t launch_kf_sort(const ChunkArgs &a, cudaStream_t st, uint32_t *launches);   // class sort -> work lists
cudaError_t launch_kf_a(const ChunkArgs &a, cudaStream_t st, uint32_t *launches);      // phase A
cudaError_t launch_kf_b(const ChunkArgs &a, cudaStream_t st, uint32_t *launches);      // phase B
cudaError_t launch_kf_rest(const ChunkArgs &a, cudaStream_t st, uint32_t *launches);   // pack-only frames + failed-frame fix-up

// position-weighted checksum of device bytes (see alacgpu_pcm_checksum)
cudaError_t launch_checksum(const uint8_t *pcm, uint64_t global_off, uint64_t len, uint64_t *d_sum, cudaStream_t st);

}  // namespace alacgpu
