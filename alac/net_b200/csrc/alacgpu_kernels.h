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
    uint64_t n_frames;
    uint32_t n_tracks;
    const uint64_t *track_first_frame;  // device, n_tracks entries
    FrameDesc *desc;
    FrameCoefs *coefs;
    uint32_t *out_len;
    uint64_t *block_sums;     // k0_scan_blocks(n_frames) entries
    uint64_t *grand_total;    // 1 entry
    uint64_t *frame_off;      // n_frames entries: unpadded exclusive scan of out_len
    uint64_t *track_start;    // n_tracks + 1 entries
    uint32_t *max_samples;    // 1 entry, pre-zeroed: max sample-frames emitted by any frame
};
cudaError_t launch_k0(const K0Args &a, cudaStream_t st, uint32_t *launches);
uint32_t k0_scan_blocks(uint64_t n_frames);

// One pipeline chunk = frames [f0, f0 + n) of the device's frame list; planes
// hold ceil(n / 32) tiles of 2 channels x ns samples x 32 lanes int32.
struct ChunkArgs {
    const uint8_t *arena;
    const FrameRef *refs;
    const TrackCfg *cfgs;
    FrameDesc *desc;
    const FrameCoefs *coefs;
    const uint64_t *frame_off;     // unpadded PCM offsets
    const uint64_t *track_shift;   // per track: padded start - unpadded start
    int32_t *planes;
    uint8_t *pcm;                  // device PCM base (global layout offset `pcm_base` maps to pcm[0])
    uint64_t pcm_base;
    uint64_t f0;
    uint32_t n;
    uint32_t ns;                   // plane stride in samples
};
cudaError_t launch_k1(const ChunkArgs &a, int lanes_per_warp, cudaStream_t st, uint32_t *launches);
cudaError_t launch_k2(const ChunkArgs &a, cudaStream_t st, uint32_t *launches);
cudaError_t launch_k3(const ChunkArgs &a, cudaStream_t st, uint32_t *launches);

// position-weighted checksum of device bytes (see alacgpu_pcm_checksum)
cudaError_t launch_checksum(const uint8_t *pcm, uint64_t global_off, uint64_t len, uint64_t *d_sum, cudaStream_t st);

}  // namespace alacgpu
