// alacgpu_device.cuh -- device-side data layout shared by the four kernels.
//
// HBM layout (per device), see DESIGN.md "Data layout":
//   arena      : every staged mdat byte, tracks back to back, each track start
//                16-byte aligned, 256 zero bytes of tail padding.
//   FrameRef[] : {arena byte offset, byte length, track} per frame -- the
//                device-resident form of the demuxer's stsz table
//                (ALACDecoder/DemuxResT.cs:28; sequential addressing of
//                ALACDecoder/AlacContext.cs:194-195).
//   TrackCfg[] : the 'alac' cookie fields (ALACDecoder/AlacFile.cs:63-93).
//   FrameDesc[]/FrameCoefs[] : K0's parse of each frame header
//                (AlacFile.cs:435-475 / :584-641).
//   planes     : int32 residual / predicted samples of one pipeline chunk,
//                stream-major: row (slot*2 + ch) = the NS samples of channel
//                ch of the chunk's frame `slot`, rows 128-byte aligned and
//                padded (NS = align32(max N) + 32).  K1 writes a lane's row 8
//                samples at a time, K2 rewrites it in place 4 at a time, K3
//                streams 8 sample-frames per thread out of the two rows.
//   pcm        : interleaved little-endian PCM, frame after frame.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace alacgpu {

constexpr int kTile = 32;            // frames per tile == lanes per warp
constexpr int kMaxFrameSamples = 16384;   // AlacFile.cs:28 BufferSize
constexpr int kMaxFramePcmBytes = 65536;  // AlacContext.cs:218

// Frame status codes == ALACGPU_FRAME_* (include/alacgpu.h).
enum : uint8_t {
    FS_OK = 0, FS_BAD_TAG = 1, FS_PRED_TYPE = 2, FS_TOO_MANY = 3, FS_OVERRUN = 4,
    FS_BAD_RSS = 5, FS_HISTORY = 6, FS_RUN_OVERFLOW = 7, FS_ORDER0_LONG = 8, FS_INTERNAL = 9
};

enum : uint8_t { FF_STEREO = 1, FF_ESCAPE = 2 };

struct FrameRef {
    uint64_t off;      // byte offset of the frame in the arena
    uint32_t len;      // bytes (stsz, truncated to what was staged)
    uint32_t track;
};

struct TrackCfg {
    int32_t sample_size;
    int32_t num_channels;
    int32_t max_samples_per_frame;
    int32_t rice_history_mult;
    int32_t rice_initial_history;
    int32_t rice_kmodifier;
    int32_t pad0, pad1;
};

struct __align__(16) FrameDesc {
    uint32_t data_bit;     // bit offset (from the frame start) of Rice stream A, or of the raw samples (escape)
    uint32_t shift_bit;    // bit offset of the interleaved wasted-byte block (valid if ub != 0)
    uint32_t out_len;      // PCM bytes this frame produces
    uint16_t n;            // samples per channel
    uint8_t flags;         // FF_*
    uint8_t ub;            // wasted bytes (0 for escape frames: AlacFile.cs:525,697)
    uint8_t status;        // FS_*
    uint8_t rss;           // read sample size: sampleSize - 8*ub (+1 stereo)
    uint8_t mix_shift;
    uint8_t mix_weight;
    uint8_t order[2];
    uint8_t quant[2];
    uint8_t rice_mod[2];
    uint8_t status0;       // K0's verdict, never changed afterwards (`status` may be raised by the entropy stage)
    uint8_t pad[5];
};
static_assert(sizeof(FrameDesc) == 32, "FrameDesc is 32 bytes");

struct __align__(16) FrameCoefs {
    int16_t c[2][32];
};

// ---- ALACGPU_CHECKED: bounds assertions (the stand-in for compute-sanitizer, which is closed on this GPU pool) --
// Built with -DALACGPU_CHECKED (libalacgpu_checked.so) every index the kernels derive from stream contents is
// compared with the extent of the buffer it goes into before it is used; a violation sets one bit of a device
// word, the access is skipped, and the runtime fails the call with ALACGPU_ERR_STATE naming the bits.  The
// release build compiles the checks away.
enum : uint32_t {
    CK_ARENA = 0,        // bitstream read outside the staged bytes + tail padding
    CK_PLANE = 1,        // plane row / sample index outside the slot's planes
    CK_LIST = 2,         // work-list / perm index outside the list
    CK_PROGRESS = 3,     // progress / done word index outside the slot's words
    CK_PCM = 4,          // PCM write outside the device's PCM buffer
    CK_RING = 5,         // shared-memory ring over- or under-run
    CK_FRAME = 6,        // frame index outside the chunk
};
#ifdef ALACGPU_CHECKED
#define ALACGPU_CHECK(word, cond, code) ((cond) ? true : (atomicOr((word), 1u << (code)), false))
#else
#define ALACGPU_CHECK(word, cond, code) (true)
#endif

// ---- big-endian bit access into the arena --------------------------------
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// n bits (1..32) at absolute bit position `pos` of the arena, MSB first.
// Two aligned 32-bit loads; the arena's tail padding keeps the second load in
// bounds.
__device__ __forceinline__ uint32_t arena_bits(const uint32_t *__restrict__ arena32, uint64_t pos, int n)
{
    const uint64_t w = pos >> 5;
    const int off = (int)(pos & 31);
    const uint32_t hi = bswap32(__ldg(arena32 + w));
    const uint32_t lo = bswap32(__ldg(arena32 + w + 1));
    const uint32_t win = __funnelshift_l(lo, hi, off);
    return win >> (32 - n);
}

__device__ __forceinline__ int32_t sext(int32_t v, int bits)
{
    // (v << (32-bits)) >> (32-bits), AlacFile.cs:278-279,309-310 (C# masks the count to 5 bits)
    const int mv = (32 - bits) & 31;
    return (int32_t)((uint32_t)v << mv) >> mv;
}

}  // namespace alacgpu
