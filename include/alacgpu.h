/*
 * alacgpu.h -- C ABI of libalacgpu.so, the B200 (sm_100a) ALAC frame decoder
 * that stands behind the teekay/ALAC.NET decode path.
 *
 * This is the only public native header.  The C# host (csharp/AlacNet/
 * AlacGpuNative.cs) mirrors it 1:1 with [DllImport("alacgpu")]; the C++ host
 * mirror (alac/net_b200/host/) and the Python ctypes binding
 * (alac/net_b200/_native.py) call exactly these entry points.
 *
 * What each group replaces in the reference (paths relative to the reference
 * repository root):
 *
 *   alacgpu_create / alacgpu_destroy
 *       the per-stream decoder object: `new AlacFile(sampleSize, channels)` +
 *       `SetInfo(codecData)` in the AlacContext constructor
 *       (ALACDecoder/AlacContext.cs:54-55; ALACDecoder/AlacFile.cs:16-20,63-93)
 *       and AlacContext.Dispose (AlacContext.cs:297-318).
 *   alacgpu_add_track
 *       the hand-over of the demuxer's tables to the decoder: DemuxResT's
 *       SampleByteSize[] (stsz), CodecData (the 'alac' cookie) and the mdat
 *       position left by QtMovieT.ReadHeader (ALACDecoder/DemuxResT.cs:22-34,
 *       ALACDecoder/QTMovieT.cs:51-109,412-523,561-613,724-734), plus the
 *       sequential `MyStream.Read(sampleByteSize, _readBuffer, 0)` that feeds
 *       every frame (AlacContext.cs:194-195): the whole mdat is staged once.
 *   alacgpu_prepare / alacgpu_decode_all
 *       the frame pump: the loop of AlacContext.Read -> UnpackSamples ->
 *       AlacFile.DecodeFrame -> FormatSamples over all frames
 *       (AlacContext.cs:163-204,214-256; AlacFile.cs:428-719), batched over
 *       every frame of every added track.
 *   alacgpu_read_frame
 *       one call of `int AlacContext.Read(byte[] buffer)` (AlacContext.cs:163):
 *       exactly one frame's little-endian interleaved PCM, 0 bytes past the end.
 *   alacgpu_frame_count / alacgpu_frame_samples / alacgpu_track_pcm_bytes
 *       what GetNumSamples / TryGetSampleInfo / SetPosition derive from the
 *       sample tables (AlacContext.cs:108-156,262-295).
 *
 * Conventions: every function returns an int32 status (0 = ALACGPU_OK, <0 =
 * error, see alacgpu_strerror); nothing throws or aborts across the boundary.
 * Handles are opaque and freed only by alacgpu_destroy.  All arrays are
 * caller-owned and only borrowed for the duration of the call.  A context is
 * single-caller (like the reference's AlacContext, which is not thread-safe);
 * distinct contexts may be used from distinct threads.
 *
 * There is NO CPU fallback: if no CUDA device is usable alacgpu_create fails
 * with ALACGPU_ERR_NO_DEVICE.
 */
#ifndef ALACGPU_H
#define ALACGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define ALACGPU_API __declspec(dllexport)
#else
#define ALACGPU_API __attribute__((visibility("default")))
#endif

#define ALACGPU_ABI_VERSION 2

/* ---- call status -------------------------------------------------------- */
enum {
    ALACGPU_OK = 0,
    ALACGPU_ERR_INVALID_ARG = -1,
    ALACGPU_ERR_NO_DEVICE = -2,       /* no usable CUDA device (no CPU fallback) */
    ALACGPU_ERR_CUDA = -3,            /* CUDA runtime error; see alacgpu_last_error */
    ALACGPU_ERR_OUT_OF_MEMORY = -4,
    ALACGPU_ERR_UNSUPPORTED = -5,     /* sample size not 16/24 ("FIXME: unimplemented sample size", AlacFile.cs:574,715), channels not 1/2 */
    ALACGPU_ERR_CAPACITY = -6,        /* caller buffer too small */
    ALACGPU_ERR_STATE = -7,           /* call order (e.g. read before prepare) */
    ALACGPU_ERR_RANGE = -8            /* track / frame index out of range */
};

/* ---- per-frame status (frame_status[] of alacgpu_decode_all) ------------ */
/* A frame the decoder cannot finish yields zero PCM of its nominal size and a
 * non-zero status; the reference throws or returns stale data there. */
enum {
    ALACGPU_FRAME_OK = 0,
    ALACGPU_FRAME_BAD_TAG = 1,          /* element tag not 0/1 (AlacFile.cs:437,577)          */
    ALACGPU_FRAME_PRED_TYPE = 2,        /* prediction type != 0 (AlacFile.cs:488-496,650,660) */
    ALACGPU_FRAME_TOO_MANY_SAMPLES = 3, /* N > 16384 or PCM > 65536 B (AlacFile.cs:28; AlacContext.cs:218) */
    ALACGPU_FRAME_OVERRUN = 4,          /* bits consumed past the stsz length                 */
    ALACGPU_FRAME_BAD_RSS = 5,          /* sampleSize - 8*wastedBytes (+1) < 1                */
    ALACGPU_FRAME_HISTORY = 6,          /* Rice history wrapped negative (AlacFile.cs:229)    */
    ALACGPU_FRAME_RUN_OVERFLOW = 7,     /* zero run past the 16384-int scratch (AlacFile.cs:240-243) */
    ALACGPU_FRAME_ORDER0_LONG = 8,      /* order 0 with N > 4096 (AlacFile.cs:264-265)        */
    ALACGPU_FRAME_INTERNAL = 9          /* decoder fault (fused-kernel hand-off timed out); never expected */
};

typedef struct alacgpu_ctx alacgpu_ctx;

/* ---- options ------------------------------------------------------------- */
/* Launch shapes (every combination yields the same bytes).  Small batches (fewer frames than the GPU has
 * lanes): one fused entropy + LPC launch per chunk (LPC warps consume residuals while the entropy lanes still
 * produce them) followed by the un-mix / pack kernel; when the inputs are resident and pcm_dst is page-locked,
 * ONE launch with entropy, LPC and pack roles that writes the PCM straight into the destination (no device
 * PCM, no D2H stage).  Big batches (>= 650 k frames per device): the frame-lane kernels -- one lane per frame
 * and channel from bitstream to PCM, no residual / sample planes in HBM except a half-width copy of channel A
 * of stereo frames. */
#define ALACGPU_FLAG_KEEP_DEVICE_PCM 0x1u    /* keep decoded PCM resident in HBM after decode_all */
#define ALACGPU_FLAG_NO_FUSION 0x2u          /* entropy, LPC and pack as three kernels */
#define ALACGPU_FLAG_NO_PACK_FUSION 0x4u     /* never the launch with pack roles: un-mix / pack stays a separate kernel */
#define ALACGPU_FLAG_NO_ZERO_COPY 0x8u       /* never write PCM straight into a page-locked destination; always device PCM + D2H copies */
#define ALACGPU_FLAG_NO_QUAD_LPC 0x10u       /* one lane per stream for every LPC stream (no multi-lane path for the tail-critical ones) */
#define ALACGPU_FLAG_FORCE_PACK_FUSION 0x20u /* the launch with pack roles even when the PCM stays in HBM (tests / A-B runs) */
#define ALACGPU_FLAG_NO_FRAME_LANES 0x40u    /* never the frame-lane kernels (one lane per frame and channel from bitstream to PCM), which big batches use by default */
#define ALACGPU_FLAG_FORCE_FRAME_LANES 0x80u /* the frame-lane kernels for every chunk, however small (tests / A-B runs) */

typedef struct alacgpu_opts {
    uint32_t struct_size;      /* sizeof(alacgpu_opts), for forward compatibility       */
    uint32_t flags;            /* ALACGPU_FLAG_*                                         */
    uint32_t chunk_frames;     /* frames per pipeline chunk; 0 = automatic: resident inputs, as many frames as the
                                  launch shape takes (32768, frame-lane kernels 262144); inputs streamed in from
                                  the host, 1/16 of the device's frames (>= 256) so copies and kernels overlap */
    uint32_t entropy_lanes;    /* frames per entropy warp: 0 = auto, or 8 / 16 / 32      */
    uint32_t reserved[4];
} alacgpu_opts;

/* The 'alac' magic-cookie fields AlacFile.SetInfo keeps (AlacFile.cs:72-92)
 * plus the container channel count the AlacFile constructor receives
 * (AlacContext.cs:54; QTMovieT.cs:508-513). */
typedef struct alacgpu_track_cfg {
    int32_t sample_size;            /* cookie byte 29: 16 or 24                  */
    int32_t num_channels;           /* cookie byte 33: container channels, 1 or 2 */
    int32_t max_samples_per_frame;  /* cookie bytes 24..27 (4096 typical)        */
    int32_t rice_history_mult;      /* cookie byte 30 (40)                       */
    int32_t rice_initial_history;   /* cookie byte 31 (10)                       */
    int32_t rice_kmodifier;         /* cookie byte 32 (14); 0..31 accepted        */
    int32_t sample_rate;            /* cookie bytes 44..47 (informational)       */
} alacgpu_track_cfg;

/* Device-side timings of the last alacgpu_decode_all, from CUDA events
 * recorded on the streams the kernels were launched on. */
typedef struct alacgpu_timing {
    float index_ms;        /* K0 header pre-pass, summed over chunks             */
    float entropy_ms;      /* K1, summed over chunks (fused launch: every stage fused into it) */
    float lpc_ms;          /* K2 (0 when fused into the entropy launch)          */
    float stereo_ms;       /* K3 (fully fused launch: only the failed-frame fix-up) */
    float kernels_ms;      /* device pipeline span: first launch -> last kernel done (includes waits for H2D when streaming) */
    float h2d_ms;          /* mdat staging copies: first copy issued -> last done */
    float d2h_ms;          /* PCM copies to the caller's buffer                  */
    float total_ms;        /* decode_all wall clock (host)                       */
    uint32_t kernel_launches;  /* kernels launched by the last prepare+decode_all */
    uint32_t chunks;
    uint64_t compressed_bytes; /* sum of stsz over all frames                    */
    uint64_t pcm_bytes;        /* PCM bytes produced                             */
    uint64_t samples;          /* channel values produced (sample-frames x container channels) */
    uint32_t internal_retries; /* 1 if the batch was decoded a second time with unfused kernels because a fused
                                  launch flagged ALACGPU_FRAME_INTERNAL (never expected; see alacgpu_decode_all) */
    uint32_t reserved;
} alacgpu_timing;

/* ---- context ------------------------------------------------------------- */
/* device_ids: CUDA ordinals (NULL / n<=0 = device 0 only).  With n > 1 the
 * global frame list is partitioned into n contiguous ranges balanced by
 * compressed bytes and each device decodes its own range; there is no
 * collective (frames are independent, AlacFile.cs:430-435). */
ALACGPU_API int32_t alacgpu_create(const int32_t *device_ids, int32_t n_devices,
                                   const alacgpu_opts *opts, alacgpu_ctx **out);
ALACGPU_API int32_t alacgpu_destroy(alacgpu_ctx *ctx);

/* ---- tracks --------------------------------------------------------------- */
/* Stage one track.  `mdat` points at the bytes that hold the frames (host
 * memory; pinned memory from alacgpu_host_alloc is copied without an
 * intermediate bounce); frame i occupies frame_sizes[i] bytes starting at
 * first_frame_offset + sum(frame_sizes[0..i)) -- the reference's sequential
 * addressing (AlacContext.cs:194-195).  Frames that extend past mdat_len are
 * truncated (short read, MyStream.cs:47-52).  Only the frame headers are
 * looked at during this call; the bytes are copied to HBM by alacgpu_prepare or
 * by the first alacgpu_decode_all after the add, so `mdat` must stay valid until
 * that call returns (frame_sizes is copied immediately).  Once that call has
 * returned the library never reads `mdat` again: tracks added later are staged on
 * their own and the earlier ones stay resident in HBM (single-device contexts; a
 * multi-device context refuses new tracks after staging with ALACGPU_ERR_STATE,
 * because its partition would have to move frames -- alacgpu_clear_tracks first). */
ALACGPU_API int32_t alacgpu_add_track(alacgpu_ctx *ctx, const alacgpu_track_cfg *cfg,
                                      const uint8_t *mdat, uint64_t mdat_len,
                                      uint64_t first_frame_offset,
                                      const uint32_t *frame_sizes, uint32_t n_frames,
                                      int32_t *track_id);
/* Same, for containers whose frames are NOT back to back: frame i occupies
 * frame_sizes[i] bytes at frame_offsets[i] of `file` (the chunk-offset addressing
 * AlacContext.SetPosition derives from stco x stsc x stsz, AlacContext.cs:262-295;
 * co64 offsets fit too).  Used by the tolerant demuxer. */
ALACGPU_API int32_t alacgpu_add_track_offsets(alacgpu_ctx *ctx, const alacgpu_track_cfg *cfg,
                                              const uint8_t *file, uint64_t file_len,
                                              const uint64_t *frame_offsets, const uint32_t *frame_sizes,
                                              uint32_t n_frames, int32_t *track_id);
/* Forget all tracks but keep device / pinned allocations for reuse. */
ALACGPU_API int32_t alacgpu_clear_tracks(alacgpu_ctx *ctx);

/* ---- decode --------------------------------------------------------------- */
/* Size of the buffer alacgpu_decode_all needs for the tracks added so far.
 * Known as soon as the tracks are added: the host reads the first seven bytes
 * of every frame (element tag, hassize flag, 32-bit sample count --
 * AlacFile.cs:435-453) while building its frame index.  Track t's PCM starts at
 * a 256-byte aligned offset; the total includes those gaps (zero bytes). */
ALACGPU_API int32_t alacgpu_total_pcm_bytes(alacgpu_ctx *ctx, uint64_t *total_pcm_bytes);

/* OPTIONAL staging step: copy every added track's mdat into HBM and run the
 * header pre-pass (K0), so that a following alacgpu_decode_all starts from
 * resident inputs.  Without it alacgpu_decode_all streams the mdat in itself,
 * chunk by chunk, overlapped with the kernels and the PCM copies. */
ALACGPU_API int32_t alacgpu_prepare(alacgpu_ctx *ctx, uint64_t *total_pcm_bytes);

/* Mark the device-side frame index stale: the next alacgpu_decode_all re-runs
 * the header pre-pass (K0) over the bytes already resident in HBM inside its
 * own pipeline, before the decode kernels; no host->device copy.  (Stages the
 * tracks first if they are not resident yet.)  Used to time the whole kernel
 * path with resident inputs. */
ALACGPU_API int32_t alacgpu_reindex(alacgpu_ctx *ctx);

/* Decode every frame of every track.  pcm_dst: caller-owned HOST buffer of
 * `cap` bytes, or NULL to leave the PCM device-resident (see
 * alacgpu_device_pcm).  track_pcm_off / track_pcm_len (n_tracks entries each,
 * optional) receive where each track's interleaved little-endian PCM landed;
 * frame_status (one int32 per frame, track-major, optional) receives the
 * ALACGPU_FRAME_* codes.  If the tracks are not resident yet (no
 * alacgpu_prepare) the call runs the whole pipeline: H2D of each chunk's mdat
 * -> header pre-pass -> decode kernels -> D2H of the chunk's PCM, with up to 16
 * chunks in flight on separate CUDA streams (3 for the frame-lane kernels, whose
 * chunks fill the machine on their own).  A frame the kernels flag
 * ALACGPU_FRAME_INTERNAL (a fused launch's hand-off timed out) is never returned
 * as zeros: the call decodes the batch again with unfused kernels first
 * (alacgpu_timing.internal_retries). */
ALACGPU_API int32_t alacgpu_decode_all(alacgpu_ctx *ctx, uint8_t *pcm_dst, uint64_t cap,
                                       uint64_t *track_pcm_off, uint64_t *track_pcm_len,
                                       int32_t *frame_status);

/* One AlacContext.Read: frame `frame_idx` of `track` as PCM bytes.  The first
 * call decodes the batch (device-resident) and later calls are served from
 * that cache.  *bytes_out = 0 and ALACGPU_OK past the last frame, like the
 * reference's `return 0` (AlacContext.cs:182-186). */
ALACGPU_API int32_t alacgpu_read_frame(alacgpu_ctx *ctx, int32_t track, uint32_t frame_idx,
                                       uint8_t *dst, uint32_t cap, uint32_t *bytes_out);

/* ---- info ----------------------------------------------------------------- */
ALACGPU_API int32_t alacgpu_track_count(alacgpu_ctx *ctx, int32_t *n_tracks);
ALACGPU_API int32_t alacgpu_frame_count(alacgpu_ctx *ctx, int32_t track, uint32_t *n_frames);
/* Samples (per channel) frame `frame_idx` decodes to, from its header
 * (AlacFile.cs:447-453); needs alacgpu_prepare. */
ALACGPU_API int32_t alacgpu_frame_samples(alacgpu_ctx *ctx, int32_t track, uint32_t frame_idx,
                                          uint32_t *n_samples);
ALACGPU_API int32_t alacgpu_track_pcm_bytes(alacgpu_ctx *ctx, int32_t track, uint64_t *off, uint64_t *len);
ALACGPU_API int32_t alacgpu_frame_status(alacgpu_ctx *ctx, int32_t track, uint32_t frame_idx, int32_t *status);
ALACGPU_API int32_t alacgpu_get_timing(alacgpu_ctx *ctx, alacgpu_timing *out);

/* Device-resident PCM of the last decode_all(NULL, ...) on device `dev_slot`
 * (index into the create-time device list): device pointer + byte length of
 * that device's shard and the shard's offset in the global PCM layout. */
ALACGPU_API int32_t alacgpu_device_pcm(alacgpu_ctx *ctx, int32_t dev_slot, void **dptr,
                                       uint64_t *shard_off, uint64_t *shard_len);
/* 64-bit position-dependent checksum of PCM bytes [off, off+len) computed ON
 * THE DEVICE over the resident PCM (sum over 8-byte little-endian words w_j of
 * w_j * (2*j+1) mod 2^64, j = index of the word in the global layout; tail
 * bytes zero-extended).  Lets full-size runs be checked without a D2H copy;
 * tests compare against the same formula over the oracle's PCM. */
ALACGPU_API int32_t alacgpu_pcm_checksum(alacgpu_ctx *ctx, uint64_t off, uint64_t len, uint64_t *sum);

/* ---- pinned host memory ---------------------------------------------------- */
ALACGPU_API int32_t alacgpu_host_alloc(uint64_t bytes, void **ptr);
ALACGPU_API int32_t alacgpu_host_free(void *ptr);

/* ---- multi-GPU partition plan (pure host arithmetic; no device touched) ---- */
/* Split n_frames frames (sizes in bytes) into n_parts contiguous ranges
 * balanced by compressed bytes; cut[0]=0 <= cut[1] <= ... <= cut[n_parts]=n_frames. */
ALACGPU_API int32_t alacgpu_plan_partition(const uint32_t *frame_sizes, uint64_t n_frames,
                                           int32_t n_parts, uint64_t *cut);

/* ---- diagnostics ----------------------------------------------------------- */
ALACGPU_API const char *alacgpu_strerror(int32_t status);
ALACGPU_API const char *alacgpu_last_error(alacgpu_ctx *ctx);   /* detail of the last failure (may be "") */
ALACGPU_API int32_t alacgpu_abi_version(void);
ALACGPU_API int32_t alacgpu_device_count(int32_t *n);

#ifdef __cplusplus
}
#endif
#endif /* ALACGPU_H */
