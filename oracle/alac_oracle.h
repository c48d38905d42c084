/*
 * alac_oracle.h -- CPU restatement of the teekay/ALAC.NET frame decoder.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle for libalacgpu: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may build, link or call it.  The product path (libalacgpu.so
 * and everything under alac/net_b200/) never references this directory.
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors
 *  and it is C#, which cannot be compiled or run in
 * this image.  The oracle is pinned instead by (1) a line-by-line reading of
 * ALACDecoder/AlacFile.cs and AlacContext.cs (every function cites the lines
 * it follows), (2) an independent Python model written from SURVEY.md
 * appendix A (pymodel/), and (3) round trips through a from-scratch encoder.
 *
 * All citations are relative to /root/reference/.
 */
#ifndef ALAC_ORACLE_H
#define ALAC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-frame status.  The numeric values are shared with include/alacgpu.h
 * (ALACGPU_FRAME_*) so parity tests can compare status words directly. */
enum {
    ALAC_ORACLE_OK = 0,
    ALAC_ORACLE_BAD_TAG = 1,        /* element tag not 0/1: AlacFile.cs:437,577 decode nothing        */
    ALAC_ORACLE_PRED_TYPE = 2,      /* prediction type != 0: AlacFile.cs:488-496 (stale), :650,660    */
    ALAC_ORACLE_TOO_MANY_SAMPLES = 3, /* N > 16384 or PCM > 65536 B: AlacFile.cs:28, AlacContext.cs:218 */
    ALAC_ORACLE_OVERRUN = 4,        /* bits consumed past the stsz frame length                        */
    ALAC_ORACLE_BAD_RSS = 5,        /* sampleSize - 8*ub (+1) < 1                                       */
    ALAC_ORACLE_HISTORY = 6,        /* rice history wrapped negative (AlacFile.cs:229 overflow)        */
    ALAC_ORACLE_RUN_OVERFLOW = 7,   /* zero run past the 16384-int scratch: AlacFile.cs:240-243        */
    ALAC_ORACLE_ORDER0_LONG = 8     /* order 0 with N > 4096: Array.Copy length, AlacFile.cs:264-265   */
};

#define ALAC_ORACLE_BUFFER_SIZE 16384   /* AlacFile.cs:28 */
#define ALAC_ORACLE_MAX_PCM_BYTES 65536 /* AlacContext.cs:218 */

/* What AlacFile.SetInfo keeps (AlacFile.cs:63-93) plus the container fields
 * AlacContext passes to the AlacFile constructor (AlacContext.cs:54). */
typedef struct alac_oracle_cfg {
    int32_t sample_size;            /* cookie byte 29 (16 or 24)                 */
    int32_t num_channels;           /* cookie byte 33: CONTAINER channel count   */
    int32_t max_samples_per_frame;  /* cookie bytes 24..27, big endian           */
    int32_t rice_history_mult;      /* cookie byte 30 (40 typical)               */
    int32_t rice_initial_history;   /* cookie byte 31 (10 typical)               */
    int32_t rice_kmodifier;         /* cookie byte 32 (14 typical)               */
} alac_oracle_cfg;

/* Stage intermediates of one frame, for per-kernel parity tests. */
typedef struct alac_oracle_stages {
    int32_t element_channels;       /* 1 or 2                                    */
    int32_t n;                      /* samples in this frame                     */
    int32_t ub;                     /* wasted bytes                              */
    int32_t escape;                 /* isnotcompressed                           */
    int32_t mix_shift, mix_weight;
    int32_t pred_type[2], quant[2], rice_mod[2], order[2];
    int32_t coef[2][32];
    int32_t residual[2][ALAC_ORACLE_BUFFER_SIZE];   /* EntropyRiceDecode output  */
    int32_t predicted[2][ALAC_ORACLE_BUFFER_SIZE];  /* predictor / raw output    */
    int32_t shift[2][ALAC_ORACLE_BUFFER_SIZE];      /* wasted-byte planes        */
    int64_t bits_consumed;
} alac_oracle_stages;

/* cookie -> cfg.  `codec_data` is DemuxResT.CodecData viewed as bytes (the
 * 'alac' atom starts at offset 12, QTMovieT.cs:487-490). */
int alac_oracle_set_info(const uint8_t *codec_data, size_t len, alac_oracle_cfg *cfg);

/* AlacFile.DecodeFrame (AlacFile.cs:428-719): frame bytes -> int[] exactly as
 * the reference fills `outbuffer` (16-bit: one int per sample; 24-bit: one
 * byte-valued int per output byte).  Returns the reference's return value
 * (`outputsize` in bytes); *status gets an ALAC_ORACLE_* code; on a non-OK
 * status the first `outputsize` bytes' worth of ints are zero (documented
 * policy, see DESIGN.md "malformed frames").  `stages` may be NULL. */
int alac_oracle_decode_frame(const alac_oracle_cfg *cfg, const uint8_t *in, size_t in_len,
                             int32_t *outbuffer, size_t outbuffer_ints,
                             int *status, alac_oracle_stages *stages);

/* AlacContext.Read (AlacContext.cs:163-172) for frame `in`: DecodeFrame +
 * FormatSamples (AlacContext.cs:214-256).  Writes little-endian interleaved
 * PCM to `pcm`; returns the byte count (0 on capacity failure). */
int alac_oracle_read_frame(const alac_oracle_cfg *cfg, const uint8_t *in, size_t in_len,
                           uint8_t *pcm, size_t pcm_cap, int *status);

/* Whole track as the reference pumps it (AlacContext.cs:179-204): frames are
 * addressed sequentially from `mdat`, frame i having stsz[i] bytes.  Writes
 * PCM contiguously; per-frame status/byte counts optional.  Returns total
 * bytes or -1 if pcm_cap is too small. */
int64_t alac_oracle_decode_track(const alac_oracle_cfg *cfg, const uint8_t *mdat, size_t mdat_len,
                                 const uint32_t *stsz, uint32_t n_frames,
                                 uint8_t *pcm, size_t pcm_cap,
                                 int32_t *frame_status, uint32_t *frame_bytes);

/* Exposed for unit tests of the quirks. */
int alac_oracle_clz(int32_t input);                       /* AlacFile.cs:170-191; clz(0)=40 */

#ifdef __cplusplus
}
#endif
#endif
