/*
 * alac_oracle.c -- CPU restatement of teekay/ALAC.NET's ALAC frame decoder.
 *
 * TEST INFRASTRUCTURE ONLY (see alac_oracle.h).  PARITY UNPINNED by the
 * reference's own tests (it has none); pinned by cross-restatement and
 * encoder round trips -- see DESIGN.md.
 *
 * C# semantics restated here:
 *   - int arithmetic wraps mod 2^32  -> computed in uint32_t and cast back;
 *   - shift counts are masked & 31; >> on int is arithmetic;
 *   - / truncates toward zero.
 * Citations: ALACDecoder/AlacFile.cs, ALACDecoder/AlacContext.cs.
 */
#include "alac_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---- C# int helpers ---------------------------------------------------- */
static inline int32_t w_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t w_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t w_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int32_t w_shl(int32_t a, int32_t n) { return (int32_t)((uint32_t)a << (n & 31)); }
static inline int32_t w_sar(int32_t a, int32_t n)
{
    /* arithmetic shift right without relying on implementation-defined >> */
    n &= 31;
    if (a >= 0) return (int32_t)((uint32_t)a >> n);
    return (int32_t)~((~(uint32_t)a) >> n);
}
static inline int32_t sign_extend(int32_t v, int32_t bits)
{
    /* (v << (32-bits)) >> (32-bits), AlacFile.cs:278-279,289-290,309-310 */
    int32_t mv = 32 - bits;
    return w_sar(w_shl(v, mv), mv);
}

/* ---- bit reader: AlacFile.cs:101-152 ----------------------------------- */
typedef struct {
    const uint8_t *buf;
    int64_t len;     /* bytes in this frame (stsz) */
    int64_t idx;     /* _ibIdx */
    int32_t acc;     /* _inputBufferBitaccumulator, 0..7 */
} bitreader;

/* The reference reads from an 80 KiB scratch that still holds older frames
 * (AlacContext.cs:64,195); bytes past the frame are defined as 0 here and a
 * frame whose consumed bits run past its stsz length is reported OVERRUN. */
static inline int32_t br_byte(const bitreader *b, int64_t i)
{
    return (i >= 0 && i < b->len) ? (int32_t)b->buf[i] : 0;
}

/* Readbits16, AlacFile.cs:101-118: 0..16 bits out of a 24-bit look-ahead. */
static int32_t br_read16(bitreader *b, int32_t bits)
{
    int32_t w24 = (br_byte(b, b->idx) << 16) | (br_byte(b, b->idx + 1) << 8) | br_byte(b, b->idx + 2);
    int32_t result = w_sar(w_shl(w24, b->acc) & 0x00ffffff, 24 - bits);
    int32_t na = b->acc + bits;
    b->idx += na >> 3;
    b->acc = na & 7;
    return result;
}

/* Readbits, AlacFile.cs:125-129: >16 bits = high 16 first, then the rest. */
static int32_t br_read(bitreader *b, int32_t n)
{
    if (n <= 16) return br_read16(b, n);
    int32_t lo_bits = n - 16;
    int32_t hi = w_shl(br_read16(b, 16), lo_bits);
    return hi | br_read16(b, lo_bits);
}

/* Readbit, AlacFile.cs:135-143 */
static int32_t br_read1(bitreader *b)
{
    int32_t r = ((br_byte(b, b->idx) << b->acc) >> 7) & 1;
    int32_t na = b->acc + 1;
    b->idx += na / 8;
    b->acc = na % 8;
    return r;
}

/* Unreadbits, AlacFile.cs:145-152 (only ever called with 1) */
static void br_unread(bitreader *b, int32_t bits)
{
    int32_t na = b->acc - bits;
    b->idx += (na >> 3);       /* na >= -1 here; -1 >> 3 == -1 in C# and gcc */
    b->acc = na & 7;
}

static inline int64_t br_bitpos(const bitreader *b) { return b->idx * 8 + b->acc; }

/* ---- CountLeadingZeros: AlacFile.cs:154-191 ----------------------------- */
/* The byte-by-byte search returns the true clz for non-zero input and
 * 32 + 8 = 40 for input 0 (the loop falls through, AlacFile.cs:190). */
int alac_oracle_clz(int32_t input)
{
    uint32_t u = (uint32_t)input;
    if (u == 0) return 40;
    int n = 0;
    while (!(u & 0x80000000u)) { u <<= 1; n++; }
    return n;
}

/* ---- EntropyDecodeValue: AlacFile.cs:193-212 ---------------------------- */
static int32_t entropy_decode_value(bitreader *b, int32_t read_sample_size, int32_t k, int32_t kmask)
{
    int32_t x = 0;                           /* :196 count 1-bits, stop after a 0 or after 9 ones */
    while (x <= 8 && br_read1(b) != 0) x++;
    if (x > 8) {                             /* :198-202 escape: raw value */
        uint32_t m = 0xffffffffu >> ((32 - read_sample_size) & 31);
        return br_read(b, read_sample_size) & (int32_t)m;
    }
    if (k == 1) return x;                    /* :203 */
    int32_t extra = br_read(b, k);           /* :205 */
    int32_t v = w_mul(x, (w_sub(w_shl(1, k), 1)) & kmask);   /* :206 */
    if (extra > 1)
        v = w_add(v, extra - 1);             /* :207-208 */
    else
        br_unread(b, 1);                     /* :210 */
    return v;
}

/* ---- EntropyRiceDecode: AlacFile.cs:214-252 ----------------------------- */
static int entropy_rice_decode(bitreader *b, int32_t *out, int32_t n, int32_t rss,
                               int32_t initial_history, int32_t kmod, int32_t mult, int32_t kmask)
{
    int32_t history = initial_history;
    int32_t count = 0;
    int32_t sign_mod = 0;
    while (count < n) {
        int32_t t = 31 - kmod - alac_oracle_clz(w_add(w_sar(history, 9), 3));   /* :221 */
        int32_t k = t < 0 ? t + kmod : kmod;                                      /* :222 */
        int32_t dv = w_add(entropy_decode_value(b, rss, k, (int32_t)0xffffffffu), sign_mod); /* :224 */
        if (br_bitpos(b) > b->len * 8) return ALAC_ORACLE_OVERRUN;   /* policy: symbol must lie inside the frame */
        int32_t half = w_add(dv, 1) / 2;                                         /* :225 */
        out[count] = (dv & 1) ? w_mul(half, -1) : half;                          /* :226 */
        sign_mod = 0;
        history = dv > 0xFFFF ? 0xFFFF
                : w_sub(w_add(history, w_mul(dv, mult)), w_sar(w_mul(history, mult), 9)); /* :229 */
        if (history < 0) return ALAC_ORACLE_HISTORY;   /* policy: wrapped history is unsupported */
        if (history < 128 && count + 1 < n) {                                    /* :231 */
            sign_mod = 1;
            k = alac_oracle_clz(history) + ((history + 16) / 64) - 24;           /* :234 */
            int32_t block = entropy_decode_value(b, 16, k, kmask);               /* :236 */
            if (br_bitpos(b) > b->len * 8) return ALAC_ORACLE_OVERRUN;
            if (block > 0) {
                if ((int64_t)count + 1 + block > ALAC_ORACLE_BUFFER_SIZE)
                    return ALAC_ORACLE_RUN_OVERFLOW;   /* reference: IndexOutOfRangeException */
                for (int32_t j = 0; j < block; j++) out[count + 1 + j] = 0;      /* :240-243 */
                count += block;
            }
            if (block > 0xFFFF) sign_mod = 0;                                    /* :246 */
            history = 0;
        }
        count++;
    }
    return ALAC_ORACLE_OK;
}

/* ---- PredictorDecompressFirAdapt: AlacFile.cs:256-336 ------------------- */
/* In place on buf (the reference returns errorBuffer itself, :260). */
static int predictor_decompress(int32_t *buf, int32_t n, int32_t rss, int32_t *coef,
                                int32_t order, int32_t quant)
{
    if (order == 0) {                                    /* :261-267 identity */
        if (n <= 1) return ALAC_ORACLE_OK;
        if ((int64_t)1 + (int64_t)(n - 1) * 4 > ALAC_ORACLE_BUFFER_SIZE)
            return ALAC_ORACLE_ORDER0_LONG;              /* Array.Copy would throw */
        return ALAC_ORACLE_OK;
    }
    if (order == 0x1f) {                                 /* :268-282 first-order delta */
        for (int32_t i = 0; i + 1 < n; i++)
            buf[i + 1] = sign_extend(w_add(buf[i], buf[i + 1]), rss);
        return ALAC_ORACLE_OK;
    }
    for (int32_t i = 0; i < order; i++)                  /* :284-293 warm-up */
        buf[i + 1] = sign_extend(w_add(buf[i], buf[i + 1]), rss);
    /* (the reference's scratch is 16384 ints, so warm-up past n is harmless
     * there; callers here always supply ALAC_ORACLE_BUFFER_SIZE ints) */
    int32_t base = 0;                                    /* bufferOutIdx */
    for (int32_t i = order + 1; i < n; i++) {            /* :297 */
        int32_t sum = 0;
        int32_t err = buf[i];                            /* :300 */
        for (int32_t j = 0; j < order; j++)              /* :301-305 */
            sum = w_add(sum, w_mul(w_sub(buf[base + order - j], buf[base]), coef[j]));
        int32_t outval = w_add(w_shl(1, quant - 1), sum);   /* :306 */
        outval = w_sar(outval, quant);                      /* :307 */
        outval = w_add(w_add(outval, buf[base]), err);      /* :308 */
        outval = sign_extend(outval, rss);                  /* :309-310 */
        buf[base + order + 1] = outval;                     /* :311 */
        if (err > 0) {                                      /* :312-331, conditionToUse = v > 0 */
            for (int32_t p = order - 1; p >= 0 && err > 0; p--) {
                int32_t val = w_sub(buf[base], buf[base + order - p]);
                int32_t sign = val < 0 ? -1 : (val > 0 ? 1 : 0);
                coef[p] = w_sub(coef[p], sign);
                val = w_mul(val, sign);
                err = w_sub(err, w_mul(w_sar(val, quant), order - p));
            }
        } else if (err < 0) {                               /* conditionToUse = v < 0, sign negated */
            for (int32_t p = order - 1; p >= 0 && err < 0; p--) {
                int32_t val = w_sub(buf[base], buf[base + order - p]);
                int32_t sign = w_mul(val < 0 ? -1 : (val > 0 ? 1 : 0), -1);
                coef[p] = w_sub(coef[p], sign);
                val = w_mul(val, sign);
                err = w_sub(err, w_mul(w_sar(val, quant), order - p));
            }
        }
        base++;
    }
    return ALAC_ORACLE_OK;
}

/* ---- SetInfo: AlacFile.cs:63-93 ----------------------------------------- */
int alac_oracle_set_info(const uint8_t *cd, size_t len, alac_oracle_cfg *cfg)
{
    if (len < 48) return -1;
    size_t p = 24;                                       /* six 4-byte fields skipped, :66-71 */
    cfg->max_samples_per_frame = (int32_t)(((uint32_t)cd[p] << 24) + (cd[p + 1] << 16) + (cd[p + 2] << 8) + cd[p + 3]);
    p += 4;
    p += 1;                                              /* _setinfo_7A */
    cfg->sample_size = cd[p++];                          /* :76 */
    cfg->rice_history_mult = cd[p++];                    /* :78 */
    cfg->rice_initial_history = cd[p++];                 /* :80 */
    cfg->rice_kmodifier = cd[p++];                       /* :82 */
    cfg->num_channels = cd[p++];                         /* :84 == QTMovieT.cs:510-511 */
    return 0;
}

/* ---- DecodeFrame: AlacFile.cs:428-719 ----------------------------------- */
typedef struct {
    int32_t a[ALAC_ORACLE_BUFFER_SIZE];
    int32_t b[ALAC_ORACLE_BUFFER_SIZE];
    int32_t sa[ALAC_ORACLE_BUFFER_SIZE];
    int32_t sb[ALAC_ORACLE_BUFFER_SIZE];
} scratch_t;

static int decode_frame_inner(const alac_oracle_cfg *cfg, const uint8_t *in, size_t in_len,
                              int32_t *out, size_t out_ints, int32_t *outputsize_p,
                              scratch_t *s, alac_oracle_stages *st)
{
    const int32_t ss = cfg->sample_size;
    const int32_t nch = cfg->num_channels;                       /* _numchannels */
    const int32_t bytespersample = (ss / 8) * nch;               /* AlacFile.cs:19 */
    int32_t n = cfg->max_samples_per_frame;                      /* :430 */
    bitreader br = { in, (int64_t)in_len, 0, 0 };
    int status = ALAC_ORACLE_OK;
    if (ss != 16 && ss != 24) return -2;                         /* :570-574, :713-715 throw */

    int32_t tag = br_read(&br, 3);                               /* :435 */
    *outputsize_p = w_mul(n, bytespersample);                    /* :436 */
    if (tag != 0 && tag != 1) return ALAC_ORACLE_BAD_TAG;        /* :437,:577 -> :718 */
    const int stereo = (tag == 1);

    br_read(&br, 4);                                             /* :442 / :584 */
    br_read(&br, 12);                                            /* :443 / :585 */
    int32_t hassize = br_read(&br, 1);
    int32_t ub = br_read(&br, 2);
    int32_t escape = br_read(&br, 1);
    if (hassize != 0) {                                          /* :447-453 / :589-595 */
        n = br_read(&br, 32);
        *outputsize_p = w_mul(n, bytespersample);
    }
    if (n < 0 || n > ALAC_ORACLE_BUFFER_SIZE ||
        (int64_t)n * bytespersample > ALAC_ORACLE_MAX_PCM_BYTES) {
        *outputsize_p = 0;
        return ALAC_ORACLE_TOO_MANY_SAMPLES;
    }
    int32_t rss = ss - ub * 8 + (stereo ? 1 : 0);                /* :454 / :596 */
    int32_t mix_shift = 0, mix_weight = 0;
    int32_t pred_type[2] = {0, 0}, quant[2] = {0, 0}, rice_mod[2] = {0, 0}, order[2] = {0, 0};
    int32_t coef[2][32];
    memset(coef, 0, sizeof coef);
    const int ech = stereo ? 2 : 1;

    if (escape == 0) {
        if (rss < 1) return ALAC_ORACLE_BAD_RSS;
        mix_shift = br_read(&br, 8);                             /* :459 / :599 */
        mix_weight = br_read(&br, 8);                            /* :460 / :600 */
        for (int c = 0; c < ech; c++) {                          /* :461-475 / :602-632 */
            pred_type[c] = br_read(&br, 4);
            quant[c] = br_read(&br, 4);
            rice_mod[c] = br_read(&br, 3);
            order[c] = br_read(&br, 5);
            for (int32_t i = 0; i < order[c]; i++) {
                int32_t t = br_read(&br, 16);
                if (t > 32767) t -= 65536;
                coef[c][i] = t;
            }
        }
        if (ub != 0) {                                           /* :476-482 / :634-641 */
            for (int32_t i = 0; i < n; i++) {
                s->sa[i] = br_read(&br, ub * 8);
                if (stereo) s->sb[i] = br_read(&br, ub * 8);
            }
        }
        if (st) {
            st->mix_shift = mix_shift; st->mix_weight = mix_weight;
            for (int c = 0; c < 2; c++) {
                st->pred_type[c] = pred_type[c]; st->quant[c] = quant[c];
                st->rice_mod[c] = rice_mod[c]; st->order[c] = order[c];
                memcpy(st->coef[c], coef[c], sizeof coef[c]);
            }
        }
        /* Status policy for malformed frames: header-level faults first,
         * then entropy faults in channel order, then overrun. */
        for (int c = 0; c < ech; c++)
            if (pred_type[c] != 0) return ALAC_ORACLE_PRED_TYPE; /* :488-496 stale / :650,:660 throw */
        for (int c = 0; c < ech; c++)
            if (order[c] == 0 && n > 1 && (int64_t)1 + (int64_t)(n - 1) * 4 > ALAC_ORACLE_BUFFER_SIZE)
                return ALAC_ORACLE_ORDER0_LONG;                  /* Array.Copy length, :264-265 */
        if (br_bitpos(&br) > (int64_t)in_len * 8) return ALAC_ORACLE_OVERRUN;
        int32_t kmask = w_sub(w_shl(1, cfg->rice_kmodifier), 1);
        for (int c = 0; c < ech; c++) {                          /* :483-487 / :643-661 */
            int32_t *buf = c == 0 ? s->a : s->b;
            int32_t mult = w_mul(rice_mod[c], cfg->rice_history_mult / 4);
            status = entropy_rice_decode(&br, buf, n, rss, cfg->rice_initial_history,
                                         cfg->rice_kmodifier, mult, kmask);
            if (status) return status;
            if (st) memcpy(st->residual[c], buf, sizeof(int32_t) * (size_t)n);
            status = predictor_decompress(buf, n, rss, coef[c], order[c], quant[c]);
            if (status) return status;
        }
    } else {                                                     /* :498-526 / :663-700 */
        if (ss <= 16) {
            for (int32_t i = 0; i < n; i++) {
                s->a[i] = sign_extend(br_read(&br, ss), ss);
                if (stereo) s->b[i] = sign_extend(br_read(&br, ss), ss);
            }
        } else {
            const int32_t m = 1 << 23;
            for (int32_t i = 0; i < n; i++) {
                for (int c = 0; c < ech; c++) {
                    int32_t v = w_shl(br_read(&br, 16), ss - 16);
                    v |= br_read(&br, ss - 16);
                    v = ((v & 0xffffff) ^ m) - m;
                    (c == 0 ? s->a : s->b)[i] = v;
                }
            }
        }
        ub = 0; mix_shift = 0; mix_weight = 0;                   /* :525 / :697-699 */
    }
    if (br_bitpos(&br) > (int64_t)in_len * 8) return ALAC_ORACLE_OVERRUN;
    if (st) {
        st->element_channels = ech; st->n = n; st->ub = ub; st->escape = escape;
        st->bits_consumed = br_bitpos(&br);
        memcpy(st->predicted[0], s->a, sizeof(int32_t) * (size_t)n);
        if (stereo) memcpy(st->predicted[1], s->b, sizeof(int32_t) * (size_t)n);
        if (ub) {
            memcpy(st->shift[0], s->sa, sizeof(int32_t) * (size_t)n);
            if (stereo) memcpy(st->shift[1], s->sb, sizeof(int32_t) * (size_t)n);
        }
    }

    /* ---- output packing ------------------------------------------------ */
    /* Ints are written exactly where the reference writes them; entries at
     * or past outputsize's worth are discarded like the reference's caller
     * does (AlacContext.cs:169 copies bytesRead bytes only). */
    const size_t ints_needed = ss == 16 ? (size_t)n * nch + 2 : (size_t)n * nch * 3 + 6;
    if (out_ints < ints_needed) return -1;
    const int32_t sh = ub * 8;
    const int32_t mask = (int32_t)~(0xFFFFFFFFu << (sh & 31));   /* :383,:407,:552 */
    for (int32_t i = 0; i < n; i++) {
        int32_t left, right;
        if (!stereo) {                                           /* :527-575 */
            left = s->a[i];
            right = 0;
            if (ss == 24 && ub != 0)
                left = w_shl(left, sh) | (s->sa[i] & mask);      /* :549-554 */
        } else {                                                 /* :338-421 */
            if (mix_weight != 0) {
                int32_t mid = s->a[i], diff = s->b[i];
                right = w_sub(mid, w_sar(w_mul(diff, mix_weight), mix_shift));
                left = w_add(right, diff);
            } else {
                left = s->a[i];
                right = s->b[i];
            }
            if (ss == 24 && ub != 0) {                           /* 16-bit ignores wasted bytes */
                left = w_shl(left, sh) | (s->sa[i] & mask);
                right = w_shl(right, sh) | (s->sb[i] & mask);
            }
        }
        if (ss == 16) {
            /* :353-354 / :534,:540 -- index i*nch and i*nch+1; with one
             * container channel the second store is overwritten by i+1. */
            out[(size_t)i * nch] = left;
            out[(size_t)i * nch + 1] = right;
        } else {
            size_t o = (size_t)i * nch * 3;                      /* :390-395 / :555-565 */
            out[o] = left & 0xFF; out[o + 1] = w_sar(left, 8) & 0xFF; out[o + 2] = w_sar(left, 16) & 0xFF;
            out[o + 3] = right & 0xFF; out[o + 4] = w_sar(right, 8) & 0xFF; out[o + 5] = w_sar(right, 16) & 0xFF;
        }
    }
    return ALAC_ORACLE_OK;
}

int alac_oracle_decode_frame(const alac_oracle_cfg *cfg, const uint8_t *in, size_t in_len,
                             int32_t *outbuffer, size_t outbuffer_ints,
                             int *status, alac_oracle_stages *stages)
{
    scratch_t *s = (scratch_t *)calloc(1, sizeof(scratch_t));
    int32_t outputsize = 0;
    if (stages) memset(stages, 0, sizeof *stages);
    int st = decode_frame_inner(cfg, in, in_len, outbuffer, outbuffer_ints, &outputsize, s, stages);
    free(s);
    if (st < 0) { if (status) *status = st; return 0; }
    if (st != ALAC_ORACLE_OK) {
        /* policy: a frame the decoder cannot finish yields zeros */
        int bps = cfg->sample_size / 8;
        size_t ints = cfg->sample_size == 16 ? (size_t)outputsize / 2 : (size_t)outputsize;
        (void)bps;
        if (ints > outbuffer_ints) ints = outbuffer_ints;
        memset(outbuffer, 0, ints * sizeof(int32_t));
    }
    if (status) *status = st;
    return outputsize;
}

/* ---- AlacContext.Read / FormatSamples: AlacContext.cs:163-172, 214-256 --- */
int alac_oracle_read_frame(const alac_oracle_cfg *cfg, const uint8_t *in, size_t in_len,
                           uint8_t *pcm, size_t pcm_cap, int *status)
{
    static const size_t FORMAT_INTS = 1024 * 80 + 8;             /* AlacContext.cs:65 */
    int32_t *fmt = (int32_t *)calloc(FORMAT_INTS, sizeof(int32_t));
    int st = 0;
    int bytes = alac_oracle_decode_frame(cfg, in, in_len, fmt, FORMAT_INTS, &st, NULL);
    if (status) *status = st;
    if (st < 0 || (size_t)(bytes < 0 ? 0 : bytes) > pcm_cap) { free(fmt); if (status) *status = -1; return 0; }
    if (bytes <= 0) { free(fmt); return 0; }
    int bps = (cfg->sample_size + 7) / 8;                        /* GetBytesPerSample, :101 */
    if (bps == 2) {                                              /* :231-242 */
        int c = 0, c2 = 0;
        for (int rem = bytes; rem > 0; rem -= 2) {
            int32_t t = fmt[c2++];
            pcm[c++] = (uint8_t)t;
            pcm[c++] = (uint8_t)((uint32_t)t >> 8);
        }
    } else if (bps == 3) {                                       /* :244-252 */
        for (int i = 0; i < bytes; i++) pcm[i] = (uint8_t)fmt[i];
    } else {
        free(fmt);
        return 0;
    }
    free(fmt);
    return bytes;
}

/* ---- UnpackSamples loop: AlacContext.cs:179-204 -------------------------- */
int64_t alac_oracle_decode_track(const alac_oracle_cfg *cfg, const uint8_t *mdat, size_t mdat_len,
                                 const uint32_t *stsz, uint32_t n_frames,
                                 uint8_t *pcm, size_t pcm_cap,
                                 int32_t *frame_status, uint32_t *frame_bytes)
{
    size_t in_off = 0, out_off = 0;
    for (uint32_t f = 0; f < n_frames; f++) {
        size_t sz = stsz[f];
        size_t avail = in_off <= mdat_len ? mdat_len - in_off : 0;
        if (sz > avail) sz = avail;                              /* short read, MyStream.cs:47 */
        int st = 0;
        int got = alac_oracle_read_frame(cfg, mdat + (in_off <= mdat_len ? in_off : mdat_len), sz,
                                         pcm + out_off, pcm_cap - out_off, &st);
        if (st < 0) return -1;
        if (frame_status) frame_status[f] = st;
        if (frame_bytes) frame_bytes[f] = (uint32_t)got;
        in_off += stsz[f];                                       /* sequential addressing, :194-195 */
        out_off += (size_t)got;
    }
    return (int64_t)out_off;
}
