/*
 * alacenc.c -- from-scratch synthetic ALAC frame encoder (test/bench input
 * generator; there is no corpus offline and the reference has no encoder).
 *
 * Written as the INVERSE of the decoder behaviour specified in SURVEY.md
 * appendix A (A.1 frame grammar, A.2 adaptive Golomb-Rice with zero-run mode,
 * A.3 sign-LMS adaptive FIR predictor, A.4 mid/side mix + wasted bytes, A.7
 * validity constraints).  It does not share code with oracle/ or with the
 * CUDA decoder, so decode(encode(pcm)) == pcm is an independent check on
 * both.
 *
 * Every per-frame choice (element type, N, wasted bytes, escape, mix, and per
 * channel order / quant / riceMod / initial coefficients) is an INPUT, so the
 * tests can sweep the whole header space, including values a real encoder
 * never emits (order 31 = delta mode, order 0, quant 0, riceMod 0 / 7,
 * pred_type != 0, unknown tags).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct alacenc_cfg {
    int32_t sample_size;            /* 16 or 24 */
    int32_t max_samples_per_frame;  /* frames with n != this get hassize=1 */
    int32_t rice_history_mult;      /* pb  (40) */
    int32_t rice_initial_history;   /* mb  (10) */
    int32_t rice_kmodifier;         /* kb  (14) */
} alacenc_cfg;

typedef struct alacenc_frame {
    int32_t tag;                    /* 0 mono element, 1 stereo element (other values: written verbatim) */
    int32_t n;                      /* samples */
    int32_t force_hassize;          /* write the 32-bit count even if n == max */
    int32_t ub;                     /* wasted bytes (low ub*8 bits go to the shift planes) */
    int32_t escape;                 /* 1: uncompressed frame; -1: auto (escape if not smaller) */
    int32_t mix_shift, mix_weight;
    int32_t pred_type[2], quant[2], rice_mod[2], order[2];
    int32_t coef[2][32];            /* initial coefficients, written as 16-bit */
    int32_t adapt_passes;           /* >0: pre-run the predictor over the frame and transmit the adapted coefs */
    int32_t end_tag;                /* 1: append 3-bit END tag (7) before byte padding */
} alacenc_frame;

/* ---- bit writer --------------------------------------------------------- */
typedef struct { uint8_t *p; size_t cap; size_t bytepos; uint64_t acc; int nacc; int overflow; } bitw;

static void bw_put(bitw *w, uint32_t v, int n)
{
    /* n in 0..32, MSB first */
    if (n <= 0) return;
    uint64_t m = n >= 32 ? 0xffffffffull : ((1ull << n) - 1ull);
    w->acc = (w->acc << n) | ((uint64_t)v & m);
    w->nacc += n;
    while (w->nacc >= 8) {
        uint8_t b = (uint8_t)(w->acc >> (w->nacc - 8));
        if (w->bytepos < w->cap) w->p[w->bytepos] = b; else w->overflow = 1;
        w->bytepos++;
        w->nacc -= 8;
    }
}

static size_t bw_finish(bitw *w)
{
    if (w->nacc > 0) {
        uint8_t b = (uint8_t)((w->acc << (8 - w->nacc)) & 0xff);
        if (w->bytepos < w->cap) w->p[w->bytepos] = b; else w->overflow = 1;
        w->bytepos++;
        w->nacc = 0;
    }
    return w->bytepos;
}

static inline int32_t sx(int32_t v, int bits)
{
    int mv = 32 - bits;
    uint32_t u = (uint32_t)v << mv;
    int32_t s = (int32_t)u;
    /* arithmetic shift */
    return s >= 0 ? (int32_t)((uint32_t)s >> mv) : (int32_t)~((~(uint32_t)s) >> mv);
}
static inline int32_t sar(int32_t a, int n)
{
    n &= 31;
    return a >= 0 ? (int32_t)((uint32_t)a >> n) : (int32_t)~((~(uint32_t)a) >> n);
}
static inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

/* decoder's leading-zero count: 40 for zero */
static int lz40(int32_t v)
{
    uint32_t u = (uint32_t)v;
    if (!u) return 40;
    return __builtin_clz(u);
}

/* ---- forward predictor (A.3 mirrored) ----------------------------------- */
/* x[0..n) -> residual e[0..n); coef is adapted in place exactly as the
 * decoder will adapt it. */
static void forward_predict(const int32_t *x, int32_t *e, int n, int rss, int32_t *coef, int order, int quant)
{
    if (n <= 0) return;
    e[0] = x[0];
    if (order == 0) { for (int i = 1; i < n; i++) e[i] = x[i]; return; }
    if (order == 31) { for (int i = 1; i < n; i++) e[i] = sx(wsub(x[i], x[i - 1]), rss); return; }
    for (int i = 1; i <= order && i < n; i++) e[i] = sx(wsub(x[i], x[i - 1]), rss);
    for (int i = order + 1; i < n; i++) {
        int b = i - order - 1;
        int32_t sum = 0;
        for (int j = 0; j < order; j++) sum = wadd(sum, wmul(wsub(x[b + order - j], x[b]), coef[j]));
        int32_t pred = wadd(sar(wadd((int32_t)(1u << ((quant - 1) & 31)), sum), quant), x[b]);
        int32_t err = sx(wsub(x[i], pred), rss);
        e[i] = err;
        if (err > 0) {
            for (int p = order - 1; p >= 0 && err > 0; p--) {
                int32_t d = wsub(x[b], x[b + order - p]);
                int32_t s = d < 0 ? -1 : (d > 0 ? 1 : 0);
                coef[p] = wsub(coef[p], s);
                err = wsub(err, wmul(sar(wmul(d, s), quant), order - p));
            }
        } else if (err < 0) {
            for (int p = order - 1; p >= 0 && err < 0; p--) {
                int32_t d = wsub(x[b], x[b + order - p]);
                int32_t s = d < 0 ? 1 : (d > 0 ? -1 : 0);
                coef[p] = wsub(coef[p], s);
                err = wsub(err, wmul(sar(wmul(d, s), quant), order - p));
            }
        }
    }
}

/* ---- adaptive Rice writer (A.2 mirrored) -------------------------------- */
static void put_symbol(bitw *w, uint32_t v, int rawbits, int k, uint32_t mask)
{
    uint32_t M = (uint32_t)(((1u << (k & 31)) - 1u) & mask);
    if (k == 1) M = 1;
    uint32_t q = M ? v / M : 9, r = M ? v % M : 0;
    if (q >= 9) {
        bw_put(w, 0x1ff, 9);
        bw_put(w, v, rawbits);
        return;
    }
    bw_put(w, (1u << q) - 1u, (int)q);
    bw_put(w, 0, 1);
    if (k != 1) {
        if (r == 0) bw_put(w, 0, k - 1);
        else bw_put(w, r + 1, k);
    }
}

static void rice_encode(bitw *w, const int32_t *e, int n, int rss, const alacenc_cfg *cfg, int rice_mod)
{
    int32_t hist = cfg->rice_initial_history;
    const int kmod = cfg->rice_kmodifier;
    const int32_t mult = rice_mod * (cfg->rice_history_mult / 4);
    const uint32_t kmask = (1u << kmod) - 1u;
    int sign_mod = 0;
    for (int i = 0; i < n; i++) {
        int t = 31 - lz40(wadd(sar(hist, 9), 3));
        int k = t < kmod ? t : kmod;
        int32_t s = e[i];
        uint32_t dv = s >= 0 ? 2u * (uint32_t)s : 2u * (uint32_t)(-(int64_t)s) - 1u;
        put_symbol(w, dv - (uint32_t)sign_mod, rss, k, 0xffffffffu);
        sign_mod = 0;
        hist = dv > 0xFFFF ? 0xFFFF : wsub(wadd(hist, wmul((int32_t)dv, mult)), sar(wmul(hist, mult), 9));
        if (hist < 128 && i + 1 < n) {
            sign_mod = 1;
            k = lz40(hist) + ((hist + 16) / 64) - 24;
            int run = 0;
            while (i + 1 + run < n && e[i + 1 + run] == 0 && run < 0xFFFF) run++;
            put_symbol(w, (uint32_t)run, 16, k, kmask);
            i += run;
            if (run > 0xFFFF) sign_mod = 0;
            hist = 0;
        }
    }
}

/* ---- one frame ----------------------------------------------------------- */
static size_t encode_frame_mode(const alacenc_cfg *cfg, const alacenc_frame *fp, int escape,
                                const int32_t *left, const int32_t *right,
                                uint8_t *out, size_t cap, int32_t *scratch)
{
    bitw w = { out, cap, 0, 0, 0, 0 };
    const int stereo = (fp->tag == 1);
    const int ech = stereo ? 2 : 1;
    const int n = fp->n;
    const int ss = cfg->sample_size;
    const int ub = fp->ub;
    const int hassize = fp->force_hassize || n != cfg->max_samples_per_frame;

    bw_put(&w, (uint32_t)fp->tag, 3);
    bw_put(&w, 0, 4);
    bw_put(&w, 0, 12);
    bw_put(&w, (uint32_t)hassize, 1);
    bw_put(&w, (uint32_t)ub, 2);
    bw_put(&w, (uint32_t)escape, 1);
    if (hassize) bw_put(&w, (uint32_t)n, 32);

    if (fp->tag != 0 && fp->tag != 1) {
        /* unknown element: a few filler bits so the frame is not empty */
        bw_put(&w, 0xA5A5, 16);
    } else if (escape) {
        for (int i = 0; i < n; i++) {
            bw_put(&w, (uint32_t)left[i] & ((1u << ss) - 1u), ss);
            if (stereo) bw_put(&w, (uint32_t)right[i] & ((1u << ss) - 1u), ss);
        }
    } else {
        const int sh = ub * 8;
        const int rss = ss - sh + (stereo ? 1 : 0);
        int32_t *pa = scratch, *pb = scratch + n, *ea = scratch + 2 * n, *eb = scratch + 3 * n;
        /* split off wasted bytes, then mix (A.4 inverse) */
        for (int i = 0; i < n; i++) {
            int32_t l = sar(left[i], sh);
            if (!stereo) { pa[i] = l; continue; }
            int32_t r = sar(right[i], sh);
            if (fp->mix_weight != 0) {
                int32_t d = wsub(l, r);
                pa[i] = wadd(r, sar(wmul(d, fp->mix_weight), fp->mix_shift));
                pb[i] = d;
            } else { pa[i] = l; pb[i] = r; }
        }
        bw_put(&w, stereo ? (uint32_t)fp->mix_shift : 0, 8);
        bw_put(&w, stereo ? (uint32_t)fp->mix_weight : 0, 8);
        int32_t coef[2][32];
        for (int c = 0; c < ech; c++) {
            int order = fp->order[c];
            for (int j = 0; j < 32; j++) coef[c][j] = j < order ? (int16_t)fp->coef[c][j] : 0;
            if (order > 0 && order < 31)
                for (int p = 0; p < fp->adapt_passes; p++) {
                    forward_predict(c ? pb : pa, c ? eb : ea, n, rss, coef[c], order, fp->quant[c]);
                    for (int j = 0; j < order; j++) {
                        if (coef[c][j] > 32767) coef[c][j] = 32767;
                        if (coef[c][j] < -32768) coef[c][j] = -32768;
                    }
                }
            bw_put(&w, (uint32_t)fp->pred_type[c], 4);
            bw_put(&w, (uint32_t)fp->quant[c], 4);
            bw_put(&w, (uint32_t)fp->rice_mod[c], 3);
            bw_put(&w, (uint32_t)order, 5);
            for (int j = 0; j < order; j++) bw_put(&w, (uint32_t)coef[c][j] & 0xffffu, 16);
        }
        if (ub) {
            const uint32_t m = (1u << sh) - 1u;
            for (int i = 0; i < n; i++) {
                bw_put(&w, (uint32_t)left[i] & m, sh);
                if (stereo) bw_put(&w, (uint32_t)right[i] & m, sh);
            }
        }
        for (int c = 0; c < ech; c++) {
            forward_predict(c ? pb : pa, c ? eb : ea, n, rss, coef[c], fp->order[c], fp->quant[c]);
            rice_encode(&w, c ? eb : ea, n, rss, cfg, fp->rice_mod[c]);
        }
    }
    if (fp->end_tag) bw_put(&w, 7, 3);
    size_t bytes = bw_finish(&w);
    if (w.overflow || bytes > cap) return 0;
    return bytes;
}

/* Returns the frame's byte length (0 on overflow of `cap`).  `left`/`right`
 * are planar int32 samples in the signed sample_size range (right ignored for
 * mono elements).  escape == -1 picks the smaller of the two encodings, like a
 * real encoder's fallback when compression expands (A.7 item 7). */
size_t alacenc_encode_frame(const alacenc_cfg *cfg, const alacenc_frame *fp,
                            const int32_t *left, const int32_t *right,
                            uint8_t *out, size_t cap)
{
    int n = fp->n > 0 ? fp->n : 1;
    int32_t *scratch = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)n);
    size_t bytes;
    if (fp->escape >= 0) {
        bytes = encode_frame_mode(cfg, fp, fp->escape, left, right, out, cap, scratch);
    } else {
        bytes = encode_frame_mode(cfg, fp, 0, left, right, out, cap, scratch);
        size_t raw = (size_t)((23 + 32 + (size_t)fp->n * (fp->tag == 1 ? 2 : 1) * cfg->sample_size + 3 + 7) / 8);
        if (bytes == 0 || bytes >= raw) {
            alacenc_frame f2 = *fp;
            f2.ub = 0;
            bytes = encode_frame_mode(cfg, &f2, 1, left, right, out, cap, scratch);
        }
    }
    free(scratch);
    return bytes;
}

/* Whole track: `frames[i]` describes frame i, which consumes frames[i].n
 * sample-frames from the planar inputs.  Frame byte sizes go to stsz; frames
 * are laid out back to back in `out`.  Returns total bytes, 0 on overflow.
 * Frames are independent, so batches are encoded in parallel (OpenMP) into
 * per-frame slots and then compacted in order. */
size_t alacenc_encode_track(const alacenc_cfg *cfg, const alacenc_frame *frames, uint32_t n_frames,
                            const int32_t *left, const int32_t *right,
                            uint8_t *out, size_t cap, uint32_t *stsz)
{
    enum { BATCH = 512 };
    size_t *spos = (size_t *)malloc(sizeof(size_t) * ((size_t)n_frames + 1));
    size_t pos = 0, slot = 0;
    for (uint32_t f = 0; f < n_frames; f++) {
        spos[f] = pos;
        pos += (size_t)frames[f].n;
        size_t need = (size_t)frames[f].n * 2 * 5 + 256;
        if (need > slot) slot = need;
    }
    uint8_t *tmp = (uint8_t *)malloc(slot * BATCH);
    size_t off = 0;
    int fail = 0;
    for (uint32_t b0 = 0; b0 < n_frames && !fail; b0 += BATCH) {
        uint32_t b1 = b0 + BATCH < n_frames ? b0 + BATCH : n_frames;
#pragma omp parallel for schedule(dynamic, 4)
        for (uint32_t f = b0; f < b1; f++) {
            size_t b = alacenc_encode_frame(cfg, &frames[f], left + spos[f], right ? right + spos[f] : NULL,
                                            tmp + (size_t)(f - b0) * slot, slot);
            stsz[f] = (uint32_t)b;
        }
        for (uint32_t f = b0; f < b1; f++) {
            if (stsz[f] == 0 || off + stsz[f] > cap) { fail = 1; break; }
            memcpy(out + off, tmp + (size_t)(f - b0) * slot, stsz[f]);
            off += stsz[f];
        }
    }
    free(tmp);
    free(spos);
    return fail ? 0 : off;
}

/* ---- seeded synthetic PCM (SURVEY.md 8(d) "value distributions") --------- */
#include <math.h>

static inline uint64_t sm64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t *s) { return (double)(sm64(s) >> 11) * (1.0 / 9007199254740992.0); }
static inline double urange(uint64_t *s, double a, double b) { return a + (b - a) * u01(s); }

/* flags: 1 = silence gaps, 2 = full-scale noise bursts, 4 = spans with zeroed
 * low 8/16 bits (24-bit only), 8 = sparse full-scale impulses (Rice escapes).
 * Output is planar: out[c * n + i].  Deterministic in (seed, n, ...) and
 * independent of the thread count. */
void alacgen_signal(uint64_t seed, int64_t n, int sample_size, int sample_rate, int channels,
                    int flags, int32_t *out)
{
    enum { BLK = 1 << 16, MAXP = 5 };
    uint64_t s = seed * 0xD1342543DE82EF95ull + 12345;
    const double full = (double)((1 << (sample_size - 1)) - 1);
    int np_ = 2 + (int)(sm64(&s) % 4);
    double f[MAXP], a[MAXP], ph[MAXP], ef[MAXP];
    for (int p = 0; p < np_; p++) {
        f[p] = urange(&s, 40.0, 6000.0);
        a[p] = urange(&s, 0.02, 0.25);
        ph[p] = urange(&s, 0.0, 6.283185307179586);
        ef[p] = urange(&s, 0.05, 0.8);
    }
    double gain[2], nz[2], a1[2], sidef[2];
    for (int c = 0; c < 2; c++) {
        gain[c] = c == 0 ? 1.0 : urange(&s, 0.6, 1.0);
        nz[c] = exp(urange(&s, log(2e-5), log(4e-3)));
        a1[c] = urange(&s, 0.5, 0.97);
        sidef[c] = urange(&s, 100.0, 900.0);
    }
    const int64_t nblk = (n + BLK - 1) / BLK;
    const double w = 6.283185307179586 / (double)sample_rate;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < nblk; b++) {
        int64_t i0 = b * BLK, i1 = i0 + BLK < n ? i0 + BLK : n;
        for (int c = 0; c < channels; c++) {
            uint64_t rs = seed ^ (0xA5A5A5A5ull * (uint64_t)(b + 1)) ^ ((uint64_t)c << 56);
            double ar = 0.0;
            for (int64_t i = i0; i < i1; i++) {
                double t = (double)i, v = 0.0;
                for (int p = 0; p < np_; p++)
                    v += a[p] * (0.6 + 0.4 * sin(w * ef[p] * t + ph[p])) * sin(w * f[p] * t + ph[p]);
                v *= gain[c];
                if (c) v += 0.15 * sin(w * sidef[c] * t);
                double g = (u01(&rs) + u01(&rs) + u01(&rs) + u01(&rs) - 2.0) * 1.7320508;
                ar = a1[c] * ar + g * nz[c];
                v += ar;
                double q = rint(v * full * 0.7);
                if (q > full) q = full;
                if (q < -full - 1) q = -full - 1;
                out[(int64_t)c * n + i] = (int32_t)q;
            }
        }
    }
    if (n <= 8192) return;
    if (flags & 1) {
        int64_t cnt = n / 400000 > 1 ? n / 400000 : 1;
        for (int64_t k = 0; k < cnt; k++) {
            int64_t st = (int64_t)(u01(&s) * (double)(n - 4096)), ln = 500 + (int64_t)(u01(&s) * 29500.0);
            for (int c = 0; c < channels; c++)
                for (int64_t i = st; i < st + ln && i < n; i++) out[(int64_t)c * n + i] = 0;
        }
    }
    if (flags & 2) {
        int64_t cnt = n / 600000 > 1 ? n / 600000 : 1;
        for (int64_t k = 0; k < cnt; k++) {
            int64_t st = (int64_t)(u01(&s) * (double)(n - 4096)), ln = 200 + (int64_t)(u01(&s) * 8800.0);
            for (int c = 0; c < channels; c++)
                for (int64_t i = st; i < st + ln && i < n; i++)
                    out[(int64_t)c * n + i] = (int32_t)floor(urange(&s, -full - 1, full + 1));
        }
    }
    if ((flags & 4) && sample_size == 24) {
        int64_t cnt = n / 150000 > 2 ? n / 150000 : 2;
        for (int64_t k = 0; k < cnt; k++) {
            int64_t st = (int64_t)(u01(&s) * (double)(n - 4096)), ln = 4096 + (int64_t)(u01(&s) * 116000.0);
            int bits = 8 * (1 + (int)(sm64(&s) & 1));
            for (int c = 0; c < channels; c++)
                for (int64_t i = st; i < st + ln && i < n; i++) {
                    int32_t v = out[(int64_t)c * n + i];
                    out[(int64_t)c * n + i] = (int32_t)((uint32_t)(v >> bits) << bits);
                }
        }
    }
}
