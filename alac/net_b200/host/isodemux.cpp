// isodemux.cpp -- see isodemux.hpp.  Host-only container parsing; no sample is decoded here.
#include "isodemux.hpp"

#include <algorithm>
#include <cstring>

namespace alacnet {

namespace {

struct Box { uint32_t type; uint64_t body, end; };   // body = first payload byte, end = one past the box

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint64_t be64(const uint8_t *p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }
constexpr uint32_t cc(char a, char b, char c, char d)
{
    return ((uint32_t)(uint8_t)a << 24) | ((uint32_t)(uint8_t)b << 16) | ((uint32_t)(uint8_t)c << 8) | (uint32_t)(uint8_t)d;
}

// next box at `pos` inside [pos, limit); false when nothing parseable is left
// (all comparisons are written so that a crafted 64-bit size cannot wrap: pos <= limit always holds)
bool next_box(const uint8_t *f, uint64_t pos, uint64_t limit, Box &b)
{
    if (pos > limit || limit - pos < 8) return false;
    uint64_t size = be32(f + pos);
    b.type = be32(f + pos + 4);
    uint64_t hdr = 8;
    if (size == 1) {                       // 64-bit largesize
        if (limit - pos < 16) return false;
        size = be64(f + pos + 8);
        hdr = 16;
    } else if (size == 0) {                // box runs to the end of its container
        size = limit - pos;
    }
    if (size < hdr || size > limit - pos) return false;
    b.body = pos + hdr;
    b.end = pos + size;
    return true;
}

struct Tables {
    bool have_alac = false;
    alacgpu_track_cfg cfg{};
    std::vector<uint32_t> sizes;
    std::vector<std::pair<uint32_t, uint32_t>> stts;                  // (count, duration)
    struct Stsc { uint32_t first_chunk, per_chunk; };
    std::vector<Stsc> stsc;
    std::vector<uint64_t> chunk_off;
};

// the 'alac' box inside the sample entry: size 'alac' ver/flags(4) + 24-byte ALACSpecificConfig
bool parse_alac_config(const uint8_t *f, const Box &b, Tables &t)
{
    if (b.end - b.body < 4 + 24) return false;
    const uint8_t *c = f + b.body + 4;
    t.cfg.max_samples_per_frame = (int32_t)be32(c);
    t.cfg.sample_size = c[5];
    t.cfg.rice_history_mult = c[6];
    t.cfg.rice_initial_history = c[7];
    t.cfg.rice_kmodifier = c[8];
    t.cfg.num_channels = c[9];
    t.cfg.sample_rate = (int32_t)be32(c + 20);
    t.have_alac = true;
    return true;
}

void parse_stsd(const uint8_t *f, const Box &b, Tables &t)
{
    if (b.end - b.body < 8) return;
    const uint32_t n = be32(f + b.body + 4);
    uint64_t pos = b.body + 8;
    for (uint32_t i = 0; i < n; i++) {
        Box e;
        if (!next_box(f, pos, b.end, e)) return;
        if (e.type == cc('a', 'l', 'a', 'c') && e.end - e.body >= 28) {
            // SoundSampleEntry: 6 reserved, 2 dref index, then version(2) ... ; v0 = 28 bytes of
            // fields, v1 adds 16, v2 adds 36.  The codec box(es) follow.
            const uint32_t version = ((uint32_t)f[e.body + 8] << 8) | f[e.body + 9];
            uint64_t child = e.body + 28 + (version == 1 ? 16 : version == 2 ? 36 : 0);
            Box c;
            while (next_box(f, child, e.end, c)) {
                if (c.type == cc('a', 'l', 'a', 'c') && parse_alac_config(f, c, t)) return;
                if (c.type == cc('w', 'a', 'v', 'e')) {      // QuickTime wraps it: wave { frma, alac, ... }
                    uint64_t w = c.body;
                    Box d;
                    while (next_box(f, w, c.end, d)) {
                        if (d.type == cc('a', 'l', 'a', 'c') && parse_alac_config(f, d, t)) return;
                        w = d.end;
                    }
                }
                child = c.end;
            }
        }
        pos = e.end;
    }
}

void parse_stbl(const uint8_t *f, const Box &stbl, Tables &t, const uint64_t file_len)
{
    uint64_t pos = stbl.body;
    Box b;
    while (next_box(f, pos, stbl.end, b)) {
        const uint64_t n_avail = b.end - b.body;
        const uint8_t *p = f + b.body;
        if (b.type == cc('s', 't', 's', 'd')) {
            parse_stsd(f, b, t);
        } else if (b.type == cc('s', 't', 't', 's') && n_avail >= 8) {
            const uint64_t n = std::min<uint64_t>(be32(p + 4), (n_avail - 8) / 8);
            for (uint64_t i = 0; i < n; i++) t.stts.push_back({be32(p + 8 + 8 * i), be32(p + 12 + 8 * i)});
        } else if (b.type == cc('s', 't', 's', 'z') && n_avail >= 12) {
            const uint32_t uniform = be32(p + 4);
            uint64_t n = be32(p + 8);
            if (uniform) {
                // a uniform table costs no file bytes, so bound it by what the file could hold: every frame
                // occupies `uniform` bytes of a file no longer than the box's container
                n = std::min<uint64_t>(n, file_len / uniform + 1);
                n = std::min<uint64_t>(n, 1u << 28);
                t.sizes.assign((size_t)n, uniform);
            } else {
                n = std::min<uint64_t>(n, (n_avail - 12) / 4);
                t.sizes.resize((size_t)n);
                for (uint64_t i = 0; i < n; i++) t.sizes[(size_t)i] = be32(p + 12 + 4 * i);
            }
        } else if (b.type == cc('s', 't', 's', 'c') && n_avail >= 8) {
            const uint64_t n = std::min<uint64_t>(be32(p + 4), (n_avail - 8) / 12);
            for (uint64_t i = 0; i < n; i++) t.stsc.push_back({be32(p + 8 + 12 * i), be32(p + 12 + 12 * i)});
        } else if (b.type == cc('s', 't', 'c', 'o') && n_avail >= 8) {
            const uint64_t n = std::min<uint64_t>(be32(p + 4), (n_avail - 8) / 4);
            for (uint64_t i = 0; i < n; i++) t.chunk_off.push_back(be32(p + 8 + 4 * i));
        } else if (b.type == cc('c', 'o', '6', '4') && n_avail >= 8) {
            const uint64_t n = std::min<uint64_t>(be32(p + 4), (n_avail - 8) / 8);
            for (uint64_t i = 0; i < n; i++) t.chunk_off.push_back(be64(p + 8 + 8 * i));
        }
        pos = b.end;
    }
}

// descend moov/trak/mdia/minf/stbl; stop at the first track that carries an ALAC sample entry
bool find_track(const uint8_t *f, uint64_t lo, uint64_t hi, int depth, Tables &t, const uint64_t file_len)
{
    uint64_t pos = lo;
    Box b;
    while (next_box(f, pos, hi, b)) {
        if (b.type == cc('s', 't', 'b', 'l')) {
            Tables cand;
            parse_stbl(f, b, cand, file_len);
            if (cand.have_alac) { t = std::move(cand); return true; }
        } else if (depth < 6 && (b.type == cc('m', 'o', 'o', 'v') || b.type == cc('t', 'r', 'a', 'k') ||
                                 b.type == cc('m', 'd', 'i', 'a') || b.type == cc('m', 'i', 'n', 'f'))) {
            if (find_track(f, b.body, b.end, depth + 1, t, file_len)) return true;
        }
        pos = b.end;
    }
    return false;
}

}  // namespace

bool IsoDemux(const uint8_t *file, size_t len, IsoTrack &out, std::string &err)
{
    Tables t;
    if (!find_track(file, 0, len, 0, t, len)) { err = "no ALAC audio track found"; return false; }
    const size_t n = t.sizes.size();
    out.cfg = t.cfg;
    out.sizes = t.sizes;
    // durations: stts runs expanded; frames beyond the table get the cookie's frame length
    out.durations.assign(n, (uint32_t)std::max(t.cfg.max_samples_per_frame, 0));
    size_t k = 0;
    for (const auto &run : t.stts)
        for (uint32_t i = 0; i < run.first && k < n; i++) out.durations[k++] = run.second;
    out.total_samples = 0;
    for (uint32_t dur : out.durations) out.total_samples += dur;
    // offsets: chunk c holds per_chunk samples (stsc run-length over chunk numbers, 1-based) back to back
    out.offsets.assign(n, 0);
    if (t.chunk_off.empty()) { err = "no chunk offset table (stco / co64)"; return false; }
    if (t.stsc.empty()) t.stsc.push_back({1, (uint32_t)n});
    size_t s = 0;
    for (size_t ci = 0; ci < t.chunk_off.size() && s < n; ci++) {
        const uint32_t chunk_no = (uint32_t)ci + 1;
        size_t r = 0;
        while (r + 1 < t.stsc.size() && t.stsc[r + 1].first_chunk <= chunk_no) r++;
        uint64_t off = t.chunk_off[ci];
        for (uint32_t j = 0; j < t.stsc[r].per_chunk && s < n; j++, s++) {
            out.offsets[s] = off;
            off += t.sizes[s];
        }
    }
    if (s < n) {                       // tables do not cover every sample: keep what is addressable
        out.offsets.resize(s);
        out.sizes.resize(s);
        out.durations.resize(s);
    }
    return true;
}

std::vector<uint8_t> WavHeader(int sample_rate, int bits, int channels, uint64_t pcm_bytes)
{
    std::vector<uint8_t> h(44);
    auto put32 = [&](size_t at, uint32_t v) { for (int i = 0; i < 4; i++) h[at + i] = (uint8_t)(v >> (8 * i)); };
    auto put16 = [&](size_t at, uint32_t v) { h[at] = (uint8_t)v; h[at + 1] = (uint8_t)(v >> 8); };
    const uint32_t data = (uint32_t)std::min<uint64_t>(pcm_bytes, 0xFFFFFFFFu - 36u);
    const uint32_t block = (uint32_t)(channels * (bits / 8));
    memcpy(&h[0], "RIFF", 4); put32(4, 36 + data); memcpy(&h[8], "WAVEfmt ", 8);
    put32(16, 16); put16(20, 1); put16(22, (uint32_t)channels); put32(24, (uint32_t)sample_rate);
    put32(28, (uint32_t)sample_rate * block); put16(32, block); put16(34, (uint32_t)bits);
    memcpy(&h[36], "data", 4); put32(40, data);
    return h;
}

}  // namespace alacnet

// ---- flat C surface for the Python tests ------------------------------------------------------
extern "C" {
__attribute__((visibility("default")))
int alacnet_iso_demux(const uint8_t *file, uint64_t len, alacgpu_track_cfg *cfg, uint64_t *offsets, uint32_t *sizes,
                      uint32_t *durations, uint32_t cap, uint32_t *n_frames, uint64_t *total_samples)
{
    alacnet::IsoTrack t;
    std::string err;
    if (!alacnet::IsoDemux(file, (size_t)len, t, err)) return -1;
    *cfg = t.cfg;
    *n_frames = (uint32_t)t.sizes.size();
    *total_samples = t.total_samples;
    for (uint32_t i = 0; i < *n_frames && i < cap; i++) { offsets[i] = t.offsets[i]; sizes[i] = t.sizes[i]; durations[i] = t.durations[i]; }
    return 0;
}
__attribute__((visibility("default")))
int alacnet_wav_header(int rate, int bits, int channels, uint64_t pcm_bytes, uint8_t *out44)
{
    const std::vector<uint8_t> h = alacnet::WavHeader(rate, bits, channels, pcm_bytes);
    memcpy(out44, h.data(), 44);
    return 44;
}
}
