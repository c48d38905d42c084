// alacgpu_decode -- batch ALAC (.m4a) -> WAV on the GPU.
//
//   alacgpu_decode [--device N] [--strict] -o OUT in1.m4a [in2.m4a ...]
//
// All inputs are demuxed on the host, handed to libalacgpu in ONE batch (alacgpu_add_track_offsets
// + alacgpu_decode_all) and written as RIFF/WAVE.  With one input OUT is the .wav path; with several
// OUT is a directory.  --strict uses the reference's own acceptance grammar (alacnet::QtMovieT)
// instead of the tolerant demuxer.  No sample is decoded on the host.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/alacgpu.h"
#include "../host/alacnet.hpp"
#include "../host/isodemux.hpp"

static bool read_file(const std::string &path, std::vector<uint8_t> &out)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const bool ok = out.empty() || fread(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

int main(int argc, char **argv)
{
    int device = 0;
    bool strict = false;
    std::string out;
    std::vector<std::string> inputs;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--strict")) strict = true;
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) out = argv[++i];
        else inputs.push_back(argv[i]);
    }
    if (inputs.empty() || out.empty()) {
        fprintf(stderr, "usage: alacgpu_decode [--device N] [--strict] -o OUT in1.m4a [in2.m4a ...]\n");
        return 2;
    }
    alacgpu_ctx *ctx = nullptr;
    const int32_t dev = device;
    int32_t rc = alacgpu_create(&dev, 1, nullptr, &ctx);
    if (rc) { fprintf(stderr, "alacgpu_create: %s\n", alacgpu_strerror(rc)); return 1; }

    std::vector<std::vector<uint8_t>> files(inputs.size());
    std::vector<alacnet::IsoTrack> tracks(inputs.size());
    for (size_t i = 0; i < inputs.size(); i++) {
        if (!read_file(inputs[i], files[i])) { fprintf(stderr, "%s: cannot read\n", inputs[i].c_str()); return 1; }
        alacnet::IsoTrack &t = tracks[i];
        if (strict) {                   // the reference's grammar and sequential addressing
            alacnet::MyStream s(files[i].data(), files[i].size());
            alacnet::DemuxResT res;
            const alacnet::MdatPosStatus st = alacnet::QtMovieT(s, res).ReadHeader();
            if (st == alacnet::MdatPosStatus::None || st == alacnet::MdatPosStatus::CannotSeekToMdatPosition) {
                fprintf(stderr, "%s: Error while loading the QuickTime movie headers.\n", inputs[i].c_str());
                return 1;
            }
            const int32_t *cd = res.CodecData;
            t.cfg.max_samples_per_frame = (cd[24] << 24) + (cd[25] << 16) + (cd[26] << 8) + cd[27];
            t.cfg.sample_size = res.SampleSize; t.cfg.num_channels = res.NumChannels; t.cfg.sample_rate = res.SampleRate;
            t.cfg.rice_history_mult = cd[30] & 0xff; t.cfg.rice_initial_history = cd[31] & 0xff; t.cfg.rice_kmodifier = cd[32] & 0xff;
            uint64_t off = (uint64_t)s.BasePosition();
            for (int32_t sz : res.SampleByteSize) { t.offsets.push_back(off); t.sizes.push_back((uint32_t)sz); off += (uint32_t)sz; }
        } else {
            std::string err;
            if (!alacnet::IsoDemux(files[i].data(), files[i].size(), t, err)) {
                fprintf(stderr, "%s: %s\n", inputs[i].c_str(), err.c_str());
                return 1;
            }
        }
        int32_t id = -1;
        rc = alacgpu_add_track_offsets(ctx, &t.cfg, files[i].data(), files[i].size(), t.offsets.data(), t.sizes.data(),
                                       (uint32_t)t.sizes.size(), &id);
        if (rc) { fprintf(stderr, "%s: %s (%s)\n", inputs[i].c_str(), alacgpu_strerror(rc), alacgpu_last_error(ctx)); return 1; }
    }
    uint64_t total = 0;
    alacgpu_total_pcm_bytes(ctx, &total);
    void *pcm = nullptr;
    if (alacgpu_host_alloc(total, &pcm)) { fprintf(stderr, "out of pinned memory\n"); return 1; }
    std::vector<uint64_t> off(inputs.size()), len(inputs.size());
    rc = alacgpu_decode_all(ctx, (uint8_t *)pcm, total, off.data(), len.data(), nullptr);
    if (rc) { fprintf(stderr, "alacgpu_decode_all: %s (%s)\n", alacgpu_strerror(rc), alacgpu_last_error(ctx)); return 1; }
    alacgpu_timing tm{};
    alacgpu_get_timing(ctx, &tm);
    for (size_t i = 0; i < inputs.size(); i++) {
        std::string path = out;
        if (inputs.size() > 1) {
            std::string base = inputs[i].substr(inputs[i].find_last_of('/') + 1);
            const size_t dot = base.find_last_of('.');
            path = out + "/" + (dot == std::string::npos ? base : base.substr(0, dot)) + ".wav";
        }
        FILE *f = fopen(path.c_str(), "wb");
        if (!f) { fprintf(stderr, "%s: cannot write\n", path.c_str()); return 1; }
        const alacgpu_track_cfg &c = tracks[i].cfg;
        const std::vector<uint8_t> h = alacnet::WavHeader(c.sample_rate ? c.sample_rate : 44100, c.sample_size, c.num_channels, len[i]);
        fwrite(h.data(), 1, h.size(), f);
        fwrite((const uint8_t *)pcm + off[i], 1, len[i], f);
        fclose(f);
    }
    fprintf(stderr, "decoded %zu file(s), %llu samples, kernels %.3f ms, total %.3f ms\n", inputs.size(),
            (unsigned long long)tm.samples, tm.kernels_ms, tm.total_ms);
    alacgpu_host_free(pcm);
    alacgpu_destroy(ctx);
    return 0;
}
