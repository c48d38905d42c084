// isodemux.hpp -- tolerant ISO-BMFF / QuickTime demuxer for ALAC-in-MP4.
//
// The reference's QtMovieT (mirrored in alacnet.hpp) only accepts a narrow grammar: 32-bit atom
// sizes, moov before mdat, a fixed child order inside minf, no atom it does not know, at most 16
// stts runs, 32-bit stco, and it addresses frames sequentially from the start of mdat
// (QTMovieT.cs:61-107,200-225,258-331; DemuxResT.cs:27; AlacContext.cs:194-195).  Real files
// (iTunes, ffmpeg) break most of those.  IsoDemux walks any box order, 64-bit sizes, skips what it
// does not need, reads stco or co64, and resolves every frame's byte offset through stsc x stco x stsz,
// so chunked and gapped layouts decode too (alacgpu_add_track_offsets).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/alacgpu.h"

namespace alacnet {

struct IsoTrack {
    alacgpu_track_cfg cfg{};
    std::vector<uint64_t> offsets;      // byte offset of every frame in the file
    std::vector<uint32_t> sizes;        // stsz
    std::vector<uint32_t> durations;    // sample-frames per frame (stts expanded)
    uint64_t total_samples = 0;         // sum of durations
};

// Parses the first ALAC audio track of `file`.  Returns false (and a reason) when there is none.
bool IsoDemux(const uint8_t *file, size_t len, IsoTrack &out, std::string &err);

// RIFF/WAVE header for `pcm_bytes` of interleaved little-endian PCM (16 or 24 bit).
std::vector<uint8_t> WavHeader(int sample_rate, int bits, int channels, uint64_t pcm_bytes);

}  // namespace alacnet
