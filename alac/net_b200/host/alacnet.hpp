// alacnet.hpp -- C++ host mirror of the reference's public decode API, written over
// the C ABI of libalacgpu.so (include/alacgpu.h).
//
// The reference's host language is C# (netstandard2.0); neither dotnet nor mono exists
// in this image, so the host side that would normally be C# is mirrored here in C++ with
// the same type names, method names, argument meaning and error behaviour, and the C#
// sources a maintainer would drop into the reference live in csharp/ (INTEGRATION.md).
//
//   alacnet::MyStream        <- ALACDecoder/MyStream.cs:20-115   (big-endian reader)
//   alacnet::DemuxResT       <- ALACDecoder/DemuxResT.cs:16-35   (sample tables)
//   alacnet::QtMovieT        <- ALACDecoder/QTMovieT.cs:51-751   (atom walker, same acceptance grammar)
//   alacnet::AlacContext     <- ALACDecoder/AlacContext.cs:20-338 (open / Read one frame / seek / info)
//   alacnet::ALACFileReader  <- AlacNetNAudioAdapter/ALACFileReader.cs:22-127 (WaveStream.Read re-chunker)
//
// No sample is decoded on the host: AlacContext hands the demuxer's tables to
// alacgpu_add_track and pulls PCM frame by frame with alacgpu_read_frame.
#pragma once

#include <cstdint>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

struct alacgpu_ctx;

namespace alacnet {

struct IOException : std::runtime_error { using std::runtime_error::runtime_error; };
struct ArgumentException : std::invalid_argument { using std::invalid_argument::invalid_argument; };
struct DecoderException : std::runtime_error { using std::runtime_error::runtime_error; };

// BinaryReader over an in-memory file + the reference's MyStream on top of it.
class MyStream {
public:
    MyStream(const uint8_t *data, size_t len) : data_(data), len_((int64_t)len) {}
    int64_t Position = 0;                                  // MyStream.cs:23 (NOT the base position after a Skip)
    bool EOF_() const { return base_pos_ >= len_; }        // MyStream.cs:27
    int Read(int size, uint8_t *buf, int start_pos);       // MyStream.cs:47-52
    void Read(int size, int32_t *buf, int start_pos);      // MyStream.cs:35-45
    int32_t ReadUint32();                                  // MyStream.cs:54-68
    int32_t ReadUint16();                                  // MyStream.cs:77-86
    int32_t ReadUint8();                                   // MyStream.cs:88-94
    void Skip(int32_t skip);                               // MyStream.cs:96-101
    int64_t Seek(int64_t pos);                             // MyStream.cs:103-114
    int64_t BasePosition() const { return base_pos_; }
    const uint8_t *data() const { return data_; }
    int64_t length() const { return len_; }

private:
    const uint8_t *data_;
    int64_t len_;
    int64_t base_pos_ = 0;
    uint8_t read_buffer_[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // MyStream.cs:22: stale bytes survive short reads
};

struct SampleInfo { int32_t SampleCount = 0, SampleDuration = 0; };            // SampleInfo.cs
struct ChunkInfo { int32_t FirstChunk = 0, SamplesPerChunk = 0, SampleDescIndex = 0; };   // ChunkInfo.cs

struct DemuxResT {                                         // DemuxResT.cs:22-34
    int32_t FormatRead = 0, NumChannels = 0, SampleSize = 0, SampleRate = 0, Format = 0;
    SampleInfo TimeToSample[16];
    int32_t NumTimeToSamples = 0;
    std::vector<int32_t> SampleByteSize;
    int32_t CodecDataLength = 0;
    int32_t CodecData[1024] = {0};
    std::vector<int32_t> Stco;
    std::vector<ChunkInfo> Stsc;
    bool HaveStco = false, HaveStsc = false;
    int32_t MdatLen = 0;
};

enum class MdatPosStatus { None = 0, Ok = 1, NoValidSaveMdatPosition = 2, CannotSeekToMdatPosition = 3 };

class QtMovieT {
public:
    QtMovieT(MyStream &s, DemuxResT &d) : qt_(s), res_(d) {}
    MdatPosStatus ReadHeader();                            // QTMovieT.cs:51-109

private:
    void ReadChunkFtyp(int32_t len);
    int ReadChunkMoov(int32_t len);
    int ReadChunkTrak(int32_t len);
    int ReadChunkMedia(int32_t len);
    int ReadChunkMediaInfo(int32_t len);
    int ReadChunkStbl(int32_t len);
    int ReadChunkStsd();
    void ProcReadChunkHdlr(int32_t len);
    void ProcReadOverChunkStts(int32_t len);
    void SkipOverChunkStsz(int32_t len);
    void ReadChunkStsc();
    void ReadChunkStco();
    void ProcReadChunkMdat(int32_t len, int skip_mdat);
    MdatPosStatus SetSavedMdat();
    MyStream &qt_;
    DemuxResT &res_;
    int64_t saved_mdat_pos_ = 0;
};

class AlacContext {
public:
    // `file` is the whole .m4a (the reference takes a System.IO.Stream; here the stream is a
    // byte range that must stay valid for the life of the context).  device: CUDA ordinal.
    AlacContext(const uint8_t *file, size_t len, int device = 0);      // AlacContext.cs:36-56
    ~AlacContext();
    AlacContext(const AlacContext &) = delete;
    AlacContext &operator=(const AlacContext &) = delete;

    int Read(uint8_t *buffer, size_t buffer_len);          // AlacContext.cs:163-172: ONE frame per call, 0 at the end
    int GetSampleRate() const;                             // :83
    int GetNumChannels() const;                            // :89
    int GetBitsPerSample() const;                          // :95
    int GetBytesPerSample() const;                         // :101
    int GetNumSamples() const;                             // :108-122  (-1 if some frame has no duration)
    int LastSampleNumber = 0;                              // :76
    void SetPosition(int64_t position);                    // :262-295
    void Dispose();                                        // :314-318

    const DemuxResT &demux() const { return res_; }
    int64_t mdat_offset() const { return mdat_pos_; }

private:
    struct Dur { int32_t size, duration; bool ok; };
    Dur TryGetSampleInfo(int samplenum) const;             // :130-156
    void Stage(int64_t first_frame_offset);
    const uint8_t *file_;
    size_t len_;
    DemuxResT res_;
    alacgpu_ctx *gpu_ = nullptr;
    int device_;
    int64_t mdat_pos_ = 0;          // where frame 0 starts (stream position after ReadHeader)
    int64_t staged_first_ = -1;     // first_frame_offset the GPU track was staged with
    int current_sample_block_ = 0;  // :62
    int offset_ = 0;                // :63
    std::vector<uint8_t> frame_;    // one decoded frame
    bool disposed_ = false;
};

struct WaveFormat {                                        // NAudio.Wave.WaveFormat(rate, bits, channels)
    int SampleRate = 0, BitsPerSample = 0, Channels = 0;
    int BlockAlign() const { return Channels * (BitsPerSample / 8); }
    int AverageBytesPerSecond() const { return SampleRate * BlockAlign(); }
};

class ALACFileReader {                                     // ALACFileReader.cs:22-127
public:
    ALACFileReader(const uint8_t *file, size_t len, int device = 0);
    int64_t Length() const { return length_; }             // :58
    int64_t Position() const;                              // :63-65
    void SetPosition(int64_t value);                       // :66-73
    const WaveFormat &GetWaveFormat() const { return fmt_; }   // :79
    int Read(uint8_t *buffer, int offset, int count);      // :89-116
    void Dispose();                                        // :118-126
    AlacContext &context() { return *ctx_; }

private:
    std::unique_ptr<AlacContext> ctx_;
    WaveFormat fmt_;
    int64_t length_ = 0;
    int leftovers_ = 0, buffer_offset_ = 0;
    std::vector<uint8_t> decompress_;
    std::mutex reposition_lock_;                           // _repositionLock, :53: Read, the Position setter and Dispose exclude each other
};

}  // namespace alacnet
