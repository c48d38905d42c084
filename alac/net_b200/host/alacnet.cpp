// alacnet.cpp -- see alacnet.hpp.  Host-only code: container parsing, table walking and
// byte shuffling.  Every PCM byte comes out of libalacgpu.so (alacgpu_read_frame).
//
// This file is a BEHAVIOURAL TWIN of the reference's C# host (ALACDecoder/QTMovieT.cs, MyStream.cs,
// AlacContext.cs:83-295, AlacNetNAudioAdapter/ALACFileReader.cs:89-126), not product logic: the shipped
// host stays the reference's own C# over P/Invoke (csharp/).  The image has no .NET, so this
// twin is what lets the parity tests drive libalacgpu.so exactly the way AlacContext / ALACFileReader would
// (same method names, same loops, the reference's quirks kept on purpose).  It decodes nothing and is not
// meant to grow.
#include "alacnet.hpp"

#include <algorithm>
#include <cstring>

#include "../../../include/alacgpu.h"

namespace alacnet {

namespace {
constexpr int32_t fourcc(char a, char b, char c, char d)
{
    return (int32_t)(((uint32_t)(uint8_t)a << 24) | ((uint32_t)(uint8_t)b << 16) | ((uint32_t)(uint8_t)c << 8) | (uint32_t)(uint8_t)d);
}
}  // namespace

// ---- MyStream ---------------------------------------------------------------------
int MyStream::Read(int size, uint8_t *buf, int start_pos)
{
    int64_t avail = base_pos_ < len_ ? len_ - base_pos_ : 0;
    int n = (int)std::min<int64_t>(std::max(size, 0), avail);
    if (n > 0) memcpy(buf + start_pos, data_ + base_pos_, (size_t)n);
    base_pos_ += n;
    Position += n;
    return n;
}

void MyStream::Read(int size, int32_t *buf, int start_pos)
{
    std::vector<uint8_t> tmp((size_t)std::max(size, 0));
    int n = Read(size, tmp.data(), 0);
    for (int i = 0; i < n; i++) buf[start_pos + i] = tmp[(size_t)i];
}

int32_t MyStream::ReadUint32()
{
    Read(4, read_buffer_, 0);
    return (int32_t)(((uint32_t)read_buffer_[0] << 24) | ((uint32_t)read_buffer_[1] << 16) |
                     ((uint32_t)read_buffer_[2] << 8) | (uint32_t)read_buffer_[3]);
}

int32_t MyStream::ReadUint16()
{
    Read(2, read_buffer_, 0);
    return (int32_t)(((uint32_t)read_buffer_[0] << 8) | (uint32_t)read_buffer_[1]);
}

int32_t MyStream::ReadUint8()
{
    int64_t pos_before = Position;
    Read(1, read_buffer_, 0);
    Position = pos_before + 1;                       // MyStream.cs:92 adds 1 whatever was read
    return read_buffer_[0];
}

void MyStream::Skip(int32_t skip)
{
    if (skip < 0) throw ArgumentException("Request to seek backwards in stream is not supported");
    base_pos_ += skip;                               // Stream.Seek may land past the end
    Position += base_pos_;                           // MyStream.cs:99-100 adds the ABSOLUTE position (reference quirk)
}

int64_t MyStream::Seek(int64_t pos)
{
    if (pos < 0) return -1;
    base_pos_ = pos;
    return base_pos_;
}

// ---- QtMovieT ---------------------------------------------------------------------
MdatPosStatus QtMovieT::ReadHeader()
{
    int found_moov = 0, found_mdat = 0;
    for (;;) {
        const int32_t chunk_len = qt_.ReadUint32();
        if (qt_.EOF_()) return MdatPosStatus::None;                       // :58
        const int32_t chunk_id = qt_.ReadUint32();
        if (chunk_id == fourcc('f', 't', 'y', 'p')) {
            ReadChunkFtyp(chunk_len);
        } else if (chunk_id == fourcc('m', 'o', 'o', 'v')) {
            if (ReadChunkMoov(chunk_len) == 0) return MdatPosStatus::None;
            if (found_mdat) return SetSavedMdat();
            found_moov = 1;
        } else if (chunk_id == fourcc('m', 'd', 'a', 't')) {
            ProcReadChunkMdat(chunk_len, found_moov ? 0 : 1);
            if (found_moov) return MdatPosStatus::Ok;                     // stream sits on the first frame (:88-91)
            found_mdat = 1;
        } else if (chunk_id == fourcc('f', 'r', 'e', 'e') || chunk_id == fourcc('j', 'u', 'n', 'k')) {
            qt_.Skip(chunk_len - 8);
        } else {
            return MdatPosStatus::None;                                   // unknown top-level atom (:103-107)
        }
    }
}

void QtMovieT::ReadChunkFtyp(int32_t len)
{
    int32_t remaining = len - 8;
    const int32_t type = qt_.ReadUint32();
    remaining -= 4;
    if (type != fourcc('M', '4', 'A', ' ')) return;                       // :116-120: rest of the atom is NOT consumed
    qt_.ReadUint32();
    remaining -= 4;
    while (remaining != 0) {                                              // compatible brands
        if (qt_.EOF_()) throw IOException("ftyp atom runs past the end of the file");   // the reference would spin here
        qt_.ReadUint32();
        remaining -= 4;
    }
}

// Container atoms share one loop shape: while (remaining != 0) { len; id; dispatch; remaining -= len; }
#define ALACNET_SUBCHUNK_HEADER()                                           \
    const int32_t sub_len = qt_.ReadUint32();                               \
    if (sub_len <= 1 || sub_len > remaining) return 0;                      \
    const int32_t sub_id = qt_.ReadUint32()

int QtMovieT::ReadChunkMoov(int32_t len)
{
    int32_t remaining = len - 8;
    while (remaining != 0) {
        ALACNET_SUBCHUNK_HEADER();
        if (sub_id == fourcc('t', 'r', 'a', 'k')) {
            if (ReadChunkTrak(sub_len) == 0) return 0;
        } else if (sub_id == fourcc('m', 'v', 'h', 'd') || sub_id == fourcc('u', 'd', 't', 'a') ||
                   sub_id == fourcc('e', 'l', 's', 't') || sub_id == fourcc('i', 'o', 'd', 's') ||
                   sub_id == fourcc('f', 'r', 'e', 'e')) {
            qt_.Skip(sub_len - 8);
        } else {
            return 0;                                                     // :714-718
        }
        remaining -= sub_len;
    }
    return 1;
}

int QtMovieT::ReadChunkTrak(int32_t len)
{
    int32_t remaining = len - 8;
    while (remaining != 0) {
        ALACNET_SUBCHUNK_HEADER();
        if (sub_id == fourcc('m', 'd', 'i', 'a')) {
            if (ReadChunkMedia(sub_len) == 0) return 0;
        } else if (sub_id == fourcc('t', 'k', 'h', 'd') || sub_id == fourcc('e', 'd', 't', 's')) {
            qt_.Skip(sub_len - 8);
        } else {
            return 0;                                                     // :169-173
        }
        remaining -= sub_len;
    }
    return 1;
}

int QtMovieT::ReadChunkMedia(int32_t len)
{
    int32_t remaining = len - 8;
    while (remaining != 0) {
        ALACNET_SUBCHUNK_HEADER();
        if (sub_id == fourcc('m', 'd', 'h', 'd')) {
            qt_.Skip(sub_len - 8);
        } else if (sub_id == fourcc('h', 'd', 'l', 'r')) {
            ProcReadChunkHdlr(sub_len);
        } else if (sub_id == fourcc('m', 'i', 'n', 'f')) {
            if (ReadChunkMediaInfo(sub_len) == 0) return 0;
        } else {
            return 0;                                                     // :367-371
        }
        remaining -= sub_len;
    }
    return 1;
}

void QtMovieT::ProcReadChunkHdlr(int32_t len)
{
    int32_t remaining = len - 8;
    for (int i = 0; i < 4; i++) qt_.ReadUint8();                          // version + flags
    remaining -= 4;
    qt_.ReadUint32(); qt_.ReadUint32();                                   // component type / subtype
    remaining -= 8;
    qt_.ReadUint32();                                                     // manufacturer
    remaining -= 4;
    qt_.ReadUint32(); qt_.ReadUint32();                                   // flags
    remaining -= 8;
    qt_.ReadUint8();                                                      // name length
    remaining -= 1;
    qt_.Skip(remaining);                                                  // negative -> ArgumentException, as in the reference
}

int QtMovieT::ReadChunkMediaInfo(int32_t len)
{
    int32_t remaining = len - 8;
    if (qt_.ReadUint32() != 16) return 0;                                 // smhd must come first, size 16 (:273-277)
    if (qt_.ReadUint32() != fourcc('s', 'm', 'h', 'd')) return 0;
    qt_.Skip(16 - 8);
    remaining -= 16;
    const int32_t dinf_size = qt_.ReadUint32();
    if (qt_.ReadUint32() != fourcc('d', 'i', 'n', 'f')) return 0;
    qt_.Skip(dinf_size - 8);
    remaining -= dinf_size;
    const int32_t stbl_size = qt_.ReadUint32();
    if (qt_.ReadUint32() != fourcc('s', 't', 'b', 'l')) return 0;
    if (ReadChunkStbl(stbl_size) == 0) return 0;
    remaining -= stbl_size;
    if (remaining != 0) qt_.Skip(remaining);
    return 1;
}

int QtMovieT::ReadChunkStbl(int32_t len)
{
    int32_t remaining = len - 8;
    while (remaining != 0) {
        ALACNET_SUBCHUNK_HEADER();
        if (sub_id == fourcc('s', 't', 's', 'd')) {
            if (ReadChunkStsd() == 0) return 0;
        } else if (sub_id == fourcc('s', 't', 't', 's')) {
            ProcReadOverChunkStts(sub_len);
        } else if (sub_id == fourcc('s', 't', 's', 'z')) {
            SkipOverChunkStsz(sub_len);
        } else if (sub_id == fourcc('s', 't', 's', 'c')) {
            ReadChunkStsc();
        } else if (sub_id == fourcc('s', 't', 'c', 'o')) {
            ReadChunkStco();
        } else {
            return 0;                                                     // :221-225
        }
        remaining -= sub_len;
    }
    return 1;
}

int QtMovieT::ReadChunkStsd()
{
    for (int i = 0; i < 4; i++) qt_.ReadUint8();                          // version + flags
    if (qt_.ReadUint32() != 1) return 0;                                  // exactly one entry (:430-434)
    const int32_t entry_size = qt_.ReadUint32();
    res_.Format = qt_.ReadUint32();
    int32_t entry_remaining = entry_size - 8;
    if (res_.Format != fourcc('a', 'l', 'a', 'c')) return 0;
    qt_.Skip(6);                entry_remaining -= 6;                     // reserved
    qt_.ReadUint16();           entry_remaining -= 2;                     // version
    qt_.ReadUint16(); qt_.ReadUint32(); entry_remaining -= 6;             // revision, vendor
    qt_.ReadUint16();           entry_remaining -= 2;                     // the "extra 16 bits"
    qt_.Skip(4);                entry_remaining -= 4;                     // channels + bits (top level)
    qt_.ReadUint16(); qt_.ReadUint16(); entry_remaining -= 4;             // compression id, packet size
    qt_.Skip(4);                entry_remaining -= 4;                     // sample rate (top level)
    res_.CodecDataLength = entry_remaining + 12 + 8;
    if (res_.CodecDataLength > 1024 || entry_remaining < 0) return 0;     // :478-482 (negative would throw in C#)
    for (int i = 0; i < res_.CodecDataLength; i++) res_.CodecData[i] = 0;
    res_.CodecData[0] = 0x0c000000;
    res_.CodecData[1] = fourcc('a', 'm', 'r', 'f');
    res_.CodecData[2] = fourcc('c', 'a', 'l', 'a');
    qt_.Read(entry_remaining, res_.CodecData, 12);                        // the 'alac' atom lands at offset 12
    res_.SampleSize = res_.CodecData[29] & 0xff;                          // :508-513
    res_.NumChannels = res_.CodecData[33] & 0xff;
    res_.SampleRate = ((res_.CodecData[44] & 0xff) << 24) | ((res_.CodecData[45] & 0xff) << 16) |
                      ((res_.CodecData[46] & 0xff) << 8) | (res_.CodecData[47] & 0xff);
    res_.FormatRead = 1;
    return 1;
}

void QtMovieT::ProcReadOverChunkStts(int32_t len)
{
    int32_t remaining = len - 8;
    for (int i = 0; i < 4; i++) qt_.ReadUint8();
    remaining -= 4;
    const int32_t n = qt_.ReadUint32();
    remaining -= 4;
    res_.NumTimeToSamples = n;
    for (int32_t i = 0; i < n; i++) {
        if (i >= 16) throw DecoderException("stts has more than 16 entries (DemuxResT.cs:27: IndexOutOfRangeException)");
        res_.TimeToSample[i].SampleCount = qt_.ReadUint32();
        res_.TimeToSample[i].SampleDuration = qt_.ReadUint32();
        remaining -= 8;
    }
    if (remaining != 0) qt_.Skip(remaining);
}

void QtMovieT::SkipOverChunkStsz(int32_t len)
{
    int32_t remaining = len - 8;
    for (int i = 0; i < 4; i++) qt_.ReadUint8();
    remaining -= 4;
    const int32_t uniform = qt_.ReadUint32();
    if (uniform != 0) {                                                   // :577-590
        const int32_t n = qt_.ReadUint32();
        if (n < 0) throw DecoderException("stsz: negative sample count");
        res_.SampleByteSize.assign((size_t)n, uniform);
        return;
    }
    remaining -= 4;
    const int32_t n = qt_.ReadUint32();
    remaining -= 4;
    if (n < 0 || (int64_t)n * 4 > qt_.length()) throw DecoderException("stsz: sample count beyond the file");
    res_.SampleByteSize.resize((size_t)n);
    for (int32_t i = 0; i < n; i++) {
        res_.SampleByteSize[(size_t)i] = qt_.ReadUint32();
        remaining -= 4;
    }
    if (remaining != 0) qt_.Skip(remaining);
}

void QtMovieT::ReadChunkStsc()
{
    qt_.Skip(4);
    const int32_t n = qt_.ReadUint32();
    if (n < 0 || (int64_t)n * 12 > qt_.length()) throw DecoderException("stsc: entry count beyond the file");
    res_.Stsc.resize((size_t)n);
    for (auto &c : res_.Stsc) {
        c.FirstChunk = qt_.ReadUint32();
        c.SamplesPerChunk = qt_.ReadUint32();
        c.SampleDescIndex = qt_.ReadUint32();
    }
    res_.HaveStsc = true;
}

void QtMovieT::ReadChunkStco()
{
    qt_.Skip(4);
    const int32_t n = qt_.ReadUint32();
    if (n < 0 || (int64_t)n * 4 > qt_.length()) throw DecoderException("stco: entry count beyond the file");
    res_.Stco.resize((size_t)n);
    for (auto &o : res_.Stco) o = qt_.ReadUint32();
    res_.HaveStco = true;
}

void QtMovieT::ProcReadChunkMdat(int32_t len, int skip_mdat)
{
    const int32_t remaining = len - 8;
    if (remaining == 0) return;
    res_.MdatLen = remaining;
    if (skip_mdat) {
        saved_mdat_pos_ = qt_.Position;
        qt_.Skip(remaining);
    }
}

MdatPosStatus QtMovieT::SetSavedMdat()
{
    if (saved_mdat_pos_ == -1) return MdatPosStatus::NoValidSaveMdatPosition;
    // MyStream.Seek returns the new POSITION, so any saved position but 0 reads as failure (:744-748)
    if (qt_.Seek(saved_mdat_pos_) != 0) return MdatPosStatus::CannotSeekToMdatPosition;
    return MdatPosStatus::Ok;
}

// ---- AlacContext ------------------------------------------------------------------
static void gpu_check(alacgpu_ctx *ctx, int32_t rc, const char *what)
{
    if (rc == ALACGPU_OK) return;
    std::string msg = std::string(what) + ": " + alacgpu_strerror(rc);
    const char *detail = ctx ? alacgpu_last_error(ctx) : "";
    if (detail && *detail) msg += std::string(" (") + detail + ")";
    // "FIXME: unimplemented sample size" (AlacFile.cs:574,715) surfaces as UNSUPPORTED at add_track
    throw DecoderException(msg);
}

AlacContext::AlacContext(const uint8_t *file, size_t len, int device) : file_(file), len_(len), device_(device)
{
    MyStream stream(file, len);
    QtMovieT qt(stream, res_);
    const MdatPosStatus st = qt.ReadHeader();
    if (st == MdatPosStatus::None || st == MdatPosStatus::CannotSeekToMdatPosition)
        throw IOException("Error while loading the QuickTime movie headers.");       // AlacContext.cs:47-51
    mdat_pos_ = stream.BasePosition();
    const int32_t dev = device_;
    gpu_check(nullptr, alacgpu_create(&dev, 1, nullptr, &gpu_), "alacgpu_create");
    Stage(mdat_pos_);
}

AlacContext::~AlacContext()
{
    if (gpu_) alacgpu_destroy(gpu_);
}

// new AlacFile(SampleSize, NumChannels) + SetInfo(CodecData) + the sequential frame addressing of
// UnpackSamples (AlacContext.cs:54-55, :194-195): frame i = SampleByteSize[i] bytes, back to back
// from `first_frame_offset`.
void AlacContext::Stage(int64_t first_frame_offset)
{
    if (staged_first_ == first_frame_offset) return;
    gpu_check(gpu_, alacgpu_clear_tracks(gpu_), "alacgpu_clear_tracks");
    const int32_t *cd = res_.CodecData;                                    // AlacFile.cs:63-93
    alacgpu_track_cfg cfg{};
    cfg.max_samples_per_frame = (int32_t)(((uint32_t)cd[24] << 24) + ((uint32_t)cd[25] << 16) + ((uint32_t)cd[26] << 8) + (uint32_t)cd[27]);
    cfg.sample_size = res_.SampleSize;          // the constructor argument; equals cookie byte 29
    cfg.rice_history_mult = cd[30] & 0xff;
    cfg.rice_initial_history = cd[31] & 0xff;
    cfg.rice_kmodifier = cd[32] & 0xff;
    cfg.num_channels = res_.NumChannels;
    cfg.sample_rate = res_.SampleRate;
    std::vector<uint32_t> sizes(res_.SampleByteSize.size());
    for (size_t i = 0; i < sizes.size(); i++) sizes[i] = (uint32_t)std::max(res_.SampleByteSize[i], 0);
    int32_t tid = -1;
    const uint64_t off = (uint64_t)std::max<int64_t>(first_frame_offset, 0);
    gpu_check(gpu_, alacgpu_add_track(gpu_, &cfg, file_, len_, off, sizes.data(), (uint32_t)sizes.size(), &tid),
              "alacgpu_add_track");
    staged_first_ = first_frame_offset;
}

int AlacContext::GetSampleRate() const { return res_.SampleRate != 0 ? res_.SampleRate : 44100; }
int AlacContext::GetNumChannels() const { return res_.NumChannels != 0 ? res_.NumChannels : 2; }
int AlacContext::GetBitsPerSample() const { return res_.SampleSize != 0 ? res_.SampleSize : 16; }
int AlacContext::GetBytesPerSample() const { return res_.SampleSize != 0 ? (res_.SampleSize + 7) / 8 : 2; }

AlacContext::Dur AlacContext::TryGetSampleInfo(int samplenum) const
{
    int accum = 0, cur = 0;
    if (samplenum < 0 || (size_t)samplenum >= res_.SampleByteSize.size()) return {0, 0, false};
    if (res_.NumTimeToSamples == 0) return {0, 0, false};
    while (res_.TimeToSample[cur].SampleCount + accum <= samplenum) {
        accum += res_.TimeToSample[cur].SampleCount;
        cur++;
        if (cur >= res_.NumTimeToSamples || cur >= 16) return {0, 0, false};
    }
    return {res_.SampleByteSize[(size_t)samplenum], res_.TimeToSample[cur].SampleDuration, true};
}

int AlacContext::GetNumSamples() const
{
    // sum of stts durations over all frames; walks the run-length table once instead of per frame
    int64_t total = 0;
    const size_t n = res_.SampleByteSize.size();
    if (n == 0) return 0;
    if (res_.NumTimeToSamples == 0) return -1;
    size_t covered = 0;
    for (int i = 0; i < res_.NumTimeToSamples && i < 16 && covered < n; i++) {
        const int64_t cnt = std::max(res_.TimeToSample[i].SampleCount, 0);
        const size_t take = (size_t)std::min<int64_t>(cnt, (int64_t)(n - covered));
        total += (int64_t)take * res_.TimeToSample[i].SampleDuration;
        covered += take;
    }
    if (covered < n) return -1;                                            // "Could not read some sample" (:121-127)
    return (int)total;
}

int AlacContext::Read(uint8_t *buffer, size_t buffer_len)
{
    if (disposed_) throw DecoderException("AlacContext used after Dispose");
    if ((size_t)current_sample_block_ >= res_.SampleByteSize.size()) return 0;       // :182-186
    const Dur info = TryGetSampleInfo(current_sample_block_);
    if (!info.ok) return 0;                                                          // :187-193
    frame_.resize(65536);
    uint32_t got = 0;
    gpu_check(gpu_, alacgpu_read_frame(gpu_, 0, (uint32_t)current_sample_block_, frame_.data(), (uint32_t)frame_.size(), &got),
              "alacgpu_read_frame");
    int32_t status = 0;
    alacgpu_frame_status(gpu_, 0, (uint32_t)current_sample_block_, &status);
    if (status == ALACGPU_FRAME_PRED_TYPE) throw DecoderException("FIXME: unhandled predicition type");   // AlacFile.cs:650,660
    current_sample_block_ += 1;
    LastSampleNumber += info.duration;
    // post-seek fix-up (:200-202): `_offset` is in ints of the decoder's int buffer -- samples for
    // 16-bit, BYTES for 24-bit (one int per byte) -- while the byte count drops by _offset * bytesPerSample.
    // Reproduced as is: 24-bit seeks land short by 2/3 of the intra-frame offset (reference bug).
    int out_bytes = (int)got - offset_ * GetBytesPerSample();
    const size_t skip = (size_t)offset_ * (GetBytesPerSample() == 2 ? 2u : 1u);
    offset_ = 0;
    if (out_bytes <= 0) return out_bytes < 0 ? 0 : 0;
    if ((size_t)out_bytes > buffer_len) throw ArgumentException("destination array was not long enough");
    memcpy(buffer, frame_.data() + skip, (size_t)out_bytes);
    return out_bytes;
}

void AlacContext::SetPosition(int64_t position)
{
    if (!res_.HaveStsc || !res_.HaveStco) throw DecoderException("SetPosition needs stsc and stco (NullReferenceException in the reference)");
    int current_position = 0, current_sample = 0;
    const int n_stsc = (int)res_.Stsc.size();
    for (int i = 0; i < n_stsc; i++) {
        const ChunkInfo &ci = res_.Stsc[(size_t)i];
        const int last_chunk = i < n_stsc - 1 ? res_.Stsc[(size_t)i + 1].FirstChunk : (int)res_.Stco.size();
        for (int chunk = ci.FirstChunk; chunk <= last_chunk; chunk++) {              // inclusive, as written (:270)
            if (chunk < 1 || (size_t)chunk > res_.Stco.size()) throw DecoderException("stco index out of range");
            int64_t pos = res_.Stco[(size_t)chunk - 1];
            int sample_count = ci.SamplesPerChunk;
            while (sample_count > 0) {
                const Dur info = TryGetSampleInfo(current_sample);
                if (!info.ok) break;
                current_position += info.duration;
                if (position < current_position) {
                    // the reference seeks the base stream to `pos` and keeps reading sequentially (:281-286)
                    int64_t before = 0;
                    for (int f = 0; f < current_sample; f++) before += res_.SampleByteSize[(size_t)f];
                    Stage(pos - before);                      // frame current_sample must start at `pos`
                    current_sample_block_ = current_sample;
                    LastSampleNumber = current_position;
                    offset_ = (int)(position - (current_position - info.duration)) * GetNumChannels();
                    return;
                }
                pos += info.size;
                current_sample++;
                sample_count--;
            }
        }
    }
}

void AlacContext::Dispose() { disposed_ = true; }

// ---- ALACFileReader ---------------------------------------------------------------
ALACFileReader::ALACFileReader(const uint8_t *file, size_t len, int device)
    : ctx_(new AlacContext(file, len, device))
{
    fmt_.SampleRate = ctx_->GetSampleRate();
    fmt_.BitsPerSample = ctx_->GetBytesPerSample() * 8;
    fmt_.Channels = ctx_->GetNumChannels();
    length_ = (int64_t)ctx_->GetNumSamples() * fmt_.BlockAlign();
    decompress_.resize((size_t)65546 * (size_t)(fmt_.BitsPerSample / 8) * (size_t)fmt_.Channels);
}

int64_t ALACFileReader::Position() const { return (int64_t)ctx_->LastSampleNumber * fmt_.BlockAlign(); }

void ALACFileReader::SetPosition(int64_t value)
{
    std::lock_guard<std::mutex> g(reposition_lock_);       // :68
    ctx_->SetPosition(value / fmt_.BlockAlign());
    leftovers_ = 0;
}

void ALACFileReader::Dispose()
{
    std::lock_guard<std::mutex> g(reposition_lock_);       // :121
    ctx_->Dispose();
}

int ALACFileReader::Read(uint8_t *buffer, int offset, int count)
{
    int bytes_read = 0;
    std::lock_guard<std::mutex> g(reposition_lock_);       // :92
    while (bytes_read < count) {
        if (leftovers_ > 0) {
            const int to_copy = std::min(leftovers_, count - bytes_read);
            memcpy(buffer + offset, decompress_.data() + buffer_offset_, (size_t)to_copy);
            leftovers_ -= to_copy;
            buffer_offset_ = leftovers_ == 0 ? 0 : buffer_offset_ + to_copy;
            bytes_read += to_copy;
            offset += to_copy;
        }
        if (bytes_read >= count) break;
        buffer_offset_ = 0;
        const int unpacked = ctx_->Read(decompress_.data(), decompress_.size());
        if (unpacked == 0) break;
        leftovers_ += unpacked;
    }
    return bytes_read;
}

}  // namespace alacnet

// ---- flat C surface for ctypes tests and non-C++ callers ---------------------------
// (not part of the drop-in boundary -- that is include/alacgpu.h; this only makes the C++
// mirror callable from the Python test-suite)
#define ALACNET_API __attribute__((visibility("default")))

enum { ALACNET_OK = 0, ALACNET_IO_EXCEPTION = -101, ALACNET_ARGUMENT_EXCEPTION = -102, ALACNET_DECODER_EXCEPTION = -103 };

static thread_local std::string g_last_error;

template <typename F>
static int guarded(F &&f)
{
    try {
        f();
        return ALACNET_OK;
    } catch (const alacnet::IOException &e) {
        g_last_error = e.what();
        return ALACNET_IO_EXCEPTION;
    } catch (const alacnet::ArgumentException &e) {
        g_last_error = e.what();
        return ALACNET_ARGUMENT_EXCEPTION;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return ALACNET_DECODER_EXCEPTION;
    }
}

extern "C" {

ALACNET_API const char *alacnet_last_error(void) { return g_last_error.c_str(); }

ALACNET_API int alacnet_context_open(const uint8_t *file, uint64_t len, int device, void **out)
{
    return guarded([&] { *out = new alacnet::AlacContext(file, (size_t)len, device); });
}
ALACNET_API void alacnet_context_close(void *h) { delete static_cast<alacnet::AlacContext *>(h); }
ALACNET_API int alacnet_context_read(void *h, uint8_t *buf, uint64_t cap, int *bytes)
{
    return guarded([&] { *bytes = static_cast<alacnet::AlacContext *>(h)->Read(buf, (size_t)cap); });
}
ALACNET_API int alacnet_context_info(void *h, int *rate, int *channels, int *bits, int *bytes_per_sample, int *num_samples)
{
    return guarded([&] {
        auto *c = static_cast<alacnet::AlacContext *>(h);
        *rate = c->GetSampleRate(); *channels = c->GetNumChannels(); *bits = c->GetBitsPerSample();
        *bytes_per_sample = c->GetBytesPerSample(); *num_samples = c->GetNumSamples();
    });
}
ALACNET_API int alacnet_context_set_position(void *h, int64_t pos)
{
    return guarded([&] { static_cast<alacnet::AlacContext *>(h)->SetPosition(pos); });
}
ALACNET_API int alacnet_context_last_sample_number(void *h) { return static_cast<alacnet::AlacContext *>(h)->LastSampleNumber; }
ALACNET_API int64_t alacnet_context_mdat_offset(void *h) { return static_cast<alacnet::AlacContext *>(h)->mdat_offset(); }
ALACNET_API int alacnet_context_frame_count(void *h) { return (int)static_cast<alacnet::AlacContext *>(h)->demux().SampleByteSize.size(); }

// demux only (no GPU): lets the CPU test-suite check the container grammar
ALACNET_API int alacnet_demux(const uint8_t *file, uint64_t len, int *status, int *sample_size, int *channels, int *rate,
                              int64_t *mdat_pos, int *n_frames, uint32_t *sizes, int sizes_cap, int *codec_data48)
{
    return guarded([&] {
        alacnet::MyStream s(file, (size_t)len);
        alacnet::DemuxResT r;
        alacnet::QtMovieT qt(s, r);
        *status = (int)qt.ReadHeader();
        *sample_size = r.SampleSize; *channels = r.NumChannels; *rate = r.SampleRate;
        *mdat_pos = s.BasePosition();
        *n_frames = (int)r.SampleByteSize.size();
        for (int i = 0; i < *n_frames && i < sizes_cap; i++) sizes[i] = (uint32_t)r.SampleByteSize[(size_t)i];
        if (codec_data48) for (int i = 0; i < 48; i++) codec_data48[i] = r.CodecData[i];
    });
}

ALACNET_API int alacnet_reader_open(const uint8_t *file, uint64_t len, int device, void **out)
{
    return guarded([&] { *out = new alacnet::ALACFileReader(file, (size_t)len, device); });
}
ALACNET_API void alacnet_reader_close(void *h) { delete static_cast<alacnet::ALACFileReader *>(h); }
ALACNET_API int alacnet_reader_read(void *h, uint8_t *buf, int offset, int count, int *bytes)
{
    return guarded([&] { *bytes = static_cast<alacnet::ALACFileReader *>(h)->Read(buf, offset, count); });
}
ALACNET_API int64_t alacnet_reader_length(void *h) { return static_cast<alacnet::ALACFileReader *>(h)->Length(); }
ALACNET_API int64_t alacnet_reader_position(void *h) { return static_cast<alacnet::ALACFileReader *>(h)->Position(); }
ALACNET_API int alacnet_reader_set_position(void *h, int64_t v)
{
    return guarded([&] { static_cast<alacnet::ALACFileReader *>(h)->SetPosition(v); });
}
ALACNET_API int alacnet_reader_format(void *h, int *rate, int *bits, int *channels, int *block_align)
{
    const alacnet::WaveFormat &f = static_cast<alacnet::ALACFileReader *>(h)->GetWaveFormat();
    *rate = f.SampleRate; *bits = f.BitsPerSample; *channels = f.Channels; *block_align = f.BlockAlign();
    return ALACNET_OK;
}

}  // extern "C"
