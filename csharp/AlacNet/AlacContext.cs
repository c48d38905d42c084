// AlacContext.cs -- drop-in replacement for ALACDecoder/AlacContext.cs whose decode path runs in
// libalacgpu.so.  Public surface identical to the reference (AlacContext.cs:20-338): constructors,
// Read(byte[]) = one frame per call, the Get* getters, LastSampleNumber, SetPosition, Dispose.
// The demuxer (QtMovieT / DemuxResT / MyStream) is the reference's own, unchanged: its tables are
// handed to alacgpu_add_track instead of to AlacFile.
//
// NOT COMPILED HERE (no .NET in the image).  The C++ twin alac/net_b200/host/alacnet.cpp has the
// same control flow line for line and IS exercised on the GPU by tests/test_gpu_host_mirror.py.
using System;
using System.IO;
using System.Linq;
using ALACdotNET.Decoder.Gpu;

namespace ALACdotNET.Decoder
{
    public class AlacContext : IDisposable
    {
        public AlacContext(Stream baseStream, bool disposeStream) : this(baseStream)
        {
            _disposeStream = disposeStream;
        }

        public unsafe AlacContext(Stream baseStream)
        {
            _demuxRes = new DemuxResT();
            _inputStream = new BinaryReader(baseStream);
            _myStream = new MyStream(_inputStream);
            var qtmovie = new QtMovieT(_myStream, _demuxRes);
            var headerRead = qtmovie.ReadHeader();
            if (headerRead == MdatPosStatus.None || headerRead == MdatPosStatus.CannotSeekToMdatPosition)
            {
                SelfDispose(true);
                throw new IOException("Error while loading the QuickTime movie headers.");
            }
            // the stream now sits on the first frame: read the frames once (the reference reads them
            // one MyStream.Read at a time, AlacContext.cs:194-195) and stage them in HBM
            var first = baseStream.Position;
            long payload = _demuxRes.SampleByteSize.Sum(s => (long)s);
            _mdat = new byte[payload];
            int got = 0;
            while (got < payload)
            {
                int n = baseStream.Read(_mdat, got, (int)Math.Min(payload - got, 1 << 20));
                if (n <= 0) break;
                got += n;
            }
            _mdatLen = got;
            _mdatFilePos = first;
            var opts = new AlacGpuOpts { StructSize = (uint)sizeof(AlacGpuOpts) };
            NativeMethods.Check(null, NativeMethods.alacgpu_create(null, 0, ref opts, out _gpu), "alacgpu_create");
            Stage(0);
        }

        private unsafe void Stage(long firstFrameOffset)
        {
            if (_stagedFirst == firstFrameOffset) return;
            NativeMethods.Check(_gpu, NativeMethods.alacgpu_clear_tracks(_gpu), "alacgpu_clear_tracks");
            var cd = _demuxRes.CodecData;                       // AlacFile.SetInfo offsets (AlacFile.cs:63-93)
            var cfg = new AlacGpuTrackCfg
            {
                MaxSamplesPerFrame = (cd[24] << 24) + (cd[25] << 16) + (cd[26] << 8) + cd[27],
                SampleSize = _demuxRes.SampleSize,
                RiceHistoryMult = cd[30] & 0xff,
                RiceInitialHistory = cd[31] & 0xff,
                RiceKModifier = cd[32] & 0xff,
                NumChannels = _demuxRes.NumChannels,
                SampleRate = _demuxRes.SampleRate
            };
            var sizes = _demuxRes.SampleByteSize.Select(s => (uint)Math.Max(s, 0)).ToArray();
            fixed (byte* p = _mdat)
            fixed (uint* ps = sizes)
            {
                NativeMethods.Check(_gpu, NativeMethods.alacgpu_add_track(_gpu, ref cfg, p, (ulong)_mdatLen,
                    (ulong)Math.Max(firstFrameOffset, 0), ps, (uint)sizes.Length, out _), "alacgpu_add_track");
                // the library borrows `p` until the first decode: run it inside the fixed block
                NativeMethods.Check(_gpu, NativeMethods.alacgpu_prepare(_gpu, out _), "alacgpu_prepare");
            }
            _stagedFirst = firstFrameOffset;
        }

        private readonly DemuxResT _demuxRes;
        private readonly BinaryReader _inputStream;
        private readonly MyStream _myStream;
        private readonly AlacGpuHandle _gpu;
        private readonly byte[] _mdat;
        private readonly long _mdatLen;
        private readonly long _mdatFilePos;
        private long _stagedFirst = -1;
        private int _currentSampleBlock;
        private int _offset;
        private readonly byte[] _frame = new byte[65536];
        private readonly bool _disposeStream;
        private bool _disposedValue;

        public int LastSampleNumber { get; private set; }

        public int GetSampleRate() => _demuxRes.SampleRate != 0 ? _demuxRes.SampleRate : 44100;
        public int GetNumChannels() => _demuxRes.NumChannels != 0 ? _demuxRes.NumChannels : 2;
        public int GetBitsPerSample() => _demuxRes.SampleSize != 0 ? _demuxRes.SampleSize : 16;
        public int GetBytesPerSample() => _demuxRes.SampleSize != 0 ? (int)Math.Ceiling((double)_demuxRes.SampleSize / 8) : 2;

        public int GetNumSamples()
        {
            int total = 0;
            for (int i = 0; i < _demuxRes.SampleByteSize.Length; i++)
            {
                var info = TryGetSampleInfo(i);
                if (info == null) return -1;
                total += info.Value.duration;
            }
            return total;
        }

        private (int size, int duration)? TryGetSampleInfo(int samplenum)
        {
            int accum = 0, cur = 0;
            if (samplenum >= _demuxRes.SampleByteSize.Length) return null;
            if (_demuxRes.NumTimeToSamples == 0) return null;
            while (_demuxRes.TimeToSample[cur].SampleCount + accum <= samplenum)
            {
                accum += _demuxRes.TimeToSample[cur].SampleCount;
                cur++;
                if (cur >= _demuxRes.NumTimeToSamples) return null;
            }
            return (_demuxRes.SampleByteSize[samplenum], _demuxRes.TimeToSample[cur].SampleDuration);
        }

        /// <summary>Reads and decodes a single ALAC frame (AlacContext.cs:163-172).</summary>
        public unsafe int Read(byte[] buffer)
        {
            if (_currentSampleBlock >= _demuxRes.SampleByteSize.Length) return 0;
            var info = TryGetSampleInfo(_currentSampleBlock);
            if (info == null) return 0;
            uint got;
            fixed (byte* p = _frame)
                NativeMethods.Check(_gpu, NativeMethods.alacgpu_read_frame(_gpu, 0, (uint)_currentSampleBlock, p, (uint)_frame.Length, out got), "alacgpu_read_frame");
            NativeMethods.alacgpu_frame_status(_gpu, 0, (uint)_currentSampleBlock, out var status);
            if ((AlacGpuFrameStatus)status == AlacGpuFrameStatus.PredType)
                throw new Exception("FIXME: unhandled predicition type");                 // AlacFile.cs:650,660
            _currentSampleBlock++;
            LastSampleNumber += info.Value.duration;
            // post-seek fix-up exactly as AlacContext.cs:200-202 (ints = samples for 16-bit, bytes for 24-bit)
            int outputBytes = (int)got - _offset * GetBytesPerSample();
            int skip = _offset * (GetBytesPerSample() == 2 ? 2 : 1);
            _offset = 0;
            if (outputBytes <= 0) return 0;
            Array.Copy(_frame, skip, buffer, 0, outputBytes);
            return outputBytes;
        }

        public void SetPosition(long position)
        {
            int currentPosition = 0;
            int currentSample = 0;
            for (int i = 0; i < _demuxRes.Stsc.Length; i++)
            {
                var chunkInfo = _demuxRes.Stsc[i];
                var lastChunk = i < _demuxRes.Stsc.Length - 1 ? _demuxRes.Stsc[i + 1].FirstChunk : _demuxRes.Stco.Length;
                for (int chunk = chunkInfo.FirstChunk; chunk <= lastChunk; chunk++)
                {
                    long pos = _demuxRes.Stco[chunk - 1];
                    int sampleCount = chunkInfo.SamplesPerChunk;
                    while (sampleCount > 0)
                    {
                        var sampleInfo = TryGetSampleInfo(currentSample);
                        if (sampleInfo == null) break;
                        currentPosition += sampleInfo.Value.duration;
                        if (position < currentPosition)
                        {
                            long before = 0;
                            for (int f = 0; f < currentSample; f++) before += _demuxRes.SampleByteSize[f];
                            Stage(pos - _mdatFilePos - before);       // frame currentSample must start at file offset `pos`
                            _currentSampleBlock = currentSample;
                            LastSampleNumber = currentPosition;
                            _offset = (int)(position - (currentPosition - sampleInfo.Value.duration)) * GetNumChannels();
                            return;
                        }
                        pos += sampleInfo.Value.size;
                        currentSample++;
                        sampleCount--;
                    }
                }
            }
        }

        protected virtual void Dispose(bool disposing) => SelfDispose(disposing);

        private void SelfDispose(bool disposing)
        {
            if (_disposedValue) return;
            _gpu?.Dispose();
            if (disposing && _disposeStream) _inputStream?.Dispose();
            _disposedValue = true;
        }

        public void Dispose() => Dispose(true);
    }
}
