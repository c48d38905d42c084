// AlacBatchDecoder.cs -- the batch entry point the reference does not have: N files in, all PCM out
// in one alacgpu_decode_all.  This is the call configs[3] / configs[4] of BASELINE.json are measured on.
// NOT COMPILED HERE (no .NET in the image); its Python twin is alac/net_b200/decoder.py BatchDecoder.
using System;
using System.Collections.Generic;
using System.IO;
using System.Linq;
using System.Runtime.InteropServices;
using ALACdotNET.Decoder.Gpu;

namespace ALACdotNET.Decoder
{
    public sealed class AlacBatchDecoder : IDisposable
    {
        private readonly AlacGpuHandle _gpu;
        private readonly List<GCHandle> _pins = new List<GCHandle>();

        public unsafe AlacBatchDecoder(int[] deviceIds = null)
        {
            var opts = new AlacGpuOpts { StructSize = (uint)sizeof(AlacGpuOpts) };
            fixed (int* ids = deviceIds)
                NativeMethods.Check(null, NativeMethods.alacgpu_create(ids, deviceIds?.Length ?? 0, ref opts, out _gpu), "alacgpu_create");
        }

        /// <summary>Demux one .m4a with the reference's QtMovieT and hand its tables to the GPU.</summary>
        public unsafe int AddFile(byte[] m4a)
        {
            var res = new DemuxResT();
            using (var ms = new MemoryStream(m4a, false))
            {
                var st = new QtMovieT(new MyStream(new BinaryReader(ms)), res).ReadHeader();
                if (st == MdatPosStatus.None || st == MdatPosStatus.CannotSeekToMdatPosition)
                    throw new IOException("Error while loading the QuickTime movie headers.");
                var cd = res.CodecData;
                var cfg = new AlacGpuTrackCfg
                {
                    MaxSamplesPerFrame = (cd[24] << 24) + (cd[25] << 16) + (cd[26] << 8) + cd[27],
                    SampleSize = res.SampleSize, RiceHistoryMult = cd[30] & 0xff, RiceInitialHistory = cd[31] & 0xff,
                    RiceKModifier = cd[32] & 0xff, NumChannels = res.NumChannels, SampleRate = res.SampleRate
                };
                var sizes = res.SampleByteSize.Select(s => (uint)Math.Max(s, 0)).ToArray();
                var pin = GCHandle.Alloc(m4a, GCHandleType.Pinned);      // borrowed until DecodeAll returns
                _pins.Add(pin);
                fixed (uint* ps = sizes)
                {
                    NativeMethods.Check(_gpu, NativeMethods.alacgpu_add_track(_gpu, ref cfg, (byte*)pin.AddrOfPinnedObject(),
                        (ulong)m4a.Length, (ulong)ms.Position, ps, (uint)sizes.Length, out var id), "alacgpu_add_track");
                    return id;
                }
            }
        }

        /// <summary>Decode every frame of every added file; returns (pcm, per-track offset, per-track length).</summary>
        public unsafe (byte[] pcm, ulong[] off, ulong[] len) DecodeAll()
        {
            NativeMethods.Check(_gpu, NativeMethods.alacgpu_total_pcm_bytes(_gpu, out var total), "alacgpu_total_pcm_bytes");
            NativeMethods.Check(_gpu, NativeMethods.alacgpu_track_count(_gpu, out var n), "alacgpu_track_count");
            var pcm = new byte[total];
            var off = new ulong[n];
            var len = new ulong[n];
            fixed (byte* p = pcm) fixed (ulong* po = off) fixed (ulong* pl = len)
                NativeMethods.Check(_gpu, NativeMethods.alacgpu_decode_all(_gpu, p, total, po, pl, null), "alacgpu_decode_all");
            foreach (var h in _pins) h.Free();
            _pins.Clear();
            return (pcm, off, len);
        }

        public void Dispose()
        {
            foreach (var h in _pins) h.Free();
            _pins.Clear();
            _gpu?.Dispose();
        }
    }
}
