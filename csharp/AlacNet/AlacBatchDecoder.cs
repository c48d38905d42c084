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
    /// <summary>All PCM of one AlacBatchDecoder.DecodeAll: interleaved little-endian, track t at
    /// [Offset[t], Offset[t] + Length[t]) of one page-locked unmanaged buffer.</summary>
    public sealed unsafe class AlacBatchPcm : IDisposable
    {
        private IntPtr _buf;
        public ulong TotalBytes { get; }
        public ulong[] Offset { get; }
        public ulong[] Length { get; }

        internal AlacBatchPcm(IntPtr buf, ulong total, ulong[] off, ulong[] len)
        {
            _buf = buf; TotalBytes = total; Offset = off; Length = len;
        }

        /// <summary>Track t as a read-only stream over the unmanaged buffer (no copy).</summary>
        public UnmanagedMemoryStream OpenTrack(int t)
        {
            if (_buf == IntPtr.Zero) throw new ObjectDisposedException(nameof(AlacBatchPcm));
            return new UnmanagedMemoryStream((byte*)_buf + Offset[t], (long)Length[t], (long)Length[t], FileAccess.Read);
        }

        /// <summary>Copy part of track t into a managed array (a caller that wants byte[] pulls it in pieces).</summary>
        public int Read(int t, ulong trackOffset, byte[] dst, int dstOffset, int count)
        {
            if (_buf == IntPtr.Zero) throw new ObjectDisposedException(nameof(AlacBatchPcm));
            if (trackOffset >= Length[t]) return 0;
            int n = (int)Math.Min((ulong)count, Length[t] - trackOffset);
            Marshal.Copy((IntPtr)((byte*)_buf + Offset[t] + trackOffset), dst, dstOffset, n);
            return n;
        }

        public void Dispose()
        {
            if (_buf != IntPtr.Zero) NativeMethods.alacgpu_host_free(_buf);
            _buf = IntPtr.Zero;
        }
    }

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
                var pin = GCHandle.Alloc(m4a, GCHandleType.Pinned);      // borrowed until the next DecodeAll returns
                _pins.Add(pin);
                fixed (uint* ps = sizes)
                {
                    NativeMethods.Check(_gpu, NativeMethods.alacgpu_add_track(_gpu, ref cfg, (byte*)pin.AddrOfPinnedObject(),
                        (ulong)m4a.Length, (ulong)ms.Position, ps, (uint)sizes.Length, out var id), "alacgpu_add_track");
                    return id;
                }
            }
        }

        /// <summary>Decode every frame of every added file.  The PCM of a batch does not fit a managed array
        /// (configs[3] is 44 GB; byte[] tops out near 2 GB), so it lands in page-locked unmanaged memory from
        /// alacgpu_host_alloc, which the GPU copies into directly; the result hands out one
        /// UnmanagedMemoryStream per track and frees the buffer on Dispose.</summary>
        public unsafe AlacBatchPcm DecodeAll()
        {
            NativeMethods.Check(_gpu, NativeMethods.alacgpu_total_pcm_bytes(_gpu, out var total), "alacgpu_total_pcm_bytes");
            NativeMethods.Check(_gpu, NativeMethods.alacgpu_track_count(_gpu, out var n), "alacgpu_track_count");
            NativeMethods.Check(_gpu, NativeMethods.alacgpu_host_alloc(Math.Max(total, 1), out var buf), "alacgpu_host_alloc");
            var off = new ulong[n];
            var len = new ulong[n];
            try
            {
                fixed (ulong* po = off) fixed (ulong* pl = len)
                    NativeMethods.Check(_gpu, NativeMethods.alacgpu_decode_all(_gpu, (byte*)buf, total, po, pl, null), "alacgpu_decode_all");
            }
            catch
            {
                NativeMethods.alacgpu_host_free(buf);
                throw;
            }
            finally
            {
                // the library read the last borrowed byte before alacgpu_decode_all returned (alacgpu.h,
                // alacgpu_add_track): the arrays may move again, and files added later are staged on their own
                foreach (var h in _pins) h.Free();
                _pins.Clear();
            }
            return new AlacBatchPcm(buf, total, off, len);
        }

        public void Dispose()
        {
            foreach (var h in _pins) h.Free();
            _pins.Clear();
            _gpu?.Dispose();
        }
    }
}
