// AlacGpuNative.cs -- P/Invoke mirror of include/alacgpu.h (ABI version 1), netstandard2.0.
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no dotnet / mono).  It is the binding a
// maintainer of teekay/ALAC.NET adds next to ALACDecoder/AlacContext.cs; every entry
// point and struct below is declared 1:1 with the C header (same order, same widths) and
// the header's layout is exercised from C/ctypes by tests/ (see INTEGRATION.md).
//
// Library resolution: netstandard2.0 has no NativeLibrary API, so "alacgpu" resolves
// through the default probing rules (libalacgpu.so next to the assembly, or LD_LIBRARY_PATH).
using System;
using System.Runtime.InteropServices;

namespace ALACdotNET.Decoder.Gpu
{
    internal enum AlacGpuStatus : int
    {
        Ok = 0, InvalidArg = -1, NoDevice = -2, Cuda = -3, OutOfMemory = -4,
        Unsupported = -5, Capacity = -6, State = -7, Range = -8
    }

    internal enum AlacGpuFrameStatus : int
    {
        Ok = 0, BadTag = 1, PredType = 2, TooManySamples = 3, Overrun = 4,
        BadRss = 5, History = 6, RunOverflow = 7, Order0Long = 8, Internal = 9
    }

    /// <summary>alacgpu_opts.flags (ALACGPU_FLAG_* in alacgpu.h); every combination yields the same bytes.</summary>
    [Flags]
    internal enum AlacGpuFlags : uint
    {
        None = 0, KeepDevicePcm = 0x1, NoFusion = 0x2, NoPackFusion = 0x4, NoZeroCopy = 0x8,
        NoQuadLpc = 0x10, ForcePackFusion = 0x20, NoFrameLanes = 0x40, ForceFrameLanes = 0x80
    }

    [StructLayout(LayoutKind.Sequential)]
    internal struct AlacGpuOpts
    {
        public uint StructSize;
        public uint Flags;
        public uint ChunkFrames;
        public uint EntropyLanes;
        public uint Reserved0, Reserved1, Reserved2, Reserved3;
    }

    /// <summary>The 'alac' cookie fields AlacFile.SetInfo keeps (AlacFile.cs:72-92).</summary>
    [StructLayout(LayoutKind.Sequential)]
    internal struct AlacGpuTrackCfg
    {
        public int SampleSize;
        public int NumChannels;
        public int MaxSamplesPerFrame;
        public int RiceHistoryMult;
        public int RiceInitialHistory;
        public int RiceKModifier;
        public int SampleRate;
    }

    [StructLayout(LayoutKind.Sequential)]
    internal struct AlacGpuTiming
    {
        public float IndexMs, EntropyMs, LpcMs, StereoMs, KernelsMs, H2dMs, D2hMs, TotalMs;
        public uint KernelLaunches, Chunks;
        public ulong CompressedBytes, PcmBytes, Samples;
        public uint InternalRetries, Reserved;
    }

    internal sealed class AlacGpuHandle : SafeHandle
    {
        public AlacGpuHandle() : base(IntPtr.Zero, true) { }
        public override bool IsInvalid => handle == IntPtr.Zero;
        protected override bool ReleaseHandle() => NativeMethods.alacgpu_destroy(handle) == 0;
    }

    internal static unsafe class NativeMethods
    {
        private const string Lib = "alacgpu";
        private const CallingConvention Cc = CallingConvention.Cdecl;

        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_create(int* deviceIds, int nDevices, ref AlacGpuOpts opts, out AlacGpuHandle ctx);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_destroy(IntPtr ctx);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_add_track(AlacGpuHandle ctx, ref AlacGpuTrackCfg cfg, byte* mdat, ulong mdatLen, ulong firstFrameOffset, uint* frameSizes, uint nFrames, out int trackId);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_add_track_offsets(AlacGpuHandle ctx, ref AlacGpuTrackCfg cfg, byte* file, ulong fileLen, ulong* frameOffsets, uint* frameSizes, uint nFrames, out int trackId);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_clear_tracks(AlacGpuHandle ctx);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_total_pcm_bytes(AlacGpuHandle ctx, out ulong total);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_prepare(AlacGpuHandle ctx, out ulong total);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_reindex(AlacGpuHandle ctx);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_decode_all(AlacGpuHandle ctx, byte* pcmDst, ulong cap, ulong* trackPcmOff, ulong* trackPcmLen, int* frameStatus);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_read_frame(AlacGpuHandle ctx, int track, uint frameIdx, byte* dst, uint cap, out uint bytesOut);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_track_count(AlacGpuHandle ctx, out int nTracks);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_frame_count(AlacGpuHandle ctx, int track, out uint nFrames);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_frame_samples(AlacGpuHandle ctx, int track, uint frameIdx, out uint nSamples);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_track_pcm_bytes(AlacGpuHandle ctx, int track, out ulong off, out ulong len);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_frame_status(AlacGpuHandle ctx, int track, uint frameIdx, out int status);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_get_timing(AlacGpuHandle ctx, out AlacGpuTiming timing);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_device_pcm(AlacGpuHandle ctx, int devSlot, out IntPtr dptr, out ulong shardOff, out ulong shardLen);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_pcm_checksum(AlacGpuHandle ctx, ulong off, ulong len, out ulong sum);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_host_alloc(ulong bytes, out IntPtr ptr);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_host_free(IntPtr ptr);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_plan_partition(uint* frameSizes, ulong nFrames, int nParts, ulong* cut);
        [DllImport(Lib, CallingConvention = Cc)] public static extern IntPtr alacgpu_strerror(int status);
        [DllImport(Lib, CallingConvention = Cc)] public static extern IntPtr alacgpu_last_error(AlacGpuHandle ctx);
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_abi_version();
        [DllImport(Lib, CallingConvention = Cc)] public static extern int alacgpu_device_count(out int n);

        public static void Check(AlacGpuHandle ctx, int rc, string what)
        {
            if (rc == 0) return;
            var msg = Marshal.PtrToStringAnsi(alacgpu_strerror(rc));
            var detail = ctx != null && !ctx.IsInvalid ? Marshal.PtrToStringAnsi(alacgpu_last_error(ctx)) : "";
            if ((AlacGpuStatus)rc == AlacGpuStatus.Unsupported)
                throw new Exception("FIXME: unimplemented sample size (" + detail + ")");      // AlacFile.cs:574,715
            throw new InvalidOperationException(what + ": " + msg + (string.IsNullOrEmpty(detail) ? "" : " (" + detail + ")"));
        }
    }
}
