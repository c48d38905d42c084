"""ctypes binding of libalacgpu.so (include/alacgpu.h), 1:1 with the C ABI.

This is the same boundary the C# host P/Invokes (csharp/AlacNet/AlacGpuNative.cs)
and the C++ host mirror links against.  There is no fallback: if the library is
missing or no CUDA device is usable, loading / alacgpu_create fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ALACGPU_LIB selects another build of the same ABI (tests: libalacgpu_checked.so, the bounds-asserting build)
LIB_PATH = os.environ.get("ALACGPU_LIB") or os.path.join(_HERE, "libalacgpu.so")

OK = 0
ERR_NAMES = {
    0: "OK", -1: "INVALID_ARG", -2: "NO_DEVICE", -3: "CUDA", -4: "OUT_OF_MEMORY",
    -5: "UNSUPPORTED", -6: "CAPACITY", -7: "STATE", -8: "RANGE",
}
FLAG_KEEP_DEVICE_PCM = 0x1
FLAG_NO_FUSION = 0x2
FLAG_NO_PACK_FUSION = 0x4
FLAG_NO_ZERO_COPY = 0x8
FLAG_NO_QUAD_LPC = 0x10
FLAG_FORCE_PACK_FUSION = 0x20
FLAG_NO_FRAME_LANES = 0x40
FLAG_FORCE_FRAME_LANES = 0x80

# every symbol include/alacgpu.h declares (tests check the export table against this)
EXPORTS = [
    "alacgpu_create", "alacgpu_destroy", "alacgpu_add_track", "alacgpu_add_track_offsets", "alacgpu_clear_tracks",
    "alacgpu_total_pcm_bytes", "alacgpu_prepare", "alacgpu_reindex", "alacgpu_decode_all", "alacgpu_read_frame", "alacgpu_track_count",
    "alacgpu_frame_count", "alacgpu_frame_samples", "alacgpu_track_pcm_bytes",
    "alacgpu_frame_status", "alacgpu_get_timing", "alacgpu_device_pcm", "alacgpu_pcm_checksum",
    "alacgpu_host_alloc", "alacgpu_host_free", "alacgpu_plan_partition", "alacgpu_strerror",
    "alacgpu_last_error", "alacgpu_abi_version", "alacgpu_device_count",
]


class Opts(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("flags", C.c_uint32), ("chunk_frames", C.c_uint32),
        ("entropy_lanes", C.c_uint32), ("reserved", C.c_uint32 * 4),
    ]


class TrackCfg(C.Structure):
    _fields_ = [
        ("sample_size", C.c_int32), ("num_channels", C.c_int32), ("max_samples_per_frame", C.c_int32),
        ("rice_history_mult", C.c_int32), ("rice_initial_history", C.c_int32),
        ("rice_kmodifier", C.c_int32), ("sample_rate", C.c_int32),
    ]


class Timing(C.Structure):
    _fields_ = [
        ("index_ms", C.c_float), ("entropy_ms", C.c_float), ("lpc_ms", C.c_float), ("stereo_ms", C.c_float),
        ("kernels_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float),
        ("kernel_launches", C.c_uint32), ("chunks", C.c_uint32), ("compressed_bytes", C.c_uint64),
        ("pcm_bytes", C.c_uint64), ("samples", C.c_uint64), ("internal_retries", C.c_uint32), ("reserved", C.c_uint32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class AlacGpuError(RuntimeError):
    def __init__(self, code: int, what: str, detail: str = ""):
        self.code = code
        super().__init__(f"{what}: {ERR_NAMES.get(code, code)}" + (f" ({detail})" if detail else ""))


_lib = None


def load() -> C.CDLL:
    """dlopen libalacgpu.so (building it is __graft_entry__.build()'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m alac.net_b200.build` "
                          "(there is no CPU fallback for the decode path)")
    L = C.CDLL(LIB_PATH)
    vp, u64p, u32p, i32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    L.alacgpu_create.argtypes = [i32p, C.c_int32, C.POINTER(Opts), C.POINTER(vp)]
    L.alacgpu_destroy.argtypes = [vp]
    L.alacgpu_add_track.argtypes = [vp, C.POINTER(TrackCfg), vp, C.c_uint64, C.c_uint64, vp, C.c_uint32, i32p]
    L.alacgpu_add_track_offsets.argtypes = [vp, C.POINTER(TrackCfg), vp, C.c_uint64, vp, vp, C.c_uint32, i32p]
    L.alacgpu_clear_tracks.argtypes = [vp]
    L.alacgpu_total_pcm_bytes.argtypes = [vp, u64p]
    L.alacgpu_prepare.argtypes = [vp, u64p]
    L.alacgpu_reindex.argtypes = [vp]
    L.alacgpu_decode_all.argtypes = [vp, vp, C.c_uint64, vp, vp, vp]
    L.alacgpu_read_frame.argtypes = [vp, C.c_int32, C.c_uint32, vp, C.c_uint32, u32p]
    L.alacgpu_track_count.argtypes = [vp, i32p]
    L.alacgpu_frame_count.argtypes = [vp, C.c_int32, u32p]
    L.alacgpu_frame_samples.argtypes = [vp, C.c_int32, C.c_uint32, u32p]
    L.alacgpu_track_pcm_bytes.argtypes = [vp, C.c_int32, u64p, u64p]
    L.alacgpu_frame_status.argtypes = [vp, C.c_int32, C.c_uint32, i32p]
    L.alacgpu_get_timing.argtypes = [vp, C.POINTER(Timing)]
    L.alacgpu_device_pcm.argtypes = [vp, C.c_int32, C.POINTER(vp), u64p, u64p]
    L.alacgpu_pcm_checksum.argtypes = [vp, C.c_uint64, C.c_uint64, u64p]
    L.alacgpu_host_alloc.argtypes = [C.c_uint64, C.POINTER(vp)]
    L.alacgpu_host_free.argtypes = [vp]
    L.alacgpu_plan_partition.argtypes = [vp, C.c_uint64, C.c_int32, vp]
    L.alacgpu_strerror.argtypes = [C.c_int32]
    L.alacgpu_strerror.restype = C.c_char_p
    L.alacgpu_last_error.argtypes = [vp]
    L.alacgpu_last_error.restype = C.c_char_p
    L.alacgpu_abi_version.argtypes = []
    L.alacgpu_device_count.argtypes = [i32p]
    for name in EXPORTS:
        if name not in ("alacgpu_strerror", "alacgpu_last_error"):
            getattr(L, name).restype = C.c_int32
    _lib = L
    return L
