"""BatchDecoder: thin Python handle on an alacgpu context (one C-ABI call per method).

The arithmetic all happens in libalacgpu.so's CUDA kernels; this class only
marshals arguments.  Mirrors what the C# AlacContext does with the same calls
(csharp/AlacNet/AlacContext.cs).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


class PinnedBuffer:
    """Page-locked host memory from alacgpu_host_alloc, viewed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        L = N.load()
        p = C.c_void_p()
        rc = L.alacgpu_host_alloc(max(1, nbytes), C.byref(p))
        if rc != N.OK:
            raise N.AlacGpuError(rc, "alacgpu_host_alloc")
        self._p = p
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(1, nbytes)).from_address(p.value))[:nbytes]

    @property
    def ptr(self) -> int:
        return self._p.value

    def free(self):
        if self._p is not None and self._p.value:
            self.array = None
            N.load().alacgpu_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _cfg_struct(cfg) -> N.TrackCfg:
    return N.TrackCfg(cfg.sample_size, cfg.num_channels, cfg.max_samples_per_frame, cfg.rice_history_mult,
                      cfg.rice_initial_history, cfg.rice_kmodifier, getattr(cfg, "sample_rate", 0))


class BatchDecoder:
    def __init__(self, devices=None, chunk_frames: int = 0, entropy_lanes: int = 0, flags: int = 0):
        self._L = N.load()
        self._h = C.c_void_p()
        opts = N.Opts(C.sizeof(N.Opts), flags, chunk_frames, entropy_lanes)
        if devices:
            ids = (C.c_int32 * len(devices))(*devices)
            rc = self._L.alacgpu_create(ids, len(devices), C.byref(opts), C.byref(self._h))
        else:
            rc = self._L.alacgpu_create(None, 0, C.byref(opts), C.byref(self._h))
        if rc != N.OK:
            self._h = C.c_void_p()
            raise N.AlacGpuError(rc, "alacgpu_create", self._L.alacgpu_strerror(rc).decode())
        self._keep = []       # arrays borrowed by the library until prepare()
        self.n_frames = 0
        self.total_pcm = 0

    # -- lifecycle -----------------------------------------------------------
    def close(self):
        if self._h:
            self._L.alacgpu_destroy(self._h)
            self._h = C.c_void_p()
        self._keep = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != N.OK:
            raise N.AlacGpuError(rc, what, self._L.alacgpu_last_error(self._h).decode())

    # -- tracks ----------------------------------------------------------------
    def add_track(self, cfg, mdat, stsz, first_frame_offset: int = 0, mdat_len: int | None = None) -> int:
        """mdat: bytes / numpy uint8 array / (ptr, len) from a PinnedBuffer."""
        stsz = np.ascontiguousarray(stsz, dtype=np.uint32)
        if isinstance(mdat, PinnedBuffer):
            ptr, ln = mdat.ptr, mdat.nbytes if mdat_len is None else mdat_len
            self._keep.append(mdat)
        else:
            arr = np.frombuffer(mdat, dtype=np.uint8) if not isinstance(mdat, np.ndarray) else mdat
            arr = np.ascontiguousarray(arr, dtype=np.uint8)
            ptr, ln = (arr.ctypes.data if arr.size else None), arr.size if mdat_len is None else mdat_len
            self._keep.append(arr)
        self._keep.append(stsz)
        tid = C.c_int32(-1)
        c = _cfg_struct(cfg)
        rc = self._L.alacgpu_add_track(self._h, C.byref(c), ptr, ln, first_frame_offset,
                                       stsz.ctypes.data if stsz.size else None, stsz.size, C.byref(tid))
        self._check(rc, "alacgpu_add_track")
        self.n_frames += int(stsz.size)
        return tid.value

    def add_track_offsets(self, cfg, file_bytes, frame_offsets, stsz) -> int:
        """frames at explicit byte offsets of `file_bytes` (chunked / gapped containers)."""
        stsz = np.ascontiguousarray(stsz, dtype=np.uint32)
        offs = np.ascontiguousarray(frame_offsets, dtype=np.uint64)
        assert offs.size == stsz.size
        arr = np.frombuffer(file_bytes, dtype=np.uint8) if not isinstance(file_bytes, np.ndarray) else file_bytes
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        self._keep += [arr, stsz, offs]
        tid = C.c_int32(-1)
        c = _cfg_struct(cfg)
        rc = self._L.alacgpu_add_track_offsets(self._h, C.byref(c), arr.ctypes.data if arr.size else None, arr.size,
                                               offs.ctypes.data if offs.size else None,
                                               stsz.ctypes.data if stsz.size else None, stsz.size, C.byref(tid))
        self._check(rc, "alacgpu_add_track_offsets")
        self.n_frames += int(stsz.size)
        return tid.value

    def clear(self):
        self._check(self._L.alacgpu_clear_tracks(self._h), "alacgpu_clear_tracks")
        self._keep = []
        self.n_frames = 0

    # -- decode ------------------------------------------------------------------
    def total_pcm_bytes(self) -> int:
        total = C.c_uint64(0)
        self._check(self._L.alacgpu_total_pcm_bytes(self._h, C.byref(total)), "alacgpu_total_pcm_bytes")
        self.total_pcm = total.value
        return total.value

    def prepare(self) -> int:
        """Optional: stage every track's mdat in HBM and run the header pre-pass."""
        total = C.c_uint64(0)
        self._check(self._L.alacgpu_prepare(self._h, C.byref(total)), "alacgpu_prepare")
        self.total_pcm = total.value
        return total.value

    def reindex(self):
        self._check(self._L.alacgpu_reindex(self._h), "alacgpu_reindex")

    def track_count(self) -> int:
        n = C.c_int32(0)
        self._check(self._L.alacgpu_track_count(self._h, C.byref(n)), "alacgpu_track_count")
        return n.value

    def decode_all(self, dst=None, want_status: bool = True):
        """dst: None (allocate numpy), a PinnedBuffer / numpy array, or False to keep
        the PCM device-resident.  -> (dst array or None, track_off, track_len, status)."""
        total = self.total_pcm_bytes()
        nt = self.track_count()
        off = np.zeros(max(1, nt), dtype=np.uint64)
        ln = np.zeros(max(1, nt), dtype=np.uint64)
        status = np.zeros(max(1, self.n_frames), dtype=np.int32) if want_status else None
        if dst is False:
            ptr, cap, arr = None, 0, None
        else:
            if dst is None:
                arr = np.empty(max(1, total), dtype=np.uint8)
            elif isinstance(dst, PinnedBuffer):
                arr = dst.array
            else:
                arr = dst
            ptr, cap = arr.ctypes.data, arr.size
        rc = self._L.alacgpu_decode_all(self._h, ptr, cap, off.ctypes.data, ln.ctypes.data,
                                        status.ctypes.data if status is not None else None)
        self._check(rc, "alacgpu_decode_all")
        return (arr[:total] if arr is not None else None, off[:nt], ln[:nt],
                status[:self.n_frames] if status is not None else None)

    def read_frame(self, track: int, frame_idx: int, buf: np.ndarray | None = None) -> bytes:
        if buf is None:
            buf = np.empty(65536, dtype=np.uint8)
        n = C.c_uint32(0)
        rc = self._L.alacgpu_read_frame(self._h, track, frame_idx, buf.ctypes.data, buf.size, C.byref(n))
        self._check(rc, "alacgpu_read_frame")
        return buf[:n.value].tobytes()

    def frame_count(self, track: int) -> int:
        n = C.c_uint32(0)
        self._check(self._L.alacgpu_frame_count(self._h, track, C.byref(n)), "alacgpu_frame_count")
        return n.value

    def frame_samples(self, track: int, frame_idx: int) -> int:
        n = C.c_uint32(0)
        self._check(self._L.alacgpu_frame_samples(self._h, track, frame_idx, C.byref(n)), "alacgpu_frame_samples")
        return n.value

    def frame_status(self, track: int, frame_idx: int) -> int:
        n = C.c_int32(0)
        self._check(self._L.alacgpu_frame_status(self._h, track, frame_idx, C.byref(n)), "alacgpu_frame_status")
        return n.value

    def track_pcm_bytes(self, track: int):
        o, l = C.c_uint64(0), C.c_uint64(0)
        self._check(self._L.alacgpu_track_pcm_bytes(self._h, track, C.byref(o), C.byref(l)), "alacgpu_track_pcm_bytes")
        return o.value, l.value

    def timing(self) -> dict:
        t = N.Timing()
        self._check(self._L.alacgpu_get_timing(self._h, C.byref(t)), "alacgpu_get_timing")
        return t.as_dict()

    def checksum(self, off: int = 0, length: int | None = None) -> int:
        s = C.c_uint64(0)
        if length is None:
            length = self.total_pcm - off
        self._check(self._L.alacgpu_pcm_checksum(self._h, off, length, C.byref(s)), "alacgpu_pcm_checksum")
        return s.value

    def device_pcm(self, slot: int = 0):
        p, o, l = C.c_void_p(), C.c_uint64(0), C.c_uint64(0)
        self._check(self._L.alacgpu_device_pcm(self._h, slot, C.byref(p), C.byref(o), C.byref(l)), "alacgpu_device_pcm")
        return p.value, o.value, l.value


def host_checksum(data, first_word: int = 0) -> int:
    """alacgpu_pcm_checksum's formula over host bytes (for tests): sum of
    w_j * (2*(first_word+j)+1) mod 2^64 over little-endian 8-byte words."""
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    pad = (-a.size) % 8
    if pad:
        a = np.concatenate([a, np.zeros(pad, dtype=np.uint8)])
    w = a.view("<u8")
    j = np.arange(first_word, first_word + w.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(np.sum(w * (j * np.uint64(2) + np.uint64(1)), dtype=np.uint64))


def plan_partition(frame_sizes, n_parts: int) -> np.ndarray:
    sizes = np.ascontiguousarray(frame_sizes, dtype=np.uint32)
    cut = np.zeros(n_parts + 1, dtype=np.uint64)
    rc = N.load().alacgpu_plan_partition(sizes.ctypes.data if sizes.size else None, sizes.size, n_parts, cut.ctypes.data)
    if rc != N.OK:
        raise N.AlacGpuError(rc, "alacgpu_plan_partition")
    return cut
