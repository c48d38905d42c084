"""ctypes view of libalacnet_host.so -- the C++ mirror of the reference's AlacContext /
QtMovieT / ALACFileReader (alac/net_b200/host/alacnet.hpp).  Same method names as the C#
classes so tests read like code written against the reference:

    ctx = AlacContext(m4a_bytes)            # throws IOException on a bad header
    n = ctx.Read(buffer)                    # one frame per call, 0 at the end
    ctx.SetPosition(sample); ctx.LastSampleNumber

All decoding happens in libalacgpu.so (which libalacnet_host.so links against).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libalacnet_host.so")


class IOException(IOError):
    pass


class ArgumentException(ValueError):
    pass


class DecoderException(RuntimeError):
    pass


_ERR = {-101: IOException, -102: ArgumentException, -103: DecoderException}
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -m alac.net_b200.build`")
        L = C.CDLL(LIB_PATH)
        L.alacnet_last_error.restype = C.c_char_p
        L.alacnet_context_open.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
        L.alacnet_context_close.argtypes = [C.c_void_p]
        L.alacnet_context_close.restype = None
        L.alacnet_context_read.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
        L.alacnet_context_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 5
        L.alacnet_context_set_position.argtypes = [C.c_void_p, C.c_int64]
        L.alacnet_context_last_sample_number.argtypes = [C.c_void_p]
        L.alacnet_context_mdat_offset.argtypes = [C.c_void_p]
        L.alacnet_context_mdat_offset.restype = C.c_int64
        L.alacnet_context_frame_count.argtypes = [C.c_void_p]
        L.alacnet_demux.argtypes = [C.c_void_p, C.c_uint64] + [C.POINTER(C.c_int)] * 4 + [C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_void_p]
        L.alacnet_reader_open.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
        L.alacnet_reader_close.argtypes = [C.c_void_p]
        L.alacnet_reader_close.restype = None
        L.alacnet_reader_read.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.alacnet_reader_length.argtypes = [C.c_void_p]
        L.alacnet_reader_length.restype = C.c_int64
        L.alacnet_reader_position.argtypes = [C.c_void_p]
        L.alacnet_reader_position.restype = C.c_int64
        L.alacnet_reader_set_position.argtypes = [C.c_void_p, C.c_int64]
        L.alacnet_reader_format.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise _ERR.get(rc, DecoderException)(load().alacnet_last_error().decode())


def _as_array(data) -> np.ndarray:
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    return np.ascontiguousarray(a, dtype=np.uint8)


def iso_demux(m4a) -> dict:
    """alacnet::IsoDemux: tolerant ISO-BMFF demux -> cookie + per-frame (offset, size, duration)."""
    from . import _native as N
    L = load()
    a = _as_array(m4a)
    cap = max(1, a.size // 4)
    cfg = N.TrackCfg()
    offs = np.zeros(cap, dtype=np.uint64)
    sizes = np.zeros(cap, dtype=np.uint32)
    durs = np.zeros(cap, dtype=np.uint32)
    n = C.c_uint32(0)
    total = C.c_uint64(0)
    L.alacnet_iso_demux.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    rc = L.alacnet_iso_demux(a.ctypes.data, a.size, C.byref(cfg), offs.ctypes.data, sizes.ctypes.data, durs.ctypes.data,
                             cap, C.byref(n), C.byref(total))
    if rc != 0:
        raise IOException("no ALAC audio track found")
    k = n.value
    return {"cfg": cfg, "offsets": offs[:k].copy(), "stsz": sizes[:k].copy(), "durations": durs[:k].copy(),
            "total_samples": total.value}


def wav_header(rate: int, bits: int, channels: int, pcm_bytes: int) -> bytes:
    L = load()
    buf = (C.c_uint8 * 44)()
    L.alacnet_wav_header.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p]
    L.alacnet_wav_header(rate, bits, channels, pcm_bytes, buf)
    return bytes(buf)


def demux(m4a) -> dict:
    """QtMovieT.ReadHeader over an in-memory file (host only, no GPU)."""
    L = load()
    a = _as_array(m4a)
    status, ss, ch, rate, nf = (C.c_int(0) for _ in range(5))
    mdat = C.c_int64(0)
    cap = max(1, a.size // 4)
    sizes = np.zeros(cap, dtype=np.uint32)
    cd = np.zeros(48, dtype=np.int32)
    _check(L.alacnet_demux(a.ctypes.data, a.size, C.byref(status), C.byref(ss), C.byref(ch), C.byref(rate),
                           C.byref(mdat), C.byref(nf), sizes.ctypes.data, cap, cd.ctypes.data))
    return {"status": status.value, "sample_size": ss.value, "num_channels": ch.value, "sample_rate": rate.value,
            "mdat_pos": mdat.value, "stsz": sizes[:nf.value].copy(), "codec_data": cd}


class AlacContext:
    def __init__(self, m4a, device: int = 0):
        self._L = load()
        self._file = _as_array(m4a)          # must outlive the native context
        self._h = C.c_void_p()
        _check(self._L.alacnet_context_open(self._file.ctypes.data, self._file.size, device, C.byref(self._h)))

    def Read(self, buffer: np.ndarray) -> int:
        n = C.c_int(0)
        _check(self._L.alacnet_context_read(self._h, buffer.ctypes.data, buffer.size, C.byref(n)))
        return n.value

    def _info(self):
        v = [C.c_int(0) for _ in range(5)]
        _check(self._L.alacnet_context_info(self._h, *[C.byref(x) for x in v]))
        return [x.value for x in v]

    def GetSampleRate(self): return self._info()[0]
    def GetNumChannels(self): return self._info()[1]
    def GetBitsPerSample(self): return self._info()[2]
    def GetBytesPerSample(self): return self._info()[3]
    def GetNumSamples(self): return self._info()[4]

    @property
    def LastSampleNumber(self) -> int:
        return self._L.alacnet_context_last_sample_number(self._h)

    def SetPosition(self, position: int):
        _check(self._L.alacnet_context_set_position(self._h, position))

    def Dispose(self):
        if self._h:
            self._L.alacnet_context_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.Dispose()
        except Exception:
            pass


class ALACFileReader:
    def __init__(self, m4a, device: int = 0):
        self._L = load()
        self._file = _as_array(m4a)
        self._h = C.c_void_p()
        _check(self._L.alacnet_reader_open(self._file.ctypes.data, self._file.size, device, C.byref(self._h)))

    @property
    def Length(self) -> int:
        return self._L.alacnet_reader_length(self._h)

    @property
    def Position(self) -> int:
        return self._L.alacnet_reader_position(self._h)

    @Position.setter
    def Position(self, value: int):
        _check(self._L.alacnet_reader_set_position(self._h, value))

    @property
    def WaveFormat(self) -> dict:
        v = [C.c_int(0) for _ in range(4)]
        self._L.alacnet_reader_format(self._h, *[C.byref(x) for x in v])
        return dict(zip(("SampleRate", "BitsPerSample", "Channels", "BlockAlign"), (x.value for x in v)))

    def Read(self, buffer: np.ndarray, offset: int, count: int) -> int:
        n = C.c_int(0)
        _check(self._L.alacnet_reader_read(self._h, buffer.ctypes.data, offset, count, C.byref(n)))
        return n.value

    def Dispose(self):
        if self._h:
            self._L.alacnet_reader_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.Dispose()
        except Exception:
            pass
