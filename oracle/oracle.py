"""ctypes wrapper around oracle/libalac_oracle.so.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under alac/ imports
this module.  PARITY UNPINNED -- see alac_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libalac_oracle.so")

BUFFER_SIZE = 16384

STATUS_NAMES = {
    0: "OK", 1: "BAD_TAG", 2: "PRED_TYPE", 3: "TOO_MANY_SAMPLES", 4: "OVERRUN",
    5: "BAD_RSS", 6: "HISTORY", 7: "RUN_OVERFLOW", 8: "ORDER0_LONG", 9: "INTERNAL(gpu only)",
}


class Cfg(C.Structure):
    _fields_ = [
        ("sample_size", C.c_int32),
        ("num_channels", C.c_int32),
        ("max_samples_per_frame", C.c_int32),
        ("rice_history_mult", C.c_int32),
        ("rice_initial_history", C.c_int32),
        ("rice_kmodifier", C.c_int32),
    ]


class Stages(C.Structure):
    _fields_ = [
        ("element_channels", C.c_int32), ("n", C.c_int32), ("ub", C.c_int32), ("escape", C.c_int32),
        ("mix_shift", C.c_int32), ("mix_weight", C.c_int32),
        ("pred_type", C.c_int32 * 2), ("quant", C.c_int32 * 2), ("rice_mod", C.c_int32 * 2),
        ("order", C.c_int32 * 2), ("coef", (C.c_int32 * 32) * 2),
        ("residual", (C.c_int32 * BUFFER_SIZE) * 2),
        ("predicted", (C.c_int32 * BUFFER_SIZE) * 2),
        ("shift", (C.c_int32 * BUFFER_SIZE) * 2),
        ("bits_consumed", C.c_int64),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "alac_oracle.c")
    hdr = os.path.join(_HERE, "alac_oracle.h")
    newest = max(os.path.getmtime(src), os.path.getmtime(hdr))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < newest:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libalac_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.alac_oracle_decode_frame.restype = C.c_int
        L.alac_oracle_decode_frame.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_size_t, C.c_void_p,
                                               C.c_size_t, C.POINTER(C.c_int), C.c_void_p]
        L.alac_oracle_read_frame.restype = C.c_int
        L.alac_oracle_read_frame.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_size_t, C.c_void_p,
                                             C.c_size_t, C.POINTER(C.c_int)]
        L.alac_oracle_decode_track.restype = C.c_int64
        L.alac_oracle_decode_track.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_size_t, C.c_void_p,
                                               C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.alac_oracle_set_info.restype = C.c_int
        L.alac_oracle_set_info.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(Cfg)]
        L.alac_oracle_clz.restype = C.c_int
        L.alac_oracle_clz.argtypes = [C.c_int32]
        _lib = L
    return _lib


def make_cfg(sample_size=16, num_channels=2, max_samples_per_frame=4096, rice_history_mult=40,
             rice_initial_history=10, rice_kmodifier=14) -> Cfg:
    return Cfg(sample_size, num_channels, max_samples_per_frame, rice_history_mult,
               rice_initial_history, rice_kmodifier)


def cfg_from(obj) -> Cfg:
    """Accept any object with the cookie fields as attributes (e.g. alacgen.TrackCfg)."""
    return make_cfg(obj.sample_size, obj.num_channels, obj.max_samples_per_frame,
                    obj.rice_history_mult, obj.rice_initial_history, obj.rice_kmodifier)


def read_frame(cfg: Cfg, frame: bytes):
    """AlacContext.Read for one frame -> (pcm bytes, status)."""
    buf = (C.c_uint8 * 65536)()
    st = C.c_int(0)
    src = (C.c_uint8 * max(1, len(frame))).from_buffer_copy(frame if frame else b"\0")
    n = lib().alac_oracle_read_frame(C.byref(cfg), src, len(frame), buf, 65536, C.byref(st))
    return bytes(buf[:n]), st.value


def decode_frame_stages(cfg: Cfg, frame: bytes):
    """-> (ints as the reference's outbuffer, outputsize, status, Stages)."""
    out = np.zeros(1024 * 80 + 8, dtype=np.int32)
    st = C.c_int(0)
    stages = Stages()
    src = (C.c_uint8 * max(1, len(frame))).from_buffer_copy(frame if frame else b"\0")
    size = lib().alac_oracle_decode_frame(C.byref(cfg), src, len(frame), out.ctypes.data, out.shape[0],
                                          C.byref(st), C.byref(stages))
    return out, size, st.value, stages


def decode_track(cfg: Cfg, mdat: bytes, stsz: np.ndarray, pcm_cap: int | None = None):
    """-> (pcm bytes, per-frame status int32[], per-frame byte counts uint32[])."""
    stsz = np.ascontiguousarray(stsz, dtype=np.uint32)
    nf = stsz.shape[0]
    if pcm_cap is None:
        pcm_cap = nf * 65536
    pcm = np.empty(max(1, pcm_cap), dtype=np.uint8)
    status = np.zeros(max(1, nf), dtype=np.int32)
    fbytes = np.zeros(max(1, nf), dtype=np.uint32)
    src = np.frombuffer(mdat, dtype=np.uint8) if len(mdat) else np.zeros(1, dtype=np.uint8)
    total = lib().alac_oracle_decode_track(C.byref(cfg), src.ctypes.data, len(mdat), stsz.ctypes.data, nf,
                                           pcm.ctypes.data, pcm_cap, status.ctypes.data, fbytes.ctypes.data)
    if total < 0:
        raise RuntimeError("oracle: pcm capacity too small or unsupported sample size")
    return pcm[:total].tobytes(), status[:nf], fbytes[:nf]
