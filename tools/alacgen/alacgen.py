"""Synthetic ALAC corpus generator: signal -> frame parameters -> encoder -> .m4a.

Test/bench input generator only (no corpus is available offline and the
reference ships none).  The frame encoder is tools/alacgen/alacenc.c (built to
libalacenc.so); this module adds

* seeded PCM signals (sinusoids + AR(1) noise + silence gaps + full-scale
  bursts; 24-bit tracks get spans with zeroed low bytes for the wasted-bits
  path) -- SURVEY.md section 8(d) "value distributions / seeds",
* per-frame encoder parameters drawn from the same PRNG,
* an .m4a muxer that writes exactly the container grammar the reference's
  demuxer accepts (QTMovieT.cs:51-751),
* the five BASELINE.json configs as `make_config(k, ...)`.

Nothing here imports oracle/ or the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libalacenc.so")

FRAME_DTYPE = np.dtype(
    [
        ("tag", "<i4"),
        ("n", "<i4"),
        ("force_hassize", "<i4"),
        ("ub", "<i4"),
        ("escape", "<i4"),
        ("mix_shift", "<i4"),
        ("mix_weight", "<i4"),
        ("pred_type", "<i4", (2,)),
        ("quant", "<i4", (2,)),
        ("rice_mod", "<i4", (2,)),
        ("order", "<i4", (2,)),
        ("coef", "<i4", (2, 32)),
        ("adapt_passes", "<i4"),
        ("end_tag", "<i4"),
    ]
)


class EncCfg(C.Structure):
    _fields_ = [
        ("sample_size", C.c_int32),
        ("max_samples_per_frame", C.c_int32),
        ("rice_history_mult", C.c_int32),
        ("rice_initial_history", C.c_int32),
        ("rice_kmodifier", C.c_int32),
    ]


def build_encoder(force: bool = False) -> str:
    src = os.path.join(_HERE, "alacenc.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-fopenmp", "-fPIC", "-std=c11", "-Wall", "-shared", "-o", _LIB_PATH, src, "-lm"]
        )
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build_encoder()
        lib = C.CDLL(_LIB_PATH)
        lib.alacenc_encode_track.restype = C.c_size_t
        lib.alacenc_encode_track.argtypes = [
            C.POINTER(EncCfg), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_size_t, C.c_void_p,
        ]
        lib.alacgen_signal.restype = None
        lib.alacgen_signal.argtypes = [C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib = lib
    return _lib


@dataclass
class TrackCfg:
    """The 'alac' cookie fields (AlacFile.cs:63-93) + container facts."""
    sample_size: int = 16
    num_channels: int = 2           # CONTAINER channel count (cookie byte 33)
    max_samples_per_frame: int = 4096
    rice_history_mult: int = 40
    rice_initial_history: int = 10
    rice_kmodifier: int = 14
    sample_rate: int = 44100


@dataclass
class Track:
    cfg: TrackCfg
    mdat: bytes                     # frames back to back (the mdat payload)
    stsz: np.ndarray                # uint32 per-frame byte sizes
    frame_samples: np.ndarray       # int32 per-frame sample counts
    pcm: bytes                      # expected interleaved little-endian PCM
    frames: np.ndarray = field(repr=False, default=None)   # FRAME_DTYPE array used

    @property
    def n_frames(self) -> int:
        return int(self.stsz.shape[0])

    @property
    def n_sample_frames(self) -> int:
        return int(self.frame_samples.sum())

    @property
    def n_samples(self) -> int:
        """channel values, the unit of BASELINE.json's metric"""
        return self.n_sample_frames * self.cfg.num_channels


# --------------------------------------------------------------------------
# encoder wrapper
# --------------------------------------------------------------------------
def encode_track(cfg: TrackCfg, frames: np.ndarray, left: np.ndarray, right: np.ndarray | None):
    """frames: FRAME_DTYPE array; left/right: planar int32.  -> (mdat bytes, stsz)."""
    lib = _load()
    frames = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
    left = np.ascontiguousarray(left, dtype=np.int32)
    if right is not None:
        right = np.ascontiguousarray(right, dtype=np.int32)
    total = int(frames["n"].sum())
    assert left.shape[0] >= total
    ec = EncCfg(cfg.sample_size, cfg.max_samples_per_frame, cfg.rice_history_mult,
                cfg.rice_initial_history, cfg.rice_kmodifier)
    # worst case: escape symbols everywhere (9 + rss bits per symbol) + headers
    cap = int(total * 2 * 5 + frames.shape[0] * 256 + 1024)
    out = np.empty(cap, dtype=np.uint8)
    stsz = np.zeros(frames.shape[0], dtype=np.uint32)
    nbytes = lib.alacenc_encode_track(
        C.byref(ec), frames.ctypes.data, frames.shape[0], left.ctypes.data,
        right.ctypes.data if right is not None else None, out.ctypes.data, cap, stsz.ctypes.data,
    )
    if nbytes == 0 and frames.shape[0] > 0:
        raise RuntimeError("alacenc: output overflow")
    return out[:nbytes].tobytes(), stsz


def pcm_bytes(cfg: TrackCfg, frames: np.ndarray, left: np.ndarray, right: np.ndarray | None) -> bytes:
    """Expected decoder output: interleaved little-endian, L first.  A mono
    element in a 2-channel container yields (sample, 0) pairs
    (AlacFile.cs:534-540, :555-565); a stereo element in a 1-channel container
    yields the left channel only (overwrite pattern of AlacFile.cs:353-354)."""
    total = int(frames["n"].sum())
    nch = cfg.num_channels
    out = np.zeros((total, nch), dtype=np.int32)
    pos = 0
    for fr in frames:
        n = int(fr["n"])
        out[pos:pos + n, 0] = left[pos:pos + n]
        if nch == 2 and int(fr["tag"]) == 1:
            out[pos:pos + n, 1] = right[pos:pos + n]
        pos += n
    return pack_pcm(out, cfg.sample_size)


def pack_pcm(interleaved: np.ndarray, sample_size: int) -> bytes:
    flat = interleaved.reshape(-1)
    if sample_size == 16:
        return flat.astype("<i2").tobytes()
    u = flat.astype(np.uint32)
    b = np.empty((flat.shape[0], 3), dtype=np.uint8)
    b[:, 0] = u & 0xFF
    b[:, 1] = (u >> 8) & 0xFF
    b[:, 2] = (u >> 16) & 0xFF
    return b.tobytes()


# --------------------------------------------------------------------------
# signals
# --------------------------------------------------------------------------
def make_signal(seed: int, n: int, sample_size: int, sample_rate: int, channels: int,
                silence: bool = True, bursts: bool = True, wasted_spans: bool = False) -> np.ndarray:
    """(channels, n) planar int32 in the signed sample_size range, from the
    seeded C generator (alacgen_signal in alacenc.c; thread-count independent)."""
    out = np.zeros((channels, max(1, n)), dtype=np.int32)
    flags = (1 if silence else 0) | (2 if bursts else 0) | (4 if wasted_spans else 0)
    _load().alacgen_signal(seed & 0xFFFFFFFFFFFFFFFF, n, sample_size, sample_rate, channels, flags,
                           out.ctypes.data)
    return out[:, :n]


# --------------------------------------------------------------------------
# frame parameter policies
# --------------------------------------------------------------------------
def default_coefs(order: np.ndarray, quant: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """Apple-style initial taps scaled to the frame's quantiser, plus jitter."""
    nf = order.shape[0]
    coef = np.zeros((nf, 32), dtype=np.int64)
    den = (1 << quant.astype(np.int64))
    coef[:, 0] = (38 * den) >> 4
    coef[:, 1] = (-29 * den) >> 4
    coef[:, 2] = (-2 * den) >> 4
    coef += rng.integers(-3, 4, size=coef.shape)
    idx = np.arange(32)[None, :]
    coef[idx >= order[:, None]] = 0
    return np.clip(coef, -32768, 32767).astype(np.int32)


def make_frames(rng: np.random.Generator, cfg: TrackCfg, total: int, stereo_element: bool,
                orders=(1, 31), quants=(1, 15), rice_mods=(4, 4), escape_prob: float = 0.0,
                adapt_passes: int = 1, mix: bool = True, end_tag: bool = True,
                auto_escape: bool = True) -> np.ndarray:
    """Split `total` sample-frames into frames of max_samples_per_frame (last one
    short, with hassize) and draw the encoder parameters."""
    nmax = cfg.max_samples_per_frame
    nf = (total + nmax - 1) // nmax
    fr = np.zeros(nf, dtype=FRAME_DTYPE)
    fr["tag"] = 1 if stereo_element else 0
    fr["n"] = nmax
    if nf:
        fr["n"][-1] = total - nmax * (nf - 1)
    for c in range(2):
        fr["order"][:, c] = rng.integers(orders[0], orders[1] + 1, size=nf)
        fr["quant"][:, c] = rng.integers(quants[0], quants[1] + 1, size=nf)
        fr["rice_mod"][:, c] = rng.integers(rice_mods[0], rice_mods[1] + 1, size=nf)
        fr["coef"][:, c, :] = default_coefs(fr["order"][:, c], fr["quant"][:, c], rng)
    if stereo_element and mix:
        sh = rng.integers(0, 5, size=nf)
        fr["mix_shift"] = sh
        w = (rng.random(nf) * ((1 << sh) + 1)).astype(np.int64)   # 0 .. 2^shift (A.7 item 2)
        fr["mix_weight"] = np.minimum(w, 1 << sh)
    esc = rng.random(nf) < escape_prob
    fr["escape"] = np.where(esc, 1, -1 if auto_escape else 0)
    fr["adapt_passes"] = adapt_passes
    fr["end_tag"] = 1 if end_tag else 0
    return fr


def assign_wasted_bytes(fr: np.ndarray, x: np.ndarray, sample_size: int, rng: np.random.Generator) -> None:
    """Wasted bytes per frame (24-bit only; the 16-bit decode path ignores them,
    AlacFile.cs:338-367).  Any ub is valid for any data because the low ub*8
    bits travel verbatim in the shift planes.  Stock encoders always send the
    low byte of 24-bit audio that way (ub=1) so the predictor stays inside
    int32; here ub is drawn 0/1/2 with weights .2/.6/.2 and raised to the
    number of all-zero low bytes of the frame."""
    if sample_size != 24:
        return
    nf = fr.shape[0]
    draw = rng.choice(np.array([0, 1, 2]), size=nf, p=[0.2, 0.6, 0.2])
    pos = 0
    for i in range(nf):
        n = int(fr["n"][i])
        seg = x[:, pos:pos + n]
        pos += n
        if int(fr["escape"][i]) == 1:
            continue
        orv = int(np.bitwise_or.reduce(seg.reshape(-1)) & 0xFFFFFF) if n else 1
        ub = int(draw[i])
        if orv & 0xFF == 0:
            ub = max(ub, 1)
            if orv & 0xFF00 == 0:
                ub = 2
        fr["ub"][i] = ub


def build_track(cfg: TrackCfg, x: np.ndarray, fr: np.ndarray) -> Track:
    left = np.ascontiguousarray(x[0])
    right = np.ascontiguousarray(x[1]) if x.shape[0] > 1 else None
    mdat, stsz = encode_track(cfg, fr, left, right)
    pcm = pcm_bytes(cfg, fr, left, right)
    return Track(cfg, mdat, stsz, fr["n"].astype(np.int32).copy(), pcm, fr)


# --------------------------------------------------------------------------
# the five BASELINE.json configs
# --------------------------------------------------------------------------
SEED_BASE = 0xA1AC0000


def track_16_stereo(seed: int, seconds: float, rate: int = 44100, realistic: bool = False) -> Track:
    rng = np.random.default_rng(seed)
    cfg = TrackCfg(16, 2, 4096, 40, 10, 14, rate)
    n = int(round(seconds * rate))
    x = make_signal(seed, n, 16, rate, 2)
    if realistic:   # what a stock encoder emits: order 4/8, quant 9, pb factor 4
        fr = make_frames(rng, cfg, n, True, orders=(8, 8), quants=(9, 9))
    else:
        fr = make_frames(rng, cfg, n, True, orders=(0, 31), quants=(1, 15), rice_mods=(1, 7))
    return build_track(cfg, x, fr)


def track_24_stereo(seed: int, seconds: float, rate: int = 96000) -> Track:
    rng = np.random.default_rng(seed)
    cfg = TrackCfg(24, 2, 4096, 40, 10, 14, rate)
    n = int(round(seconds * rate))
    x = make_signal(seed, n, 24, rate, 2, wasted_spans=True)
    fr = make_frames(rng, cfg, n, True, orders=(1, 31), quants=(1, 15), rice_mods=(1, 7))
    assign_wasted_bytes(fr, x, 24, rng)
    return build_track(cfg, x, fr)


def track_16_mono_mixed(seed: int, seconds: float, rate: int = 44100) -> Track:
    """config 3: one third normal, one third uncompressed, one third
    heavy-tailed residuals (Rice escapes), i.i.d. per frame."""
    rng = np.random.default_rng(seed)
    cfg = TrackCfg(16, 1, 4096, 40, 10, 14, rate)
    n = int(round(seconds * rate))
    x = make_signal(seed, n, 16, rate, 1, bursts=False).copy()
    fr = make_frames(rng, cfg, n, False, orders=(0, 31), quants=(1, 15), rice_mods=(1, 7),
                     auto_escape=False)
    kind = rng.integers(0, 3, size=fr.shape[0])
    fr["escape"] = np.where(kind == 1, 1, 0)
    pos = 0
    for i in range(fr.shape[0]):
        nn = int(fr["n"][i])
        if kind[i] == 2:   # sparse full-scale impulses on a quiet bed -> 9-ones escapes
            seg = x[0, pos:pos + nn]
            hits = rng.random(nn) < 0.08
            seg[hits] = rng.integers(-32768, 32768, size=int(hits.sum()))
        pos += nn
    return build_track(cfg, x, fr)


def track_24_mono(seed: int, seconds: float, rate: int = 96000) -> Track:
    rng = np.random.default_rng(seed)
    cfg = TrackCfg(24, 1, 4096, 40, 10, 14, rate)
    n = int(round(seconds * rate))
    x = make_signal(seed, n, 24, rate, 1, wasted_spans=True)
    fr = make_frames(rng, cfg, n, False, orders=(1, 31), quants=(1, 15), rice_mods=(1, 7))
    assign_wasted_bytes(fr, x, 24, rng)
    return build_track(cfg, x, fr)


def corpus_track(i: int, seconds: float) -> Track:
    """config 4/5 track i: kind by i mod 10."""
    seed = SEED_BASE + i
    k = i % 10
    if k <= 4:
        return track_16_stereo(seed, seconds, 44100)
    if k <= 6:
        rng = np.random.default_rng(seed)
        cfg = TrackCfg(16, 1, 4096, 40, 10, 14, 44100)
        n = int(round(seconds * 44100))
        x = make_signal(seed, n, 16, 44100, 1)
        fr = make_frames(rng, cfg, n, False, orders=(0, 31), quants=(1, 15), rice_mods=(1, 7))
        return build_track(cfg, x, fr)
    if k <= 8:
        return track_24_stereo(seed, seconds, 48000)
    return track_24_mono(seed, seconds, 96000)


def make_config(k: int, scale: float = 1.0, n_tracks: int | None = None) -> list[Track]:
    """BASELINE.json configs[k-1].  `scale` shortens durations for tests;
    `n_tracks` bounds the number of UNIQUE tracks generated for configs 4/5
    (callers replicate them physically to reach 1,000 / 10,000)."""
    if k == 1:
        return [track_16_stereo(SEED_BASE + 1, 60.0 * scale)]
    if k == 2:
        return [track_24_stereo(SEED_BASE + 2, 600.0 * scale)]
    if k == 3:
        return [track_16_mono_mixed(SEED_BASE + 3, 60.0 * scale)]
    if k == 4:
        n = n_tracks if n_tracks is not None else 1000
        return [track_16_stereo(SEED_BASE + 1000 + i, 252.0 * scale) for i in range(n)]
    if k == 5:
        n = n_tracks if n_tracks is not None else 10000
        return [corpus_track(i, 252.0 * scale) for i in range(n)]
    raise ValueError(k)


# --------------------------------------------------------------------------
# .m4a muxer
# --------------------------------------------------------------------------
def _atom(kind: bytes, payload: bytes) -> bytes:
    return struct.pack(">I", 8 + len(payload)) + kind + payload


def alac_cookie(cfg: TrackCfg, max_frame_bytes: int = 0, avg_bit_rate: int = 0) -> bytes:
    """24-byte ALACSpecificConfig; offsets match AlacFile.SetInfo (AlacFile.cs:72-92)."""
    return struct.pack(
        ">IBBBBBBHIII", cfg.max_samples_per_frame, 0, cfg.sample_size, cfg.rice_history_mult,
        cfg.rice_initial_history, cfg.rice_kmodifier, cfg.num_channels, 255,
        max_frame_bytes, avg_bit_rate, cfg.sample_rate,
    )


def mux_m4a(track: Track, mdat_first: bool = False, free_atom: bool = False,
            uniform_stsz: bool = False) -> bytes:
    cfg = track.cfg
    nf = track.n_frames
    ftyp = _atom(b"ftyp", b"M4A " + struct.pack(">I", 0) + b"M4A mp42isom")
    mvhd = _atom(b"mvhd", bytes(100))
    tkhd = _atom(b"tkhd", bytes(84))
    mdhd = _atom(b"mdhd", bytes(24))
    hdlr = _atom(b"hdlr", bytes(4) + bytes(4) + b"soun" + bytes(12) + b"SoundHandler\x00")
    smhd = _atom(b"smhd", bytes(8))
    dinf = _atom(b"dinf", _atom(b"dref", bytes(4) + struct.pack(">I", 1) + _atom(b"url ", b"\x00\x00\x00\x01")))
    alac = _atom(b"alac", bytes(4) + alac_cookie(cfg, int(track.stsz.max()) if nf else 0))
    entry = (bytes(6) + struct.pack(">H", 1) + bytes(8) + struct.pack(">HH", cfg.num_channels, cfg.sample_size)
             + bytes(4) + struct.pack(">I", (cfg.sample_rate & 0xFFFF) << 16) + alac)
    stsd = _atom(b"stsd", bytes(4) + struct.pack(">I", 1) + _atom(b"alac", entry))
    # stts: run-length over frame durations (<= 16 entries: DemuxResT.cs:27)
    runs = []
    for d in track.frame_samples.tolist():
        if runs and runs[-1][1] == d:
            runs[-1][0] += 1
        else:
            runs.append([1, d])
    assert len(runs) <= 16, "reference holds at most 16 stts entries"
    stts = _atom(b"stts", bytes(4) + struct.pack(">I", len(runs)) + b"".join(struct.pack(">II", c, d) for c, d in runs))
    stsc = _atom(b"stsc", bytes(4) + struct.pack(">I", 1) + struct.pack(">III", 1, nf, 1))
    if uniform_stsz:
        assert nf and int(track.stsz.min()) == int(track.stsz.max())
        stsz = _atom(b"stsz", bytes(4) + struct.pack(">II", int(track.stsz[0]), nf))
    else:
        stsz = _atom(b"stsz", bytes(4) + struct.pack(">II", 0, nf) + track.stsz.astype(">u4").tobytes())

    def moov_with(stco_off: int) -> bytes:
        stco = _atom(b"stco", bytes(4) + struct.pack(">II", 1, stco_off))
        stbl = _atom(b"stbl", stsd + stts + stsc + stsz + stco)
        minf = _atom(b"minf", smhd + dinf + stbl)
        mdia = _atom(b"mdia", mdhd + hdlr + minf)
        trak = _atom(b"trak", tkhd + mdia)
        return _atom(b"moov", mvhd + trak)

    free = _atom(b"free", bytes(16)) if free_atom else b""
    mdat = _atom(b"mdat", track.mdat)
    if mdat_first:
        off = len(ftyp) + len(free) + 8
        return ftyp + free + mdat + moov_with(off)
    moov_len = len(moov_with(0))
    off = len(ftyp) + moov_len + len(free) + 8
    return ftyp + moov_with(off) + free + mdat


def mux_m4a_ex(track: Track, chunk_frames: int = 7, gap: int = 13, co64: bool = False, mdat_first: bool = False,
               extra_atoms: bool = True, split_stts: bool = False, large_mdat: bool = False) -> bytes:
    """A container the REFERENCE's demuxer rejects but real encoders write: frames grouped into chunks
    of `chunk_frames` with `gap` junk bytes between chunks (stsc/stco addressing), optional co64, mdat
    before moov, 64-bit mdat size, unknown atoms (wide, meta, udta children, sgpd inside stbl) and one
    stts run per frame.  Exercises alacnet::IsoDemux + alacgpu_add_track_offsets."""
    cfg = track.cfg
    nf = track.n_frames
    offs = np.concatenate([[0], np.cumsum(track.stsz.astype(np.int64))])
    # mdat payload: chunks separated by junk
    payload = bytearray()
    chunk_rel = []
    for c0 in range(0, nf, chunk_frames):
        c1 = min(nf, c0 + chunk_frames)
        payload += bytes([0xEE]) * gap
        chunk_rel.append(len(payload))
        payload += track.mdat[offs[c0]:offs[c1]]
    payload += bytes([0xEE]) * gap
    n_chunks = len(chunk_rel)
    last = nf - chunk_frames * (n_chunks - 1)
    ftyp = _atom(b"ftyp", b"M4A " + struct.pack(">I", 512) + b"M4A mp42isom")
    wide = _atom(b"wide", b"") if extra_atoms else b""
    mvhd = _atom(b"mvhd", bytes(100))
    tkhd = _atom(b"tkhd", bytes(84))
    mdhd = _atom(b"mdhd", bytes(24))
    hdlr = _atom(b"hdlr", bytes(4) + bytes(4) + b"soun" + bytes(12) + b"SoundHandler\x00")
    smhd = _atom(b"smhd", bytes(8))
    dinf = _atom(b"dinf", _atom(b"dref", bytes(4) + struct.pack(">I", 1) + _atom(b"url ", b"\x00\x00\x00\x01")))
    alac = _atom(b"alac", bytes(4) + alac_cookie(cfg, int(track.stsz.max()) if nf else 0))
    entry = (bytes(6) + struct.pack(">H", 1) + bytes(8) + struct.pack(">HH", cfg.num_channels, cfg.sample_size)
             + bytes(4) + struct.pack(">I", (cfg.sample_rate & 0xFFFF) << 16) + alac)
    stsd = _atom(b"stsd", bytes(4) + struct.pack(">I", 1) + _atom(b"alac", entry))
    if split_stts:
        runs = [[1, int(d)] for d in track.frame_samples.tolist()]
    else:
        runs = []
        for d in track.frame_samples.tolist():
            if runs and runs[-1][1] == d:
                runs[-1][0] += 1
            else:
                runs.append([1, d])
    stts = _atom(b"stts", bytes(4) + struct.pack(">I", len(runs)) + b"".join(struct.pack(">II", c, d) for c, d in runs))
    stsc_runs = [(1, chunk_frames, 1)]
    if last != chunk_frames:
        stsc_runs.append((n_chunks, last, 1))
    stsc = _atom(b"stsc", bytes(4) + struct.pack(">I", len(stsc_runs)) + b"".join(struct.pack(">III", *r) for r in stsc_runs))
    stsz = _atom(b"stsz", bytes(4) + struct.pack(">II", 0, nf) + track.stsz.astype(">u4").tobytes())
    sgpd = _atom(b"sgpd", bytes(12)) if extra_atoms else b""
    udta = _atom(b"udta", _atom(b"meta", bytes(4) + _atom(b"hdlr", bytes(25)) + _atom(b"ilst", b""))) if extra_atoms else b""
    mdat_hdr = 16 if large_mdat else 8

    def moov_with(base: int) -> bytes:
        if co64:
            stco = _atom(b"co64", bytes(4) + struct.pack(">I", n_chunks) + b"".join(struct.pack(">Q", base + r) for r in chunk_rel))
        else:
            stco = _atom(b"stco", bytes(4) + struct.pack(">I", n_chunks) + b"".join(struct.pack(">I", base + r) for r in chunk_rel))
        stbl = _atom(b"stbl", stsd + stts + sgpd + stsz + stsc + stco)      # not the reference's order either
        minf = _atom(b"minf", smhd + dinf + stbl)
        mdia = _atom(b"mdia", mdhd + hdlr + minf)
        trak = _atom(b"trak", tkhd + _atom(b"edts", _atom(b"elst", bytes(16))) + mdia)
        return _atom(b"moov", mvhd + trak + udta)

    if large_mdat:
        mdat = struct.pack(">I", 1) + b"mdat" + struct.pack(">Q", 16 + len(payload)) + bytes(payload)
    else:
        mdat = _atom(b"mdat", bytes(payload))
    if mdat_first:
        base = len(ftyp) + len(wide) + mdat_hdr
        return ftyp + wide + mdat + moov_with(base)
    moov_len = len(moov_with(0))
    base = len(ftyp) + len(wide) + moov_len + mdat_hdr
    return ftyp + wide + moov_with(base) + mdat


if __name__ == "__main__":  # small smoke: write config 1 to a file
    import sys
    tr = make_config(1, scale=float(sys.argv[2]) if len(sys.argv) > 2 else 0.05)[0]
    data = mux_m4a(tr)
    out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/alacgen_c1.m4a"
    with open(out, "wb") as f:
        f.write(data)
    print(out, len(data), "bytes;", tr.n_frames, "frames; ratio", len(tr.mdat) / max(1, len(tr.pcm)))
