"""CPU: the C oracle against the independent Python model (pymodel/) and the
encoder's input, on randomised small frames.  Pins the oracle twice: two
restatements written separately from the reference source must agree byte for
byte, and both must invert a from-scratch encoder."""
import numpy as np
import pytest

from pymodel import alac_model as M


def _cookie(cfg):
    return M.Cookie(cfg.sample_size, cfg.num_channels, cfg.max_samples_per_frame,
                    cfg.rice_history_mult, cfg.rice_initial_history, cfg.rice_kmodifier)


def _small_track(gen, rng, sample_size, container_ch, stereo_element, nmax, n_frames, **kw):
    cfg = gen.TrackCfg(sample_size, container_ch, nmax, kw.pop("hist_mult", 40), kw.pop("init_hist", 10),
                       kw.pop("kmod", 14), 44100)
    total = nmax * (n_frames - 1) + int(rng.integers(1, nmax + 1))
    chans = 2 if stereo_element else 1
    x = gen.make_signal(int(rng.integers(1, 1 << 31)), total, sample_size, 44100, chans,
                        wasted_spans=(sample_size == 24)).copy()
    if kw.pop("loud", False):       # full-scale noise: escapes and history clamps
        lim = 1 << (sample_size - 1)
        x[:, ::3] = rng.integers(-lim, lim, size=x[:, ::3].shape)
    fr = gen.make_frames(rng, cfg, total, stereo_element, **kw)
    if sample_size == 24:
        gen.assign_wasted_bytes(fr, x, 24, rng)
    return gen.build_track(cfg, x, fr)


CASES = [
    # sample_size, container channels, stereo element, extra frame policy
    (16, 2, True, dict(orders=(0, 31), quants=(0, 15), rice_mods=(0, 7))),
    (16, 1, False, dict(orders=(0, 31), quants=(1, 15), rice_mods=(0, 7))),
    (16, 2, False, dict(orders=(1, 8), quants=(4, 12), rice_mods=(4, 4))),       # mono element, stereo container
    (16, 1, True, dict(orders=(1, 8), quants=(4, 12), rice_mods=(4, 4))),        # stereo element, mono container
    (24, 2, True, dict(orders=(0, 31), quants=(1, 15), rice_mods=(1, 7))),
    (24, 1, False, dict(orders=(0, 31), quants=(1, 15), rice_mods=(1, 7))),
    (24, 2, False, dict(orders=(2, 30), quants=(1, 15), rice_mods=(1, 7))),
    (16, 2, True, dict(orders=(1, 31), quants=(1, 15), rice_mods=(1, 7), escape_prob=0.5)),
    (24, 2, True, dict(orders=(1, 31), quants=(1, 15), rice_mods=(1, 7), escape_prob=0.5)),
    (16, 2, True, dict(orders=(4, 8), quants=(9, 9), rice_mods=(4, 4), loud=True, auto_escape=False)),
    (24, 2, True, dict(orders=(4, 8), quants=(9, 9), rice_mods=(7, 7), loud=True, auto_escape=False)),
    (16, 2, True, dict(orders=(4, 8), quants=(9, 9), rice_mods=(4, 4), kmod=6, hist_mult=63, init_hist=200)),
    (16, 1, False, dict(orders=(30, 31), quants=(0, 2), rice_mods=(0, 1), end_tag=False)),
]


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_oracle_equals_model_equals_encoder_input(idx, gen, oracle):
    ss, cch, stereo, kw = CASES[idx]
    rng = np.random.default_rng(1000 + idx)
    t = _small_track(gen, rng, ss, cch, stereo, nmax=int(rng.integers(40, 200)), n_frames=6, **dict(kw))
    ref, status, fbytes = oracle.decode_track(oracle.cfg_from(t.cfg), t.mdat, t.stsz)
    assert (status == 0).all(), status
    model = M.decode_track(_cookie(t.cfg), t.mdat, t.stsz)
    assert model == ref, "C oracle and Python model disagree"
    # quant 0 (1 << 31 rounding term) and friends are quirks the encoder mirrors, so the round trip holds too
    assert ref == t.pcm, "oracle does not invert the encoder"


def test_full_size_frame_stereo24(gen, oracle):
    """one 4096-sample 24-bit stereo frame with high orders through both restatements"""
    rng = np.random.default_rng(77)
    t = _small_track(gen, rng, 24, 2, True, nmax=4096, n_frames=1, orders=(24, 30), quants=(1, 15), rice_mods=(1, 7))
    ref, status, _ = oracle.decode_track(oracle.cfg_from(t.cfg), t.mdat, t.stsz)
    assert status[0] == 0
    assert M.decode_track(_cookie(t.cfg), t.mdat, t.stsz) == ref == t.pcm


def test_stage_intermediates_match_model(gen, oracle):
    """residuals and predictor outputs of the oracle's stage dump vs the model's functions"""
    rng = np.random.default_rng(5)
    t = _small_track(gen, rng, 16, 2, True, nmax=300, n_frames=2, orders=(3, 12), quants=(3, 9), rice_mods=(2, 6))
    ck = _cookie(t.cfg)
    pos = 0
    for sz in t.stsz:
        frame = t.mdat[pos:pos + int(sz)]
        pos += int(sz)
        _, _, st, stages = oracle.decode_frame_stages(oracle.cfg_from(t.cfg), frame)
        assert st == 0
        n = stages.n
        # replay the model on the oracle's residuals: predictor outputs must match
        for c in range(2):
            e = list(stages.residual[c][:n])
            coef = list(stages.coef[c][:stages.order[c]])
            o = M.predict(e, n, 17, coef, stages.order[c], stages.quant[c])
            assert o == list(stages.predicted[c][:n])


# ---- primitive / quirk tests on hand-built bits -------------
class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, v, n):
        for i in range(n - 1, -1, -1):
            self.bits.append((v >> i) & 1)
        return self

    def bytes(self):
        b = self.bits + [0] * ((-len(self.bits)) % 8)
        return bytes(int("".join(map(str, b[i:i + 8])), 2) for i in range(0, len(b), 8))


def test_clz_quirk(oracle):
    L = oracle.lib()
    assert L.alac_oracle_clz(0) == 40 == M.clz_quirk(0)            # AlacFile.cs:190 falls off the table
    for v in (1, 2, 3, 127, 128, 255, 256, 0x7FFF, 0x8000, 0xFFFF, 0x10000, 0x7FFFFFFF, -1, -12345):
        assert L.alac_oracle_clz(v) == M.clz_quirk(v), v
    assert M.clz_quirk(1) == 31 and M.clz_quirk(0x7FFFFFFF) == 1 and M.clz_quirk(-1) == 0


def _mono16_header(w, n, nmax, order=0, quant=0, rice_mod=4, coefs=(), ub=0, escape=0):
    w.put(0, 3).put(0, 4).put(0, 12)
    w.put(1 if n != nmax else 0, 1).put(ub, 2).put(escape, 1)
    if n != nmax:
        w.put(n, 32)
    if not escape:
        w.put(0, 8).put(0, 8).put(0, 4).put(quant, 4).put(rice_mod, 3).put(order, 5)
        for c in coefs:
            w.put(c & 0xFFFF, 16)
    return w


def _both(oracle, cfg_kw, frame):
    cfg = oracle.make_cfg(**cfg_kw)
    pcm, st = oracle.read_frame(cfg, frame)
    ck = M.Cookie(cfg_kw.get("sample_size", 16), cfg_kw.get("num_channels", 2),
                  cfg_kw.get("max_samples_per_frame", 4096), cfg_kw.get("rice_history_mult", 40),
                  cfg_kw.get("rice_initial_history", 10), cfg_kw.get("rice_kmodifier", 14))
    return pcm, st, M.read_frame(ck, frame)


def test_rice_symbols_by_hand(oracle):
    """k=1 unary values, remainder 0 (k-1 bits) vs non-0 (k bits), nine-ones escape; order 0 so
    the residuals are the samples.  history starts at 10 -> k=1 (A.2)."""
    kw = dict(sample_size=16, num_channels=1, max_samples_per_frame=4, rice_initial_history=10)
    w = _mono16_header(BitWriter(), 4, 4)
    # history 10, mult 40: k = min(31 - clz((h>>9)+3), 14) = 1 while h < 512
    w.put(0b0, 1)            # dv 0 -> 0 ; history 10 - 0 = 10 (<128) -> zero-run symbol follows
    # run symbol: k = clz(10)+ (26/64=0) - 24 = 28 - 24 = 4 ; M = 15 ; run 0 = "0" + remainder 0 -> 3 zero bits
    w.put(0b0, 1).put(0, 3)
    # next value carries signModifier 1: coded dv-1. want sample -1 -> dv 1 -> code 0
    w.put(0b0, 1)
    # history = 0 + 1*40 = 40 -> run symbol with k = clz(40) + 56/64 - 24 = 2 (M = 3):
    # run 1 -> q=0, r=1 -> "0" + (r+1 = 2 in 2 bits)
    w.put(0b0, 1).put(2, 2)
    # last sample (index 3): coded value 2 (+1 signmod) = dv 3 -> -2 ; k=1 so unary 110
    w.put(0b110, 3)
    frame = w.bytes() + b"\0\0"
    pcm, st, model = _both(oracle, kw, frame)
    assert st == 0
    assert pcm == model == np.array([0, -1, 0, -2], dtype="<i2").tobytes()


def test_escape_symbol_and_raw_width(oracle):
    kw = dict(sample_size=16, num_channels=1, max_samples_per_frame=2)
    w = _mono16_header(BitWriter(), 2, 2)
    w.put(0x1FF, 9).put(0x8001 & 0xFFFF, 16)      # nine ones then 16 raw bits: dv = 0x8001 -> -(0x8002/2)
    w.put(0x1FF, 9).put(0x0004, 16)               # history was clamped? dv<=0xFFFF so no; second escape: dv 4 -> 2
    frame = w.bytes() + b"\0\0\0"
    pcm, st, model = _both(oracle, kw, frame)
    assert st == 0 and pcm == model
    v = np.frombuffer(pcm, "<i2")
    assert v[0] == np.int16(-(0x8002 // 2)) and v[1] == 2


def test_uncompressed_frames_16_and_24(oracle):
    w = _mono16_header(BitWriter(), 3, 3, escape=1)
    for s in (-32768, 32767, -2):
        w.put(s & 0xFFFF, 16)
    pcm, st, model = _both(oracle, dict(sample_size=16, num_channels=1, max_samples_per_frame=3), w.bytes())
    assert st == 0 and pcm == model == np.array([-32768, 32767, -2], "<i2").tobytes()
    # 24-bit stereo escape frame: 16 + 8 bit reads, (x ^ m) - m sign extension (AlacFile.cs:676-691)
    w = BitWriter().put(1, 3).put(0, 16).put(0, 1).put(0, 2).put(1, 1)
    vals = [(-8388608, 8388607), (-1, 1)]
    for l, r in vals:
        w.put(l & 0xFFFFFF, 24).put(r & 0xFFFFFF, 24)
    pcm, st, model = _both(oracle, dict(sample_size=24, num_channels=2, max_samples_per_frame=2), w.bytes())
    exp = b"".join(int(v & 0xFFFFFF).to_bytes(3, "little") for pair in vals for v in pair)
    assert st == 0 and pcm == model == exp


def test_mono_element_in_stereo_container_zero_fills(oracle):
    w = _mono16_header(BitWriter(), 2, 2, escape=1)
    w.put(5, 16).put(0xFFFF, 16)
    pcm, st, model = _both(oracle, dict(sample_size=16, num_channels=2, max_samples_per_frame=2), w.bytes())
    assert st == 0 and pcm == model == np.array([5, 0, -1, 0], "<i2").tobytes()


def test_sixteen_bit_ignores_wasted_bytes_on_output(oracle, gen):
    """quirk 3: ub != 0 on a 16-bit stream still shrinks rss and is parsed, but the shift planes are
    never merged (Deinterlace16 has no such path).  Build by hand: escape=0, ub=1, order 0."""
    n = 2
    w = _mono16_header(BitWriter(), n, n, ub=1)
    w.put(0xAA, 8).put(0xBB, 8)          # wasted-byte plane, ignored on output
    # two residuals with rss = 8: dv 6 -> 3 (unary 1111110), then run symbol etc. keep it simple: k=1
    w.put(0b1111110, 7)                  # dv 6 -> +3, history = 10 + 6*40 - (10*40>>9 = 0) = 250 (>=128)
    w.put(0b110, 3)                      # dv 2 -> +1
    pcm, st, model = _both(oracle, dict(sample_size=16, num_channels=1, max_samples_per_frame=n), w.bytes() + b"\0\0")
    assert st == 0 and pcm == model == np.array([3, 1], "<i2").tobytes()


def test_first_sample_is_not_sign_extended(oracle):
    """quirk 6: o[0] = e[0] verbatim (AlacFile.cs:259-260); visible through 24-bit packing when the
    escape symbol delivers a raw value wider than the sample."""
    n = 2
    w = BitWriter().put(0, 3).put(0, 16).put(0, 1).put(0, 2).put(0, 1)      # mono, full size, compressed
    w.put(0, 16).put(0, 4).put(4, 4).put(4, 3).put(31, 5)                    # order 31 (delta mode)
    for _ in range(31):
        w.put(0, 16)
    w.put(0x1FF, 9).put(0xFFFFFE, 24)     # dv = 0xFFFFFE -> sample 0x7FFFFF (in range), then
    w.put(0b0, 1)                         # history = 0xFFFF (dv > 0xFFFF) -> k = min(31-clz(130), 14) = 7: "0" + 6 zero bits = dv 0
    w.put(0, 6)
    pcm, st, model = _both(oracle, dict(sample_size=24, num_channels=1, max_samples_per_frame=n), w.bytes() + b"\0\0\0")
    assert st == 0 and pcm == model
    assert pcm[:3] == (0x7FFFFF).to_bytes(3, "little")


def test_truncated_and_malformed_frames_have_policy_status(oracle):
    cfg = oracle.make_cfg(sample_size=16, num_channels=2, max_samples_per_frame=16)
    pcm, st = oracle.read_frame(cfg, bytes([0b01000000, 0, 0, 0]))        # tag 2
    assert st == 1 and pcm == bytes(16 * 4)
    pcm, st = oracle.read_frame(cfg, b"")                                  # empty frame: tag 0, runs out of bits
    assert st == 4 and pcm == bytes(16 * 4)
    w = BitWriter().put(1, 3).put(0, 16).put(1, 1).put(0, 2).put(0, 1).put(20000, 32)   # N beyond 16384
    pcm, st = oracle.read_frame(cfg, w.bytes())
    assert st == 3 and pcm == b""
