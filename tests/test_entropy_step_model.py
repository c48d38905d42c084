"""CPU: the arithmetic of the GPU entropy step (alac/net_b200/csrc/k1_entropy.cuh, ALACGPU_ENTROPY_STEP)
restated in Python with 32-bit wraparound, plugged into the independent model in place of its
EntropyRiceDecode, and checked against the encoder's input.

What it pins, without a GPU: one symbol per step with the kind of the field (value / zero-run length /
raw field after nine 1-bits) as state; the count of leading 1-bits and floor(log2(history)) taken from the
exponent of an exactly representable float; k carried as k + 127; "consume k - 1 bits unless the field is
>= 2"; the sign modifier across runs; zero runs as a jump of the output index over a cleared buffer; the
step never moving the cursor by more than 32 bits; and the stop conditions (negative history, run past
the frame buffer) that the kernel reports as frame status 6 / 7.  The CUDA code itself is covered by the
GPU parity tests; this keeps its algorithm honest on the CPU tier."""
import struct

import numpy as np
import pytest

from pymodel import alac_model as M

U32 = 0xFFFFFFFF
MAX_FRAME_SAMPLES = 16384


def _exp_of(bits_4b: int) -> int:
    """exponent field of (float(bits) - 2^23): 127 + floor(log2(v)) for 0 < v < 2^23, 0 for v == 0"""
    f = struct.unpack("<f", struct.pack("<I", bits_4b & U32))[0]
    g = np.float32(f) - np.float32(8388608.0)
    return (struct.unpack("<I", struct.pack("<f", float(g)))[0] >> 23) & 0x1FF


def _s32(v: int) -> int:
    v &= U32
    return v - (1 << 32) if v & 0x80000000 else v


class _Window:
    """32-bit big-endian window at a bit position (zero fill past the data), like BitCursor::peek."""

    def __init__(self, br):
        self.value, self.nbits, self.pos = br.value, br.nbits, br.pos

    def peek(self) -> int:
        end = self.pos + 32
        if end <= self.nbits:
            return (self.value >> (self.nbits - end)) & U32
        have = max(0, self.nbits - self.pos)
        return ((self.value & ((1 << have) - 1)) << (32 - have)) & U32 if have else 0


def step_rice_decode(br, n, rss, initial_history, kmod, mult, mask, one_step_escape=False, lane_mode=False, consume=None):
    """Drop-in for alac_model.rice_decode built from the kernel's step.  `mask` is the run-length
    multiplier mask (1 << kmod) - 1 the model passes for the zero-run symbol (AlacFile.cs:236).

    lane_mode: the variant of the frame-lane kernels (kf_frame.cu, ALACGPU_ENTROPY_STEP_F): the residual stays in
    a register (`have`, `e`) until the predictor of the same lane consumes it -- which happens only in rounds
    where every lane of the warp has one, modelled by `consume(round) -> bool` -- a lane holding a residual
    (or still owing zeros of a run) sits the step out, and a zero run becomes `pend` zero residuals handed out one per round (clipped to the
    frame) while the symbol index jumps as before."""
    assert mask == (1 << kmod) - 1
    win = _Window(br)
    out = [0] * n                                   # the cleared plane row
    kcap = kmod + 127
    kk_after_run = min(128, kcap)
    rssh = 32 - rss
    i, nc, h, smm1 = 0, n, initial_history, U32
    kk = min(_exp_of(0x4B000000 | ((initial_history >> 9) + 3)), kcap)
    mk = ((1 << (kk - 127)) - 1) & U32
    mm = mk
    run, raw = False, False
    steps = 0
    have, pend, e_reg, j, rounds = False, 0, 0, 0, 0
    while i < nc or (lane_mode and (have or pend)):
        if lane_mode:
            rounds += 1
            assert rounds <= 16 * n + 64
            if have or pend or i >= nc:             # the lane sits this step out: nothing of its state moves
                if not have and pend:
                    have, e_reg, pend = True, 0, pend - 1
                if have and (consume is None or consume(rounds)):
                    out[j] = e_reg
                    j += 1
                    have = False
                continue
        steps += 1
        assert steps <= 4 * n, "at most four steps per sample: value and run length, each with its raw field"
        w = win.peek()
        ex = _exp_of(((w >> 23) ^ 0x1FF) | 0x4B000000)
        esc = ex == 0
        x = (135 - ex) & U32
        s0 = (kk - ex + 8) & U32
        s1 = (s0 + 1) & U32
        sh = s1 & 31                                # shf.l.wrap
        e = ((w >> (32 - sh)) if sh else 0) & mk
        em = max(e, 1)
        rice = (x * mm + smm1 + em) & U32
        rsh = 16 if run else rssh
        rawv = ((w >> rsh) + smm1 + 1) & U32
        dv = rawv if raw else rice
        cons = s1 if e >= 2 else s0
        if esc or raw:
            cons = (32 - rsh) if raw else 9
        # not in the kernel yet: when nine 1-bits and the raw field fit the 32-bit window
        # together (16-bit material, and every run-length escape), the escape needs no second step
        fused_esc = one_step_escape and esc and not raw and 9 + (32 - rsh) <= 32
        if fused_esc:
            dv = ((((w << 9) & U32) >> rsh) + smm1 + 1) & U32
            cons = 9 + (32 - rsh)
        cons &= U32
        assert cons <= 32, "a step moves the cursor by at most one word"
        win.pos += cons
        pend_raw = esc and not raw and not fused_esc
        is_val = not pend_raw and not run
        is_run = not pend_raw and run
        hb = _s32(h - (_s32(h * mult) >> 9))
        hn = _s32(dv * mult + hb)
        if dv > 0xFFFF:
            hn = 0xFFFF
        isum = (i + dv) & U32
        if is_val:
            if lane_mode:
                have, e_reg = True, _s32((dv >> 1) ^ (-(dv & 1) & U32))
            else:
                out[i] = _s32((dv >> 1) ^ (-(dv & 1) & U32))
            i += 1
        if is_run:
            pend = min(isum, nc) - i                # zeros of the run that lie inside the frame
            i = isum
        to_run = is_val and (hn & U32) < 128 and i < nc
        if is_val and hn < 0:
            nc = 0                                  # the kernel stops the lane: frame status 6
            raise M.Unmodelled("negative rice history")
        if is_run and i > MAX_FRAME_SAMPLES:
            raise M.Unmodelled("zero run beyond the scratch buffer")      # frame status 7
        if is_val:
            h = 0 if to_run else hn
            smm1 = U32
        if is_run:
            smm1 = U32 if dv > 0xFFFF else 0
        kkv = min(_exp_of(((hn >> 9) + 0x4B000003) & U32), kcap)
        hz = hn & U32
        flo = hz.bit_length() - 1 if hz else -1
        kkz = 143 if hz == 0 else (((hz + 16) >> 6) - flo + 134) & U32
        kk = kkz if to_run else (kk_after_run if is_run else kkv)
        mk = ((1 << ((kk + 1) & 31)) - 1) & U32        # shf.l.wrap(2, 2, kk) - 1
        mm = mk & (mask if to_run else U32)
        run = run if pend_raw else to_run
        raw = pend_raw
        if lane_mode:
            if not have and pend:
                have, e_reg, pend = True, 0, pend - 1
            if have and (consume is None or consume(rounds)):
                out[j] = e_reg
                j += 1
                have = False
    if lane_mode:
        assert j == n and not have and not pend, "every residual was handed out exactly once"
    br.pos = win.pos
    return out


CASES = [
    (16, 2, dict(orders=(0, 31), quants=(0, 15), rice_mods=(0, 7))),
    (24, 2, dict(orders=(0, 31), quants=(1, 15), rice_mods=(1, 7))),
    (24, 1, dict(orders=(0, 31), quants=(1, 15), rice_mods=(1, 7))),
    (16, 2, dict(orders=(4, 8), quants=(9, 9), rice_mods=(4, 4), loud=True, auto_escape=False)),
    (24, 2, dict(orders=(4, 8), quants=(9, 9), rice_mods=(7, 7), loud=True, auto_escape=False)),
    (16, 2, dict(orders=(4, 8), quants=(9, 9), rice_mods=(4, 4), kmod=6, hist_mult=63, init_hist=200)),
    (16, 2, dict(orders=(4, 8), quants=(9, 9), rice_mods=(4, 4), kmod=0, auto_escape=False)),
    (16, 2, dict(orders=(4, 8), quants=(9, 9), rice_mods=(4, 4), kmod=1, auto_escape=False)),
    (16, 2, dict(orders=(4, 8), quants=(9, 9), rice_mods=(4, 4), kmod=31, hist_mult=255, init_hist=255)),
    (16, 2, dict(orders=(4, 8), quants=(9, 9), rice_mods=(0, 7), hist_mult=3, auto_escape=False)),
    (16, 1, dict(orders=(1, 4), quants=(9, 9), rice_mods=(4, 4), quiet=True)),
]


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_step_state_machine_inverts_the_encoder(idx, gen, monkeypatch):
    ss, ch, kw = CASES[idx]
    kw = dict(kw)
    rng = np.random.default_rng(4000 + idx)
    cfg = gen.TrackCfg(ss, ch, 256, kw.pop("hist_mult", 40), kw.pop("init_hist", 10), kw.pop("kmod", 14), 44100)
    total = 256 * 5 + 77
    x = gen.make_signal(int(rng.integers(1, 1 << 31)), total, ss, 44100, ch, wasted_spans=(ss == 24)).copy()
    if kw.pop("loud", False):
        lim = 1 << (ss - 1)
        x[:, ::3] = rng.integers(-lim, lim, size=x[:, ::3].shape)
    if kw.pop("quiet", False):          # long zero runs, short zero runs and isolated small values
        x[:] = 0
        x[:, rng.integers(0, total, size=60)] = rng.integers(-3, 4, size=(ch, 60))
    fr = gen.make_frames(rng, cfg, total, ch == 2, **kw)
    if ss == 24:
        gen.assign_wasted_bytes(fr, x, 24, rng)
    t = gen.build_track(cfg, x, fr)
    ck = M.Cookie(cfg.sample_size, cfg.num_channels, cfg.max_samples_per_frame, cfg.rice_history_mult,
                  cfg.rice_initial_history, cfg.rice_kmodifier)
    want = M.decode_track(ck, t.mdat, t.stsz)
    assert want == t.pcm
    calls = []

    def counted(*a):
        calls.append(1)
        return step_rice_decode(*a)

    monkeypatch.setattr(M, "rice_decode", counted)
    assert M.decode_track(ck, t.mdat, t.stsz) == t.pcm, "the step state machine does not reproduce the model"
    # the encoder may store a frame uncompressed unless told not to; every compressed channel went through the step
    assert len(calls) > 0 and (CASES[idx][2].get("auto_escape", True) or len(calls) == ch * t.n_frames)


@pytest.mark.parametrize("idx", [0, 3, 5, 10])
def test_one_step_escape_variant(idx, gen, monkeypatch):
    """groundwork: the escape folded into one step where the window allows it decodes the same streams"""
    ss, ch, kw = CASES[idx]
    kw = dict(kw)
    rng = np.random.default_rng(5000 + idx)
    cfg = gen.TrackCfg(ss, ch, 256, kw.pop("hist_mult", 40), kw.pop("init_hist", 10), kw.pop("kmod", 14), 44100)
    total = 256 * 4 + 9
    x = gen.make_signal(int(rng.integers(1, 1 << 31)), total, ss, 44100, ch).copy()
    if kw.pop("loud", False):
        lim = 1 << (ss - 1)
        x[:, ::3] = rng.integers(-lim, lim, size=x[:, ::3].shape)
    if kw.pop("quiet", False):
        x[:] = 0
        x[:, rng.integers(0, total, size=40)] = rng.integers(-3, 4, size=(ch, 40))
    t = gen.build_track(cfg, x, gen.make_frames(rng, cfg, total, ch == 2, **kw))
    ck = M.Cookie(cfg.sample_size, cfg.num_channels, cfg.max_samples_per_frame, cfg.rice_history_mult,
                  cfg.rice_initial_history, cfg.rice_kmodifier)
    monkeypatch.setattr(M, "rice_decode", lambda *a: step_rice_decode(*a, one_step_escape=True))
    assert M.decode_track(ck, t.mdat, t.stsz) == t.pcm


def test_float_exponent_is_floor_log2_on_the_whole_domain():
    """every leading-ones pattern of the 9-bit prefix, and every history value the k formula can see"""
    for v9 in range(512):
        ex = _exp_of(v9 | 0x4B000000)
        assert ex == (0 if v9 == 0 else 127 + v9.bit_length() - 1)
    for v in list(range(3, 70000)) + [(1 << 23) - 1, 1 << 22, (1 << 22) + 1]:
        assert _exp_of(0x4B000000 | v) == 127 + v.bit_length() - 1


@pytest.mark.parametrize("idx", [0, 1, 3, 5, 9, 10])
def test_frame_lane_variant_hands_out_every_residual_once(idx, gen, monkeypatch):
    """kf_frame.cu: residual in a register until the lane's predictor takes it, zero runs as pending zeros,
    a waiting lane's state frozen -- with the consumer (the warp-wide predictor round) arriving at irregular
    intervals"""
    ss, ch, kw = CASES[idx]
    kw = dict(kw)
    rng = np.random.default_rng(6000 + idx)
    cfg = gen.TrackCfg(ss, ch, 256, kw.pop("hist_mult", 40), kw.pop("init_hist", 10), kw.pop("kmod", 14), 44100)
    total = 256 * 4 + 31
    x = gen.make_signal(int(rng.integers(1, 1 << 31)), total, ss, 44100, ch, wasted_spans=(ss == 24)).copy()
    if kw.pop("loud", False):
        lim = 1 << (ss - 1)
        x[:, ::3] = rng.integers(-lim, lim, size=x[:, ::3].shape)
    if kw.pop("quiet", False):
        x[:] = 0
        x[:, rng.integers(0, total, size=40)] = rng.integers(-3, 4, size=(ch, 40))
    fr = gen.make_frames(rng, cfg, total, ch == 2, **kw)
    if ss == 24:
        gen.assign_wasted_bytes(fr, x, 24, rng)
    t = gen.build_track(cfg, x, fr)
    ck = M.Cookie(cfg.sample_size, cfg.num_channels, cfg.max_samples_per_frame, cfg.rice_history_mult,
                  cfg.rice_initial_history, cfg.rice_kmodifier)
    pattern = rng.integers(0, 3, size=997)
    for consume in (None, lambda r: pattern[r % 997] == 0):
        monkeypatch.setattr(M, "rice_decode", lambda *a: step_rice_decode(*a, lane_mode=True, consume=consume))
        assert M.decode_track(ck, t.mdat, t.stsz) == t.pcm
