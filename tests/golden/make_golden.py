#!/usr/bin/env python
"""Regenerates tests/golden/alac_golden.npz.

The reference (C#) cannot run in this image and ships no vectors, so the golden
set is produced by the two independent restatements: every vector is decoded by
the pure-Python model (pymodel/alac_model.py, written from the reference source)
AND by the C oracle (oracle/alac_oracle.c); the script refuses to write a vector
on which they disagree, or which does not reproduce the encoder's input PCM.

    python tests/golden/make_golden.py

Stored per vector: cookie fields, the frame bytes (mdat), stsz, expected PCM.
Deterministic (fixed seeds); ~150 KB.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O            # noqa: E402
from pymodel import alac_model as M       # noqa: E402
from tools.alacgen import alacgen as G    # noqa: E402

# name, sample_size, container channels, stereo element, nmax, frames, frame policy
VECTORS = [
    ("s16_stereo_all_orders", 16, 2, True, 512, 5, dict(orders=(0, 31), quants=(1, 15), rice_mods=(1, 7))),
    ("s16_stereo_quant0_ricemod0", 16, 2, True, 256, 4, dict(orders=(1, 30), quants=(0, 0), rice_mods=(0, 0))),
    ("s16_mono", 16, 1, False, 512, 4, dict(orders=(0, 31), quants=(1, 15), rice_mods=(1, 7))),
    ("s16_mono_in_stereo_container", 16, 2, False, 256, 3, dict(orders=(2, 8), quants=(4, 10), rice_mods=(4, 4))),
    ("s16_stereo_in_mono_container", 16, 1, True, 256, 3, dict(orders=(2, 8), quants=(4, 10), rice_mods=(4, 4))),
    ("s24_stereo_wasted", 24, 2, True, 512, 6, dict(orders=(1, 31), quants=(1, 15), rice_mods=(1, 7))),
    ("s24_mono_wasted", 24, 1, False, 512, 4, dict(orders=(1, 31), quants=(1, 15), rice_mods=(1, 7))),
    ("s16_escape_frames", 16, 2, True, 256, 4, dict(orders=(1, 8), quants=(4, 10), rice_mods=(4, 4), escape_prob=0.6)),
    ("s24_escape_frames", 24, 2, True, 256, 4, dict(orders=(1, 8), quants=(4, 10), rice_mods=(4, 4), escape_prob=0.6)),
    ("s16_loud_rice_escapes", 16, 2, True, 512, 3, dict(orders=(4, 8), quants=(9, 9), rice_mods=(4, 4), loud=True, auto_escape=False)),
    ("s24_loud_rice_escapes", 24, 2, True, 512, 3, dict(orders=(4, 8), quants=(9, 9), rice_mods=(7, 7), loud=True, auto_escape=False)),
    ("s16_kmod6_hist63", 16, 2, True, 256, 3, dict(orders=(4, 8), quants=(9, 9), rice_mods=(4, 4), kmod=6, hist_mult=63, init_hist=200)),
    ("s16_full_frame_4096", 16, 2, True, 4096, 2, dict(orders=(8, 30), quants=(6, 12), rice_mods=(3, 5))),
]


def build(name, ss, cch, stereo, nmax, nf, kw, seed):
    kw = dict(kw)
    rng = np.random.default_rng(seed)
    cfg = G.TrackCfg(ss, cch, nmax, kw.pop("hist_mult", 40), kw.pop("init_hist", 10), kw.pop("kmod", 14), 44100)
    total = nmax * (nf - 1) + int(rng.integers(1, nmax + 1))
    x = G.make_signal(seed * 7 + 1, total, ss, 44100, 2 if stereo else 1, wasted_spans=(ss == 24)).copy()
    if kw.pop("loud", False):
        lim = 1 << (ss - 1)
        x[:, ::3] = rng.integers(-lim, lim, size=x[:, ::3].shape)
    fr = G.make_frames(rng, cfg, total, stereo, **kw)
    if ss == 24:
        G.assign_wasted_bytes(fr, x, 24, rng)
    return G.build_track(cfg, x, fr)


def main():
    out = {}
    names = []
    for i, (name, ss, cch, stereo, nmax, nf, kw) in enumerate(VECTORS):
        t = build(name, ss, cch, stereo, nmax, nf, kw, 0x601D + i)
        ck = M.Cookie(t.cfg.sample_size, t.cfg.num_channels, t.cfg.max_samples_per_frame,
                      t.cfg.rice_history_mult, t.cfg.rice_initial_history, t.cfg.rice_kmodifier)
        model = M.decode_track(ck, t.mdat, t.stsz)
        ref, status, _ = O.decode_track(O.cfg_from(t.cfg), t.mdat, t.stsz)
        if not (status == 0).all() or model != ref or ref != t.pcm:
            raise SystemExit(f"{name}: model / oracle / encoder input disagree -- not writing golden data")
        names.append(name)
        out[f"{name}.cfg"] = np.array([ck.sample_size, ck.num_channels, ck.max_samples_per_frame,
                                       ck.rice_history_mult, ck.rice_initial_history, ck.rice_kmodifier], dtype=np.int32)
        out[f"{name}.mdat"] = np.frombuffer(t.mdat, dtype=np.uint8)
        out[f"{name}.stsz"] = t.stsz.astype(np.uint32)
        out[f"{name}.pcm"] = np.frombuffer(ref, dtype=np.uint8)
        print(f"{name}: {t.n_frames} frames, {len(t.mdat)} B in, {len(ref)} B PCM")
    out["names"] = np.array(names)
    path = os.path.join(HERE, "alac_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
