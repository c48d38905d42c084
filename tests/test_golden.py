"""Committed golden vectors (tests/golden/alac_golden.npz, made by make_golden.py):
the C oracle and the Python model on CPU, libalacgpu on the GPU."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "alac_golden.npz")


def vectors():
    z = np.load(GOLDEN)
    for name in z["names"].tolist():
        yield name, z[f"{name}.cfg"].tolist(), z[f"{name}.mdat"].tobytes(), z[f"{name}.stsz"], z[f"{name}.pcm"].tobytes()


NAMES = [v[0] for v in vectors()]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name, oracle):
    for n, cfg, mdat, stsz, pcm in vectors():
        if n != name:
            continue
        got, status, _ = oracle.decode_track(oracle.make_cfg(*cfg), mdat, stsz)
        assert (status == 0).all() and got == pcm


def test_python_model_reproduces_golden_small_vectors():
    from pymodel import alac_model as M
    for n, cfg, mdat, stsz, pcm in vectors():
        if len(pcm) > 12000:          # pure-Python loops: keep the CPU suite fast
            continue
        assert M.decode_track(M.Cookie(*cfg), mdat, stsz) == pcm, n


class _Cfg:
    def __init__(self, v):
        (self.sample_size, self.num_channels, self.max_samples_per_frame, self.rice_history_mult,
         self.rice_initial_history, self.rice_kmodifier) = v
        self.sample_rate = 44100


@pytest.mark.gpu
def test_gpu_reproduces_golden_one_batch():
    """all golden vectors as ONE batch (mixed 16/24-bit, mono/stereo tracks in one decode_all)"""
    from alac.net_b200 import BatchDecoder
    vs = list(vectors())
    with BatchDecoder(devices=[0]) as dec:
        for n, cfg, mdat, stsz, pcm in vs:
            dec.add_track(_Cfg(cfg), mdat, stsz)
        out, off, ln, status = dec.decode_all()
        assert (status == 0).all()
        for (n, cfg, mdat, stsz, pcm), o, l in zip(vs, off, ln):
            assert out[int(o):int(o + l)].tobytes() == pcm, n


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_reproduces_golden_single(name):
    from alac.net_b200 import BatchDecoder
    for n, cfg, mdat, stsz, pcm in vectors():
        if n != name:
            continue
        with BatchDecoder(devices=[0]) as dec:
            dec.add_track(_Cfg(cfg), mdat, stsz)
            out, off, ln, status = dec.decode_all()
            assert (status == 0).all() and out[:int(ln[0])].tobytes() == pcm
