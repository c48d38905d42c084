"""GPU parity: libalacgpu (through the C ABI) vs the oracle, byte for byte."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _decode(tracks, **kw):
    from alac.net_b200 import BatchDecoder
    with BatchDecoder(**kw) as dec:
        for t in tracks:
            dec.add_track(t.cfg, t.mdat, t.stsz)
        pcm, off, ln, status = dec.decode_all()
        timing = dec.timing()
        return [pcm[int(o):int(o + l)].tobytes() for o, l in zip(off, ln)], status, timing


def _oracle(oracle, t):
    return oracle.decode_track(oracle.cfg_from(t.cfg), t.mdat, t.stsz)


@pytest.mark.parametrize("k,scale", [(1, 0.1), (2, 0.02), (3, 0.2)])
def test_configs_small(k, scale, gen, oracle):
    tracks = gen.make_config(k, scale=scale)
    got, status, timing = _decode(tracks)
    pos = 0
    for t, g in zip(tracks, got):
        ref, st, fb = _oracle(oracle, t)
        assert ref == t.pcm, "oracle does not reproduce the encoder's input"
        assert np.array_equal(status[pos:pos + t.n_frames], st)
        assert len(g) == len(ref)
        if g != ref:
            a, b = np.frombuffer(g, np.uint8), np.frombuffer(ref, np.uint8)
            bad = np.nonzero(a != b)[0]
            raise AssertionError(f"config {k}: {bad.size} bytes differ, first at {bad[:8]}")
        pos += t.n_frames
    assert timing["kernel_launches"] > 0
