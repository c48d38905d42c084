"""Helper run in a SUBPROCESS by test_gpu_parity.py: one context over the devices named on the command line,
with whatever tuning environment the test set (e.g. ALACGPU_SMEM_PAD: launch attributes are per device)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from alac.net_b200 import BatchDecoder          # noqa: E402
from oracle import oracle                       # noqa: E402
from tools.alacgen import alacgen as gen        # noqa: E402


def main():
    devices = [int(a) for a in sys.argv[1:]] or [0]
    oracle.build()
    gen.build_encoder()
    tracks = gen.make_config(2, scale=0.02) + gen.make_config(1, scale=0.2)
    with BatchDecoder(devices=devices) as dec:
        for t in tracks:
            dec.add_track(t.cfg, t.mdat, t.stsz)
        pcm, off, ln, status = dec.decode_all()
    assert (status == 0).all()
    for t, o_, l_ in zip(tracks, off, ln):
        ref, _, _ = oracle.decode_track(oracle.cfg_from(t.cfg), t.mdat, t.stsz)
        assert pcm[int(o_):int(o_ + l_)].tobytes() == ref, "PCM differs from the oracle"
    if len(devices) > 1:
        # a multi-device context refuses tracks once its tracks are staged (alacgpu.h, alacgpu_add_track)
        from alac.net_b200 import AlacGpuError
        with BatchDecoder(devices=devices) as dec:
            dec.add_track(tracks[0].cfg, tracks[0].mdat, tracks[0].stsz)
            dec.prepare()
            try:
                dec.add_track(tracks[1].cfg, tracks[1].mdat, tracks[1].stsz)
                raise SystemExit("a staged multi-device context accepted another track")
            except AlacGpuError as e:
                assert e.code == -7, e
            dec.clear()
            dec.add_track(tracks[1].cfg, tracks[1].mdat, tracks[1].stsz)
            pcm, off, ln, status = dec.decode_all()
            ref, _, _ = oracle.decode_track(oracle.cfg_from(tracks[1].cfg), tracks[1].mdat, tracks[1].stsz)
            assert pcm[int(off[0]):int(off[0] + ln[0])].tobytes() == ref
    print("multidev case ok")


if __name__ == "__main__":
    main()
