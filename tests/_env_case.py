"""Helper run in a SUBPROCESS by test_gpu_parity.py: the library reads its tuning environment variables
(ALACGPU_QUAD_MIN_LAST / _FIRST, ALACGPU_NO_TAPER, ...) once per process, so each setting needs its own.
Decodes a few small workloads through the C ABI and compares them with the oracle, byte for byte."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from alac.net_b200 import BatchDecoder          # noqa: E402
from oracle import oracle                       # noqa: E402
from tools.alacgen import alacgen as gen        # noqa: E402


def main():
    oracle.build()
    gen.build_encoder()
    cases = [gen.make_config(2, scale=0.02), gen.make_config(1, scale=0.1), gen.make_config(3, scale=0.2)]
    for tracks in cases:
        for resident in (False, True):
            with BatchDecoder(devices=[0]) as dec:
                for t in tracks:
                    dec.add_track(t.cfg, t.mdat, t.stsz)
                if resident:
                    dec.prepare()
                pcm, off, ln, status = dec.decode_all()
                if os.environ.get("ALACGPU_TEST_INJECT_INTERNAL"):
                    assert dec.timing()["internal_retries"] == 1, "the unfused retry did not run"
            pos = 0
            for t, o_, l_ in zip(tracks, off, ln):
                ref, st, _ = oracle.decode_track(oracle.cfg_from(t.cfg), t.mdat, t.stsz)
                assert np.array_equal(status[pos:pos + t.n_frames], st)
                assert pcm[int(o_):int(o_ + l_)].tobytes() == ref, "PCM differs from the oracle"
                pos += t.n_frames
    print("env case ok")


if __name__ == "__main__":
    main()
