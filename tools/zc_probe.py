import sys, time, numpy as np
sys.path.insert(0,'.')
from alac.net_b200 import BatchDecoder, PinnedBuffer
from tools.alacgen import alacgen as g
t = g.make_config(2, scale=1.0)[0]
pb = PinnedBuffer(len(t.mdat)); pb.array[:] = np.frombuffer(t.mdat, np.uint8)
for flags in (0, 8, 4):
    with BatchDecoder(devices=[0], flags=flags) as dec:
        dec.add_track(t.cfg, pb, t.stsz)
        total = dec.prepare()
        out = PinnedBuffer(total)
        for mode, dst in (("resident", False), ("pinned", out)):
            ts=[]
            for i in range(6):
                t0=time.perf_counter(); dec.decode_all(dst, want_status=False); ts.append((time.perf_counter()-t0)*1e3)
            tm=dec.timing()
            print(f"flags={flags} {mode}: wall {np.median(ts):.2f} ms  kernels {tm['kernels_ms']:.2f} d2h {tm['d2h_ms']:.2f} chunks {tm['chunks']}")
        assert out.array[:len(t.pcm)].tobytes()==t.pcm
