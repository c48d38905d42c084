import sys, numpy as np
sys.path.insert(0, '.')
from alac.net_b200 import BatchDecoder
from tools.alacgen import alacgen as g
from oracle import oracle as o
tracks = [g.make_config(1, scale=0.03)[0], g.make_config(2, scale=0.003)[0], g.make_config(3, scale=0.05)[0]]
rng = np.random.default_rng(1)
sizes = np.array([0, 1, 2, 7, 64, 300, 2000, 0, 9000], dtype=np.uint32)
garbage = rng.integers(0, 256, size=int(sizes.sum()), dtype=np.uint8); garbage[1] &= 0x3F
tracks.append(g.Track(tracks[0].cfg, garbage.tobytes(), sizes, np.zeros(sizes.size, np.int32), b""))
t = tracks[0]; cut = int(np.cumsum(t.stsz)[t.n_frames // 2] - 100)
tracks.append(g.Track(t.cfg, t.mdat[:cut], t.stsz, t.frame_samples, b""))
for kw in ({}, {"chunk_frames": 32}, {"entropy_lanes": 8}):
    with BatchDecoder(devices=[0], **kw) as dec:
        for t in tracks: dec.add_track(t.cfg, t.mdat, t.stsz)
        pcm, off, ln, st = dec.decode_all()
        for t, o_, l_ in zip(tracks, off, ln):
            ref, rst, _ = o.decode_track(o.cfg_from(t.cfg), t.mdat, t.stsz)
            assert pcm[int(o_):int(o_+l_)].tobytes() == ref
        dec.read_frame(0, 3); dec.checksum()
print("SAN_OK")
