#!/usr/bin/env python
"""e2e_probe.py -- one end-to-end step on configs[1] with the pipeline schedule printed (ALACGPU_HOST_TIMING)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alac.net_b200 import BatchDecoder, PinnedBuffer
from tools.alacgen import alacgen as g
g.build_encoder()
tr = g.track_24_stereo(g.SEED_BASE + 2, 600.0)
pb = PinnedBuffer(len(tr.mdat)); pb.array[:] = np.frombuffer(tr.mdat, dtype=np.uint8)
dec = BatchDecoder(devices=[0], flags=int(os.environ.get("FLAGS", "0")), chunk_frames=int(os.environ.get("CHUNK", "0")))
dec.add_track(tr.cfg, pb, tr.stsz)
out = PinnedBuffer(dec.total_pcm_bytes())
for it in range(4):
    if it == 3:
        os.environ["ALACGPU_HOST_TIMING"] = "1"
    t0 = time.perf_counter()
    dec.clear()
    t1 = time.perf_counter()
    dec.add_track(tr.cfg, pb, tr.stsz)
    t2 = time.perf_counter()
    dec.decode_all(out, want_status=False)
    t3 = time.perf_counter()
    print(f"step {it}: clear {1e3*(t1-t0):.2f} add_track {1e3*(t2-t1):.2f} decode_all {1e3*(t3-t2):.2f} ms", flush=True)
dec.timing()
assert out.array[:len(tr.pcm)].tobytes() == tr.pcm
