import sys, numpy as np
sys.path.insert(0,'.')
from alac.net_b200 import BatchDecoder
from tools.alacgen import alacgen as g
from oracle import oracle as o
t = g.make_config(1, scale=0.05)[0]
ref, st, fb = o.decode_track(o.cfg_from(t.cfg), t.mdat, t.stsz)
for flags in (2, 2|16):
    with BatchDecoder(devices=[0], flags=flags) as dec:
        dec.add_track(t.cfg, t.mdat, t.stsz)
        pcm, off, ln, status = dec.decode_all()
    got = pcm[:len(ref)].tobytes()
    a = np.frombuffer(got, '<i2').reshape(-1,2); b = np.frombuffer(ref, '<i2').reshape(-1,2)
    print("flags", flags, "equal", got==ref, "status", status[:8])
    if got != ref:
        pos=0
        for f in range(t.n_frames):
            n=int(t.frame_samples[f]); fa=a[pos:pos+n]; fbb=b[pos:pos+n]; pos+=n
            bad=np.nonzero((fa!=fbb).any(axis=1))[0]
            fr=t.frames[f]
            if bad.size:
                print("frame",f,"orders",fr['order'],"quant",fr['quant'],"mix",fr['mix_shift'],fr['mix_weight'],"first bad sample",bad[0],"nbad",bad.size, "L/R bad", (fa[:,0]!=fbb[:,0]).sum(), (fa[:,1]!=fbb[:,1]).sum())
