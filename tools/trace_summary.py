#!/usr/bin/env python
"""trace_summary.py <trace file written with ALACGPU_TRACE=...> -- when each role of the fused launch starts and ends."""
import sys
import numpy as np
hdr = open(sys.argv[1]).readline().split()
eb, grid, n1, n4 = int(hdr[2]), int(hdr[4]), int(hdr[6]), int(hdr[8])
d = np.loadtxt(sys.argv[1], dtype=np.int64)
d = d[d[:, 1] > 0]
t0 = d[:, 1].min()
w, s, e = d[:, 0], (d[:, 1] - t0) / 1e3, (d[:, 2] - t0) / 1e3
qw = (n4 + 7) // 8
roles = {"entropy": w < eb * 4, "lpc workers": w >= eb * 4}
print(f"eblocks {eb} grid {grid} one-lane streams {n1} quad streams {n4}; total {e.max():.0f} us")
for name, m in roles.items():
    if not m.any():
        continue
    dur = e[m] - s[m]
    live = dur > 5
    print(f"{name:13s} warps {m.sum():5d} (working {live.sum():5d}) start {s[m].min():7.0f}..{s[m].max():7.0f} us  end p50 {np.percentile(e[m][live],50):7.0f} p90 {np.percentile(e[m][live],90):7.0f} max {e[m][live].max():7.0f} us")
if d.shape[1] > 3:      # per-SM view: which warps an SM hosted and when it went idle
    sm = d[:, 3]
    ends = np.array([e[sm == k].max() for k in range(int(sm.max()) + 1) if (sm == k).any()])
    nw = np.array([((sm == k) & ((e - s) > 5)).sum() for k in range(int(sm.max()) + 1) if (sm == k).any()])
    print(f"SMs {len(ends)}: last warp ends p10 {np.percentile(ends,10):.0f} p50 {np.percentile(ends,50):.0f} p90 {np.percentile(ends,90):.0f} max {ends.max():.0f} us; working warps per SM min {nw.min()} max {nw.max()}")
    busy = np.array([(e[sm == k] - s[sm == k]).sum() for k in range(int(sm.max()) + 1) if (sm == k).any()])
    print(f"warp-us per SM: min {busy.min():.0f} p50 {np.percentile(busy,50):.0f} max {busy.max():.0f}")
    blk = (w // 4).astype(int)
    first = {int(b): int(sm[blk == b][0]) for b in np.unique(blk)[:300:37]}
    print("block -> SM samples:", first)
if len(sys.argv) > 2:   # one-lane warps in work-list order (heaviest first): end time per decile
    m = roles["lpc workers"]
    ee = e[m][(e[m] - s[m]) > 5]
    for k in range(0, len(ee), max(1, len(ee) // 20)):
        print(f"  one-lane warp #{k}: end {ee[k]:.0f} us")
