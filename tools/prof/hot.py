#!/usr/bin/env python
"""hot.py <source.csv> [min_samples] [lo hi] -- per-instruction stall samples from `ncu --page source --csv`."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
mins=int(sys.argv[2]) if len(sys.argv)>2 else 25
lo=int(sys.argv[3],16) if len(sys.argv)>3 else 0
hi=int(sys.argv[4],16) if len(sys.argv)>4 else 1<<30
iS=hdr.index('# Samples'); iE=hdr.index('Instructions Executed'); iT=hdr.index('Avg. Threads Executed')
src=hdr.index('Source'); ia=hdr.index('Address')
cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data=rows[2:]
base=int(data[0][ia],16)
tot=sum(int(r[iS]) for r in data)
agg={}
for r in data:
    for i in cols:
        if r[i] not in ('','0'): agg[hdr[i]]=agg.get(hdr[i],0)+int(r[i])
print('total samples',tot,'instrs',len(data), sorted(agg.items(), key=lambda kv:-kv[1])[:8])
for r in data:
    a=int(r[ia],16)-base
    if lo<=a<=hi and int(r[iS])>=mins:
        st={hdr[i][6:]:int(r[i]) for i in cols if r[i] not in ('','0')}
        top=sorted(st.items(), key=lambda kv:-kv[1])[:2]
        print(f"{a:05x} {int(r[iS]):5d} {int(r[iE]):8d} {r[iT]:>3s} {r[src][:58]:58s} {top}")
