#!/usr/bin/env python
"""Summarise `ncu --page source --csv` (SASS view): opcode mix and hottest contiguous regions."""
import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1]))]
hdr=rows[1]; data=[r for r in rows[2:] if len(r)==len(hdr) and r[0]!="Address"]
iS=hdr.index('Source'); iN=hdr.index('Instructions Executed'); iSm=hdr.index('# Samples')
tot=sum(int(r[iN]) for r in data); tots=sum(int(r[iSm]) for r in data)
print('total warp-inst', tot, 'sass lines', len(data), 'samples', tots)
mix=collections.Counter(); smp=collections.Counter()
for r in data:
    t=r[iS].split()
    op=t[1] if t[0].startswith('@') else t[0]
    op=op.split('.')[0]
    mix[op]+=int(r[iN]); smp[op]+=int(r[iSm])
for op,n in mix.most_common(22): print(f'{op:10s} {n:12d} {100*n/tot:5.1f}%  samples {100*smp[op]/max(tots,1):5.1f}%')
print()
prev=None; start=0; regions=[]
for k,r in enumerate(data):
    n=int(r[iN])
    if prev is None or abs(n-prev)>0.02*max(n,prev,1):
        if prev is not None: regions.append((start,k-1,prev))
        start=k
    prev=n
regions.append((start,len(data)-1,prev))
regions=[(a,b,n,(b-a+1)*n, sum(int(data[k][iSm]) for k in range(a,b+1))) for a,b,n in regions]
for a,b,n,w,s in sorted(regions,key=lambda x:-x[4])[:int(sys.argv[2]) if len(sys.argv)>2 else 14]:
    ops=collections.Counter()
    for k in range(a,b+1):
        t=data[k][iS].split(); op=t[1] if t[0].startswith('@') else t[0]; ops[op.split('.')[0]]+=1
    print(f'sass {a}-{b} ({b-a+1} instr) x {n} exec = {100*w/tot:.1f}% inst, {100*s/max(tots,1):.1f}% samples | {dict(ops.most_common(6))}')
