#!/usr/bin/env python
"""Stall-reason totals from `ncu --page source --csv` (SASS view)."""
import csv, sys
rows=[r for r in csv.reader(open(sys.argv[1]))]
hdr=rows[1]; data=[r for r in rows[2:] if len(r)==len(hdr) and r[0]!='Address']
tots={}
for h in hdr:
    if h.startswith('stall_') and 'Not Issued' not in h:
        i=hdr.index(h); tots[h]=sum(int(r[i]) for r in data)
s=sum(tots.values())
for h,v in sorted(tots.items(), key=lambda x:-x[1])[:9]: print(f'{h:28s} {v:8d} {100*v/s:5.1f}%')
