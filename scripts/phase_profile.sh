#!/bin/bash
# phase_profile.sh <ncu-rep> <kernel-substr> : per-function share of samples / instructions in aba_deriv_coop.cuh
ncu -i $1 --page source --csv > /tmp/_src.csv 2>/dev/null
python scripts/ncu_summary.py $1 | python -c "
import csv,sys
r=list(csv.reader(sys.stdin))
for h,v in zip(r[0],r[1]):
    if any(k in h for k in ('duration','inst_executed.sum','registers','dynamic','fp64_cycles','stalled','dram__bytes','bank_conflicts','wavefronts')): print(h,'=',v)"
python - "$2" <<'PY'
import re,sys,subprocess,collections
src=open('pinocchio_b200/csrc/aba_deriv_coop.cuh').read().splitlines()
# function ranges
starts=[(k+1,re.search(r'(\w+)\(',l).group(1)) for k,l in enumerate(src) if l.startswith('BRBD_DI') or l.startswith('aba_derivatives_coop_kernel(')]
out=subprocess.run(['python','scripts/line_profile.py','/tmp/_src.csv',sys.argv[1],'aba_deriv_coop.cuh','1000'],capture_output=True,text=True,env=dict(__import__('os').environ,INNER='1')).stdout
S=collections.Counter();I=collections.Counter()
for l in out.splitlines():
    m=re.match(r'aba_deriv_coop.cuh:(\d+)\s+samples\s+([\d.]+)%.*inst\s+([\d.]+)%',l)
    if not m: continue
    ln=int(m.group(1)); name='?'
    for a,n in starts:
        if ln>=a: name=n
    if name=='aba_derivatives_coop_config':
        t=src[ln-1]; mm=re.search(r'(coop_\w+)<',t); name='cfg:'+(mm.group(1) if mm else 'other')
    S[name]+=float(m.group(2)); I[name]+=float(m.group(3))
for n,_ in S.most_common(): print(f'{n:40s} samples {S[n]:5.1f}%  inst {I[n]:5.1f}%')
PY
