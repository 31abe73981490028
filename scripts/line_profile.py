#!/usr/bin/env python
"""Per-source-line profile of one kernel: joins `ncu --page source --csv` (SASS view, per-instruction samples)
with `nvdisasm -gi` line info of the built library (outermost frame in the given source file).

usage: line_profile.py <ncu_source.csv> <kernel-substring> <source-file-basename> [top]
"""
import csv, collections, os, re, subprocess, sys, tempfile
ncu_csv, kern, srcfile = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "pinocchio_b200", "csrc", "libpinocchio_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# locate function
start = None
for k, l in enumerate(dis):
    if l.startswith("//---") and ".text." in l and kern in l:
        start = k; break
assert start is not None, "kernel not found in disassembly"
off2line = {}
cur = None
inner = os.environ.get("INNER") == "1"   # INNER=1: innermost frame in the file (default: outermost)
fresh = True
for l in dis[start + 1:]:
    if l.startswith("//---") and ".text." in l: break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if os.path.basename(m.group(1)) == srcfile and (not inner or fresh):
            cur = int(m.group(2))   # last such frame = outermost in that file
            fresh = False
        continue
    m = re.match(r'\s*/\*([0-9a-f]+)\*/', l)
    if m:
        off2line[int(m.group(1), 16)] = cur
        fresh = True
rows = [r for r in csv.reader(open(ncu_csv))]
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]
iA, iN, iS = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iL = hdr.index("stall_long_sb")
base = int(data[0][iA], 16)
# the csv may list the kernel more than once; keep the first copy
n = len(off2line)
data = data[:n] if len(data) >= n else data
inst = collections.Counter(); smp = collections.Counter(); lsb = collections.Counter()
for r in data:
    ln = off2line.get(int(r[iA], 16) - base)
    inst[ln] += int(r[iN]); smp[ln] += int(r[iS]); lsb[ln] += int(r[iL])
ti, ts = sum(inst.values()), sum(smp.values())
src = open(os.path.join(root, "pinocchio_b200", "csrc", srcfile)).read().splitlines()
print(f"total warp-inst {ti}  samples {ts}")
for ln, s in smp.most_common(top):
    text = src[ln - 1].strip()[:90] if ln else "?"
    print(f"{srcfile}:{ln}  samples {100*s/ts:5.1f}%  (long_sb {100*lsb[ln]/ts:4.1f}%)  inst {100*inst[ln]/ti:5.1f}%  | {text}")
