#!/usr/bin/env python
"""Joins an ncu source-page CSV (`ncu -i x.ncu-rep --page source --csv`) with `nvdisasm --print-line-info` of the same
cubin and prints, per block of N SASS instructions of one kernel: warp-instructions executed per unit of work, share of the
stall samples, the source-line range and the three dominant stall reasons. Development aid (how the phases of the traversal
kernel were weighed).

  python tools/ncu_blocks.py <kernel.sass> <source.csv> <mangled kernel name> <units of work> [block size]
"""
import csv,re,sys,collections
sass_path, csv_path, kernel, per = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
B = int(sys.argv[5]) if len(sys.argv)>5 else 100
lines=open(sass_path).read().split("\n")
start=[i for i,l in enumerate(lines) if l.startswith(".text."+kernel+":")][0]
cur=None; seq=[]
for l in lines[start+1:]:
    if l.startswith("//-----") or l.startswith(".text."): break
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split("/")[-1],int(m.group(2))); continue
    m=re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);",l)
    if m: seq.append(cur)
rows=list(csv.reader(open(csv_path))); hdr=rows[1]
ie=hdr.index("Instructions Executed"); si=hdr.index("# Samples")
stalls=[(i,h) for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data=[r for r in rows[2:] if len(r)==len(hdr)]
tot=sum(int(r[ie]) for r in data); stot=sum(int(r[si]) for r in data)
print("total instr/unit %.0f samples %d"%(tot/per, stot))
for b in range(0,len(data),B):
    blk=data[b:b+B]
    ins=sum(int(r[ie]) for r in blk); smp=sum(int(r[si]) for r in blk)
    if ins/per<5 and smp/stot<0.003: continue
    st=collections.Counter()
    for r in blk:
        for i,h in stalls: st[h]+=int(r[i])
    lc=collections.Counter()
    for i in range(b,min(b+B,len(seq))):
        if seq[i] and seq[i][0] in ('search.cuh','search_fast.cuh'): lc[seq[i][1]]+=1
    lns=sorted(lc)
    rng="%d-%d"%(lns[0],lns[-1]) if lns else "-"
    top=" ".join("%s=%.0f%%"%(k.replace('stall_',''),100*v/max(smp,1)) for k,v in st.most_common(3))
    print("%5d  ins/unit %6.0f  smp %5.1f%%  lines %-10s %s"%(b,ins/per,100*smp/stot,rng,top))
