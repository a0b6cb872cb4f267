#!/usr/bin/env python3
"""Basic-block execution profile of a kernel from an .ncu-rep (SASS page): for every run of
instructions with the same execution count, its share of the executed warp instructions, the
average active threads and its share of the stall samples.  usage: ncu_blocks.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))

hdr=rows[1]; ie=hdr.index("Instructions Executed"); te=hdr.index("Thread Instructions Executed"); sm=hdr.index("# Samples")
data=[(i,r[1].strip(),float(r[ie]),float(r[te]),float(r[sm])) for i,r in enumerate(rows[2:])]
tot=sum(d[2] for d in data); stot=sum(d[4] for d in data)
print("total",tot, "samples", stot)
# segment into runs of (approximately) equal exec count: basic blocks
segs=[]; cur=None
for d in data:
    if cur is None or abs(d[2]-cur['cnt'])>0.02*max(cur['cnt'],1):
        cur={'start':d[0],'cnt':d[2],'n':0,'inst':0,'tinst':0,'smp':0,'ops':{}}
        segs.append(cur)
    cur['n']+=1; cur['inst']+=d[2]; cur['tinst']+=d[3]; cur['smp']+=d[4]; cur['end']=d[0]
    op=d[1].split()[0] if not d[1].startswith('@') else d[1].split()[1]
    op=op.split('.')[0]
    cur['ops'][op]=cur['ops'].get(op,0)+1
cum=0
for s in segs:
    cum+=s['inst']
    if s['inst']/tot>0.003:
        ops=sorted(s['ops'].items(),key=lambda kv:-kv[1])[:8]
        print(f"{s['start']:5d}-{s['end']:5d} n={s['n']:4d} cnt={s['cnt']:10.0f} inst={100*s['inst']/tot:5.2f}% cum={100*cum/tot:5.1f}% thr/inst={s['tinst']/max(s['inst'],1):5.1f} smp={100*s['smp']/stot:5.2f}% ", ' '.join(f"{k}:{v}" for k,v in ops))
