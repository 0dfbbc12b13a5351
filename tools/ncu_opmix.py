#!/usr/bin/env python
"""Opcode mix per source file / hot line from an .ncu-rep (cuda,sass correlated view). Usage: ncu_opmix.py rep [frames]"""
import csv, io, subprocess, sys
from collections import defaultdict, Counter
rep=sys.argv[1]; frames=float(sys.argv[2]) if len(sys.argv)>2 else 1996000.0
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source=cuda,sass"],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
fname='?'; H=None; cur=None; byfile=defaultdict(Counter); byline=defaultdict(Counter)
for r in rows:
    if not r: continue
    if r[0] in("File Name","File Path"): fname=r[1].split('/')[-1]; continue
    if r[0]=="Line No": H={h:i for i,h in enumerate(r)}; continue
    if H is None or len(r)<8: continue
    if r[0]!="":
        try: cur=int(r[0])
        except Exception: cur=None
        continue
    if not r[2].startswith("0x"): continue
    try: n=int(r[H["Instructions Executed"]] or 0)
    except Exception: continue
    op=[o for o in r[3].strip().split() if not o.startswith('@')]
    opc=op[0].split('.')[0] if op else '?'
    byfile[fname][opc]+=n; byline[(fname,cur)][opc]+=n
for f,c in byfile.items():
    tot=sum(c.values())
    if tot/frames<1: continue
    print(f, "total %.1f/frame"%(tot/frames)); print("  "+"  ".join(f"{k}:{v/frames:.1f}" for k,v in c.most_common(24)))
print()
for key,c in sorted(byline.items(), key=lambda kv:-sum(kv[1].values()))[:int(sys.argv[3]) if len(sys.argv)>3 else 14]:
    print(key[0][4:9], key[1], "%.1f"%(sum(c.values())/frames), "  ".join(f"{k}:{v/frames:.1f}" for k,v in c.most_common(8)))
