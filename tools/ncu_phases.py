#!/usr/bin/env python
"""Executed warp-instructions per frame between consecutive BAR.SYNC instructions of the profiled kernel (SASS order),
with the opcode mix of each segment. Usage: ncu_phases.py rep [frames]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]; frames = float(sys.argv[2]) if len(sys.argv) > 2 else 1996000.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H = None; seg = []; cur = [0, 0, Counter(), None, Counter()]
STALLS = None
for r in rows:
    if not r: continue
    if "Source" in r and "Instructions Executed" in r:
        H = {h: i for i, h in enumerate(r)}
        STALLS = [h for h in r if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if H is None or len(r) < len(H): continue
    try: n = int(r[H["Instructions Executed"]] or 0); s = int(r[H["# Samples"]] or 0)
    except ValueError: continue
    src = r[H["Source"]].strip(); t = src.split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0] if t else "?"
    if cur[3] is None: cur[3] = r[H["Address"]] if "Address" in H else "?"
    cur[0] += n; cur[1] += s; cur[2][op] += n
    for h in STALLS:
        try: cur[4][h[6:]] += int(r[H[h]] or 0)
        except ValueError: pass
    if src.startswith("BAR.SYNC") or "BAR.SYNC" in src:
        seg.append(cur); cur = [0, 0, Counter(), None, Counter()]
seg.append(cur)
tot = sum(c[0] for c in seg); ts = sum(c[1] for c in seg)
print(f"total {tot/frames:.1f} warp-inst/frame, {ts} samples")
for i, (n, s, ops, a, st) in enumerate(seg):
    print(f"seg {i:2d} @{a}: {n/frames:7.1f} inst/frame  {100*s/max(ts,1):5.1f}% samples   " +
          " ".join(f"{k}:{v/frames:.1f}" for k, v in ops.most_common(12)))
    if s * 50 > ts:
        print("        stalls (% of the segment's samples): " + " ".join(f"{k}:{100*v/max(s,1):.0f}" for k, v in st.most_common(8)))
