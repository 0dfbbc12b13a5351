#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares from an .ncu-rep (needs -lineinfo). Usage: ncu_lines.py rep [frames]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; frames = float(sys.argv[2]) if len(sys.argv) > 2 else 1996000.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = "?"; H = None
agg = defaultdict(lambda: [0, 0, ""])   # (file,line) -> [inst, samples, text]
tot_i = tot_s = 0
for r in rows:
    if not r: continue
    if r[0] in ("File Name", "File Path"): fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": H = {h: i for i, h in enumerate(r)}; ia = r.index("Address"); continue
    if H is None or len(r) < 8: continue
    try:
        ln = int(r[0]); n = int(r[H["Instructions Executed"]] or 0); s = int(r[H["# Samples"]] or 0)
    except Exception:
        continue
    if r[ia] == "": # a pure source row (aggregated) -> skip, we sum SASS rows
        agg[(fname, ln)][2] = r[1].strip()[:90]
        continue
    a = agg[(fname, ln)]; a[0] += n; a[1] += s; tot_i += n; tot_s += s
print(f"total warp-inst/frame {tot_i/frames:.1f}, samples {tot_s}")
byfile = defaultdict(lambda: [0, 0])
for (f, l), (n, s, t) in agg.items():
    byfile[f][0] += n; byfile[f][1] += s
for f, (n, s) in byfile.items():
    print(f"{f:20s} inst/frame {n/frames:8.1f}  samples {100*s/max(tot_s,1):5.1f}%")
print()
for (f, l), (n, s, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{f:16s}:{l:4d} inst/frame {n/frames:7.1f} samples {100*s/max(tot_s,1):5.1f}%  {t}")
