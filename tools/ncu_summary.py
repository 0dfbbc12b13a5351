#!/usr/bin/env python
"""Summarise an .ncu-rep of the fused kernel: key raw metrics, SASS opcode mix, stall reasons, top stall lines.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [frames_per_launch]"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 2000 * 998.0

def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout

raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
H, U = raw[0], raw[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.avg', 'launch__grid_size', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
print(f"# {rep}  frames/launch={frames:.0f}")
vals = {}
for k in keys:
    if k in H:
        i = H.index(k)
        vals[k] = raw[2][i]
        print(f"{k:72s} {U[i]:16s} {[r[i] for r in raw[2:]]}")
try:
    ie = float(vals['smsp__inst_executed.sum'].replace(',', ''))
    print(f"warp instructions per frame: {ie / frames:.1f}")
    tr = float(vals['dram__bytes_read.sum']) + float(vals['dram__bytes_write.sum'])
    print(f"dram traffic per frame (unit of the two rows above): {tr / frames * 1e6:.1f} B" if 'Mbyte' in U[H.index('dram__bytes_read.sum')] else tr)
except Exception as e:
    print("derived:", e)

sass = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source=sass"]))))
hi = [i for i, r in enumerate(sass) if r and r[0] == "Address"][0]
SH = sass[hi]; ci = {h: i for i, h in enumerate(SH)}
tot = 0; byop = Counter(); stall = Counter(); samples = 0; lines = []
stall_cols = [h for h in SH if h.startswith('stall_') and 'Not Issued' not in h]
for r in sass[hi + 1:]:
    try:
        n = int(r[ci['Instructions Executed']])
    except Exception:
        continue
    tot += n
    op = [o for o in r[ci['Source']].strip().split() if not o.startswith('@')]
    byop[op[0].split('.')[0] if op else '?'] += n
    s = int(r[ci['# Samples']] or 0)
    samples += s
    lines.append((s, r[ci['Source']].strip()[:70], n))
    for c in stall_cols:
        stall[c] += int(r[ci[c]] or 0)
scale = 1.0
try:
    scale = tot / ie
except Exception:
    pass
print(f"\nSASS-page instruction total {tot} (x{scale:.2f} of smsp__inst_executed.sum); opcode mix per frame (rescaled):")
for k, v in byop.most_common(28):
    print(f"  {k:10s} {v / scale / frames:8.1f}")
print("\nstall samples:")
for k, v in stall.most_common(10):
    print(f"  {k:26s} {100.0 * v / max(samples, 1):5.1f}%")
print("\ntop sampled SASS lines:")
for s, src, n in sorted(lines, reverse=True)[:25]:
    print(f"  {100.0 * s / samples:5.2f}%  x{n / scale / frames:7.2f}/frame  {src}")
