#!/usr/bin/env python
"""Per-source-line warp-instruction counts of an .ncu-rep (sorted by instructions). usage: ncu_lines.py rep [top] [warps]"""
import collections, csv, subprocess, io, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45; warps = float(sys.argv[3]) if len(sys.argv) > 3 else 32768.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, 0]); cur = None; h = None
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': h = r; continue
    if h is None or not r[0].isdigit(): continue
    try:
        iex = int(r[h.index('Instructions Executed')] or 0); tex = int(r[h.index('Thread Instructions Executed')] or 0); smp = int(r[h.index('# Samples')] or 0)
    except Exception: continue
    k = (cur, int(r[0]), r[1].strip()[:90]); agg[k][0] += iex; agg[k][1] += tex; agg[k][2] += smp
tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values())
print("total warp-inst %d (%.0f per warp), samples %d" % (tot, tot / warps, tots))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst (%6.0f/warp) thr %4.1f smp %4.1f%% | %s:%d %s" % (100 * v[0] / tot, v[0] / warps, v[1] / max(v[0], 1), 100 * v[2] / max(tots, 1), k[0], k[1], k[2]))
