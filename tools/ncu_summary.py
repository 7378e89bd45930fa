#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + instruction / stall-sample share per source line.
usage: tools/ncu_summary.py report.ncu-rep [top_n]"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic']
for w in want:
    if w in hdr:
        i = hdr.index(w); print("%-70s %-14s %s" % (w, units[i], [r[i][:60] for r in data]))
stall = [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct')]
vals = sorted(((float(data[0][hdr.index(h)] or 0), h) for h in stall), reverse=True)[:8]
for v, h in vals: print("  stall %-60s %.1f" % (h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, 0]); cur = None; h = None
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': h = r; continue
    if h is None or not r[0].isdigit(): continue
    try:
        iex = int(r[h.index('Instructions Executed')] or 0); tex = int(r[h.index('Thread Instructions Executed')] or 0)
        smp = int(r[h.index('# Samples')] or 0)
    except Exception: continue
    k = (cur, int(r[0]), r[1].strip()[:100]); agg[k][0] += iex; agg[k][1] += tex; agg[k][2] += smp
tot = sum(v[0] for v in agg.values()) or 1; tots = sum(v[2] for v in agg.values()) or 1
print("source lines: total warp-inst %d, samples %d" % (tot, tots))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print("%5.1f%% smp %5.1f%% inst thr %4.1f | %s:%d %s" % (100 * v[2] / tots, 100 * v[0] / tot, v[1] / max(v[0], 1), k[0], k[1], k[2]))
