#!/usr/bin/env python3
"""Summarise an ncu report of k_step (developer tool): key counters, stall mix, per-function instruction/sample shares.
usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [16|32]"""
import csv, io, re, subprocess, sys, bisect, os
rep = sys.argv[1]; lanes = sys.argv[2] if len(sys.argv) > 2 else "16"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
print("| metric | value | unit |\n|---|---|---|")
for k in keys:
    if k in hdr: i = hdr.index(k); print(f"| {k} | {vals[i]} | {units[i]} |")
st = {}
for i, h in enumerate(hdr):
    if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
        try: st[h[33:]] = float(vals[i])
        except ValueError: pass
tot = sum(st.values()) or 1
print("\nWarp-stall sampling: " + ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; data = rows[2:]
ia, ii, isamp, isrc, ith = (hdr.index(x) for x in ('Address', 'Instructions Executed', '# Samples', 'Source', 'Thread Instructions Executed'))
cols = {k: hdr.index(k) for k in ['stall_no_inst', 'stall_wait', 'stall_short_sb', 'stall_barrier', 'stall_long_sb', 'stall_branch_resolving']}
base = int(data[0][ia], 16); entries = {base}
for r in data:
    m = re.search(r'CALL\.REL\.NOINC (0x[0-9a-f]+)', r[isrc])
    if m: entries.add(int(m.group(1), 16))
entries = sorted(entries)
# names from the current build (same order as in the cubin)
names = []
try:
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "sass_funcs.py"), lanes], capture_output=True, text=True).stdout
    names = [(l.split()[0], int(l.split()[1])) for l in out.splitlines() if l.split() and l.split()[0] != 'total']
except Exception: pass
agg = {}
for r in data:
    a = int(r[ia], 16); e = entries[bisect.bisect_right(entries, a) - 1]
    d = agg.setdefault(e, dict(n=0, inst=0, samp=0, th=0, **{k: 0 for k in cols}))
    d['n'] += 1; d['inst'] += int(r[ii] or 0); d['samp'] += int(r[isamp] or 0); d['th'] += int(r[ith] or 0)
    for k, c in cols.items(): d[k] += int(r[c] or 0)
ti = sum(d['inst'] for d in agg.values()); ts = sum(d['samp'] for d in agg.values())
bysize = {}
for nm, n in names: bysize.setdefault(n, []).append(nm)
print("\n| function | SASS | inst % | samples % | thr/inst | no_inst | wait | short_sb | barrier | long_sb | branch |\n|---|---|---|---|---|---|---|---|---|---|---|")
for e, d in sorted(agg.items()):
    s_ = max(d['samp'], 1); nm = (bysize.get(d['n']) or [hex(e - base)])[0]
    print(f"| {nm} | {d['n']} | {100*d['inst']/ti:.1f} | {100*d['samp']/ts:.1f} | {d['th']/max(d['inst'],1):.1f} | " + " | ".join(str(round(100 * d[k] / s_)) for k in cols) + " |")
print(f"\ntotal warp-instructions {ti}, samples {ts}")
