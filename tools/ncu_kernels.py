#!/usr/bin/env python3
"""Per-launch table of an ncu report (developer tool): duration, grid, registers, issue / tensor-pipe / DRAM numbers of every captured launch.
usage: python tools/ncu_kernels.py report.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"), ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("lts__t_bytes.sum", "L2 bytes"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %")]
idx = [(hdr.index(k), n) for k, n in cols if k in hdr]
units = rows[1]
print("| " + " | ".join(n + (f" [{units[i]}]" if units[i] and n not in ("kernel",) else "") for i, n in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    vals = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.replace("(anonymous namespace)::", "").split("(")[0][:26]
        else:
            try:
                f = float(v.replace(",", "")); v = f"{f:.1f}" if abs(f) < 1e4 else f"{f:.3g}"
            except ValueError:
                pass
        vals.append(v)
    print("| " + " | ".join(vals) + " |")
