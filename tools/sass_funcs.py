#!/usr/bin/env python3
"""Per-function SASS instruction counts of the env kernels (developer tool): python tools/sass_funcs.py [16|32]."""
import re, subprocess, sys, os, tempfile
lanes = sys.argv[1] if len(sys.argv) > 1 else "16"
so = os.path.join(os.path.dirname(__file__), "..", "robosuite_benchmark_b200", "csrc", "librsb_cuda.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, "rsb_cuda16.sm_100a.cubin" if lanes == "16" else "rsb_cuda.sm_100a.cubin")
sass = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout
cur, counts, order, in_step = None, {}, [], False
for line in sass.splitlines():
    m = re.match(r"^(\S+):\s*$", line)
    if line.startswith("//---") and ".text." in line:
        in_step = "6k_step" in line
        cur = "k_step(main)" if in_step else None
        if cur: counts[cur] = 0; order.append(cur)
        continue
    if not in_step: continue
    if m and not m.group(1).startswith(".L_"):
        name = m.group(1)
        mm = re.search(r"\$_ZN\d+_INTERNAL_[0-9a-z_]+?cu_[0-9a-f]{8}(\d+)([A-Za-z_0-9]+)", name)
        if mm: cur = mm.group(2)[:int(mm.group(1))]
        elif "internal" in name: cur = name.split("$")[-1][:40]
        else: continue
        counts.setdefault(cur, 0); order.append(cur)
        continue
    if cur and re.match(r"^\s*/\*[0-9a-f]{4,}\*/", line): counts[cur] += 1
tot = sum(counts.values())
for k in dict.fromkeys(order): print(f"{k:28s} {counts[k]:6d}")
print(f"{'total':28s} {tot:6d}  ({tot*16/1024:.0f} KB)")
