#!/usr/bin/env python
"""Per-CUDA-source-line summary of an ncu report (needs -lineinfo + --import-source on):
   python profiles/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [min_pct]"""
import csv, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
seen = 0
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name":
        seen += 1
        if seen > 1:
            rows = rows[:i]
            break
    if r and r[0] == "Line No" and len(r) > 8:
        hdr = r
lines = [r for r in rows if hdr and len(r) == len(hdr) and r[0].isdigit()]
ci = {n: hdr.index(n) for n in ("# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
                                "Warp Stall Sampling (All Samples)")}
gi = [k for k, n in enumerate(hdr) if n == "L2 Theoretical Sectors Global"]
tot_i = sum(int(r[ci["Instructions Executed"]] or 0) for r in lines)
tot_s = sum(int(r[ci["# Samples"]] or 0) for r in lines)
tot_w = sum(int(r[ci["L1 Wavefronts Shared"]] or 0) for r in lines)
print("kernel %s: instr=%d samples=%d smem_wavefronts=%d" % (kern, tot_i, tot_s, tot_w))
print("%6s %6s %6s %8s  line" % ("inst%", "smp%", "wf%", "wf/ideal"))
for r in lines:
    ni, ns, nw = int(r[ci["Instructions Executed"]] or 0), int(r[ci["# Samples"]] or 0), int(r[ci["L1 Wavefronts Shared"]] or 0)
    nwi = int(r[ci["L1 Wavefronts Shared Ideal"]] or 0)
    pi, ps, pw = 100.0 * ni / max(tot_i, 1), 100.0 * ns / max(tot_s, 1), 100.0 * nw / max(tot_w, 1)
    if max(pi, ps, pw) >= minpct:
        print("%6.1f %6.1f %6.1f %8s  L%-4s %s" % (pi, ps, pw, ("%.1f" % (nw / nwi)) if nwi else "-", r[0], r[1].strip()[:100]))
