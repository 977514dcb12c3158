#!/usr/bin/env python
"""Summarise ncu reports (gpurun_out/*.ncu-rep, scratch) into tracked files:
   python profiles/summarize.py OUT.md REPORT.ncu-rep [REPORT2 ...]
writes OUT.md (one block per captured kernel) and updates profiles/traffic.json
(DRAM bytes per score column of each kernel, from dram__bytes_read/write.sum)."""
import csv, json, os, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/smem throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}

out_md, reps = sys.argv[1], sys.argv[2:]
cols_per_launch = int(os.environ.get("COLS", "296"))
here = os.path.dirname(os.path.abspath(__file__))
tpath = os.path.join(here, "traffic.json")
traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
lines = []
seen = {}
for rep in reps:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        short = name.split("(")[0].replace("void ", "").split("<")[0]
        lines.append("### `%s`  (%s, %d score columns per launch)\n" % (short, os.path.basename(rep), cols_per_launch))
        lines.append("| metric | value |\n|---|---|")
        dram = 0.0
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                lines.append("| %s | %s %s |" % (label, r[i], units[i]))
                if key.startswith("dram__bytes"):
                    dram += float(r[i]) * UNIT.get(units[i], 1.0)
        lines.append("")
        if short in seen and seen[short] >= dram:
            continue  # keep the heaviest launch of a kernel (e.g. the first ranking pass, not the re-rank)
        seen[short] = dram
        traffic[short] = {"dram_bytes_per_column": dram / cols_per_launch, "columns_in_capture": cols_per_launch,
                          "source": os.path.basename(rep) + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
        pipes = {}
        for key, label in (("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
                           ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_smem_pct"),
                           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
                           ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
                           ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
                           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct")):
            if key in hdr:
                try:
                    pipes[label] = round(float(r[hdr.index(key)]), 1)
                except ValueError:
                    pass
        traffic[short]["pipes"] = pipes
open(out_md, "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(tpath, "w"), indent=1)
print("wrote", out_md, "and", tpath)
