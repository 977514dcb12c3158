#!/usr/bin/env python
"""One markdown block per kernel launch of an ncu report (raw page), without touching traffic.json:
   python profiles/ncu_table.py REPORT.ncu-rep > OUT.md"""
import csv, subprocess, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
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
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA sub-pipe %"),
    ("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "shared pipe %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/smem throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "s": 1.0, "ns": 1e-9}
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        short = name.split("(")[0].replace("void ", "")
        print("### `%s`  (%s, launch id %s)\n" % (short, rep.split("/")[-1], r[hdr.index("ID")]))
        print("| metric | value |\n|---|---|")
        dram, dur = 0.0, None
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print("| %s | %s %s |" % (label, r[i], units[i]))
                if key.startswith("dram__bytes"):
                    dram += float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
                if key == "gpu__time_duration.sum":
                    dur = float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        if dur:
            print("| DRAM bytes / duration | %.1f GB/s |" % (dram / dur / 1e9))
        print()
