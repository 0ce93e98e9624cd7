#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full --import-source on) into a small text summary for profiles/:
key raw metrics per captured launch + executed-instruction mix + stall reasons + hottest SASS lines.
usage: python tools/ncu_export.py gpurun_out/prof.ncu-rep profiles/name.txt [cells_per_launch]"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
cells = sys.argv[3] if len(sys.argv) > 3 else None
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "launch__shared_mem_per_block_dynamic",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = [f"# summary of {rep} (ncu --set full --clock-control none; per captured launch)"]
for r in rows[2:]:
    lines.append("")
    lines.append("kernel: " + r[hdr.index("Kernel Name")])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"  {k} = {r[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
tmp = out + ".src.csv"
open(tmp, "w").write(src)
summ = subprocess.run([sys.executable, __file__.replace("ncu_export.py", "ncu_src_summary.py"), tmp] + ([cells] if cells else []),
                      capture_output=True, text=True).stdout
import os
os.remove(tmp)
lines.append("")
lines.append("# executed instruction mix / stall reasons / hottest SASS lines (source page of the first captured launch)")
lines.append(summ)
open(out, "w").write("\n".join(lines))
print("wrote", out)
