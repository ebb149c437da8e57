"""Prints a compact summary of an .ncu-rep (raw page) -- the numbers quoted in profiles/*.md.
usage: python scripts/ncu_summary.py file.ncu-rep [more_regex]"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^gpu__dram_throughput\.avg\.pct",
    r"^launch__registers_per_thread$", r"^launch__occupancy_limit", r"^launch__grid_size$", r"^launch__block_size$",
    r"^launch__shared_mem_per_block_dynamic$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__inst_executed\.sum$", r"^smsp__issue_active\.avg\.pct", r"^sm__inst_executed_pipe_fma\.avg\.pct",
    r"^sm__pipe_fma_cycles_active\.avg\.pct_of_peak_sustained_active", r"^sm__inst_executed_pipe_lsu",
    r"^l1tex__data_pipe_lsu_wavefronts\.avg\.pct", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared(_op_(ld|st))?\.sum$",
    r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_(ld|st))?\.sum$", r"^l1tex__t_sectors_pipe_lsu_mem_global_op_(ld|st)\.sum$",
    r"^lts__t_sectors_srcunit_tex_op_(read|write)\.sum$", r"^lts__t_sectors\.sum$", r"^lts__throughput\.avg\.pct",
    r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$", r"^sm__throughput\.avg\.pct",
    r"^l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed", r"^sm__cycles_elapsed\.avg$", r"^smsp__cycles_active\.avg$",
    r"^sass__inst_executed_shared_(loads|stores)$", r"^smsp__sass_inst_executed_op_shared", r"local_(loads|stores)$",
    r"^sm__sass_inst_executed_op_(global|shared)", r"^smsp__inst_executed_pipe_(fma|alu|lsu|uniform|xu|fmaheavy|fmalite)",
    r"tma", r"^smsp__thread_inst_executed_per_inst_executed\.ratio$",
]
if len(sys.argv) > 2:
    KEYS += sys.argv[2:]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = dict(zip(hdr, r)).get("Kernel Name", "?")
    print("==", name[:100])
    for h, u, v in zip(hdr, units, r):
        if any(re.search(k, h) for k in KEYS) and v not in ("", "0"):
            print(f"  {h} [{u}] = {v}")
