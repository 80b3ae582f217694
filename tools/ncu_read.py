"""Prints the key metrics of an .ncu-rep (raw page) -- used to write the profiles/ summaries.
usage: python tools/ncu_read.py file.ncu-rep [more regex]"""
import csv
import io
import re
import subprocess
import sys

KEYS = r"gpu__time_duration.sum|dram__bytes_read.sum$|dram__bytes_write.sum$|dram__throughput.avg.pct|gpu__dram_throughput|sm__pipe_tensor.*cycles_active.avg.pct|sm__inst_executed_pipe_xu|sm__pipe_xu|sm__warps_active.avg.pct|launch__registers_per_thread|launch__occupancy_limit|sm__throughput.avg.pct|smsp__issue_active.avg.pct|sm__inst_executed_pipe_(fma|alu|xu|lsu|uniform|tc|tmem).*pct|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|l1tex__throughput.avg.pct|lts__throughput.avg.pct|smsp__average_warp.*stall|smsp__warp_issue_stalled.*_per_warp_active|launch__grid_size|launch__block_size|sm__cycles_active.avg$|smsp__inst_executed.sum$|sm__pipe_fma_cycles_active|sm__pipe_alu_cycles_active|sm__pipe_fmaheavy|sm__inst_executed_pipe_tensor|dram__cycles_active.avg.pct"
path = sys.argv[1]
extra = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2:]
for v in vals:
    print("== kernel:", v[hdr.index("Kernel Name")][:100])
    for h, u, x in zip(hdr, units, v):
        if re.search(r"\.(max|min|sum)\.pct|realtime|Triage|pct_of_peak_sustained_elapsed", h) and not (extra and re.search(extra, h)):
            continue
        if re.search(KEYS, h) or (extra and re.search(extra, h)):
            print("  %-80s %-14s %s" % (h, u, x))
