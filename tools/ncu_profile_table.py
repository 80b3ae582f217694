"""Builds profiles/<round>_ncu_summary.md and <round>_ncu_traffic.json from gpurun_out/prof2_*.ncu-rep (ncu --set full)."""
import csv
import io
import json
import re
import subprocess
import sys

reports = sys.argv[2:]
tag = sys.argv[1]
WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "hmma_pct",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active": "tc_inst_pct",
    "smsp__issue_active.avg.pct": "issue_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__cycles_active.avg": "sm_cycles",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_pct",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}
rows_out, traffic = [], {}
for path in reports:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, v = rows[0], rows[1], rows[2]
    d = {}
    for h, u, x in zip(hdr, units, v):
        if h in WANT and x != "":
            val = float(x.replace(",", ""))
            d[WANT[h]] = val * UNIT.get(u, 1)
    det = subprocess.run(["ncu", "-i", path, "--page", "details"], capture_output=True, text=True).stdout
    m = re.search(r"highest-utilized pipeline \(([\d.]+)%\)", det)
    tcp = re.search(r"(\w+) is the highest-utilized pipeline", det)
    kname = re.sub(r"\(.*", "", v[hdr.index("Kernel Name")])
    d["kernel"] = kname
    d["top_pipe"] = "%s %s%%" % (tcp.group(1), m.group(1)) if m and tcp else "-"
    f = re.search(r"SM Frequency\s+Ghz\s+([\d.]+)", det)
    d["sm_ghz"] = float(f.group(1)) if f else None
    rows_out.append((path, d))
    traffic[kname] = {"dram_bytes": d.get("dram_read", 0) + d.get("dram_write", 0), "duration_s": d.get("duration"), "report": path}
md = ["| report | kernel | grid x block | regs | duration | SM GHz | DRAM read | DRAM write | DRAM %% of peak | busiest pipe | issue slots busy |",
      "|---|---|---|---|---|---|---|---|---|---|---|"]
for path, d in rows_out:
    md.append("| %s | `%s` | %d x %d | %d | %.1f us | %s | %.1f MB | %.1f MB | %.1f | %s | %.1f %% |" % (
        path.split("/")[-1], d["kernel"][:60], d.get("grid", 0), d.get("block", 0), d.get("regs", 0), d["duration"] * 1e6, d["sm_ghz"],
        d.get("dram_read", 0) / 1e6, d.get("dram_write", 0) / 1e6, d.get("dram_pct", 0), d["top_pipe"], d.get("issue_pct", 0)))
open("profiles/%s_ncu_summary.md" % tag, "w").write("\n".join(md) + "\n")
json.dump(traffic, open("profiles/%s_ncu_traffic.json" % tag, "w"), indent=1)
print("\n".join(md))
