"""Aggregates the SASS-level sampling of an .ncu-rep: total samples per stall reason, and the hottest instructions.
usage: python tools/ncu_stalls.py file.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
reasons = [k for k in rows[0].keys() if k.startswith("stall_") and "Not Issued" not in k]
tot = {r: 0 for r in reasons}
total = 0
for r in rows:
    n = int(r["# Samples"] or 0)
    total += n
    for k in reasons:
        tot[k] += int(r[k] or 0)
print("total samples", total, " instructions executed", sum(int(r["Instructions Executed"] or 0) for r in rows))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print("  %-24s %8d  %5.1f%%" % (k, v, 100.0 * v / max(1, total)))
print("hottest instructions:")
for i, r in sorted(enumerate(rows), key=lambda ir: -int(ir[1]["# Samples"] or 0))[:top]:
    top_reason = max(reasons, key=lambda k: int(r[k] or 0))
    print("  #%-5d %6s smp %5.1f%%  exec %9s  %-22s %s" % (i, r["# Samples"], 100.0 * int(r["# Samples"]) / total, r["Instructions Executed"],
                                                   top_reason, r["Source"].strip()[:90]))
