"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
usage: python tools/ncu_launch_summary.py gpurun_out/launches.csv [first_id last_id] [--by-grid]
--by-grid keeps launches of one kernel with different grid sizes apart (e.g. the decode-step GEMMs of different shapes)."""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"(vc::)?([A-Za-z0-9_]+)(<[^(]*>)?\(", name)
    if m and m.group(1):
        return m.group(2) + (m.group(3) or "")
    if "at::" in name or "at::native" in name:
        m2 = re.search(r"at::native::([A-Za-z0-9_]+)", name)
        return "torch:" + (m2.group(1) if m2 else name[:40])
    return name[:70]


def main():
    by_grid = "--by-grid" in sys.argv
    argv = [a for a in sys.argv if a != "--by-grid"]
    path = argv[1]
    lo = int(argv[2]) if len(argv) > 2 else 0
    hi = int(argv[3]) if len(argv) > 3 else 1 << 60
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        i = int(r["ID"])
        if lo <= i <= hi:
            name = short(r["Kernel Name"])
            if by_grid:
                name += " grid " + r["Grid Size"].replace(" ", "")
            rows.append((i, name, float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for i, n, t, g, b in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    print("launches %d, total %.3f ms (ids %d..%d)" % (len(rows), tot / 1e6, rows[0][0], rows[-1][0]))
    print("%-70s %6s %10s %7s %9s" % ("kernel", "count", "total_ms", "share", "avg_us"))
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %6d %10.3f %6.1f%% %9.1f" % (n[:70], c, t / 1e6, 100 * t / tot, t / c / 1e3))


if __name__ == "__main__":
    main()
