#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum ... --csv --log-file x.csv`) -> text table.
usage: tools/launch_summary.py launches.csv "command line that was profiled" > summary.txt"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = {k: i for i, k in enumerate(rows[h])}
agg = collections.OrderedDict()
n = 0
for r in rows[h + 1:]:
    if len(r) < len(H) or r[H["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[H["Kernel Name"]]).replace("srt::", "srt::")[:60]
    v = float(r[H["Metric Value"]].replace(",", ""))
    unit = r[H["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ms; n += 1
tot = sum(a[1] for a in agg.values())
print("launch list of `%s` (ncu --metrics gpu__time_duration.sum --clock-control none, %d launches, %.2f ms in kernels)" % (sys.argv[2] if len(sys.argv) > 2 else "?", n, tot))
print("%-62s %5s %12s %7s" % ("kernel", "n", "total ms", "share"))
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s %5d %12.3f %6.2f%%" % (k, c, ms, 100 * ms / tot))
