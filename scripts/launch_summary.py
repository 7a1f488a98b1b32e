"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): time and launches per kernel name.
    python scripts/launch_summary.py gpurun_out/launches.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iu = hdr.index("Metric Unit")
acc = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)
    a = acc[r[ik][:110]]
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in acc.values())
print("%d launches, %.1f us total" % (sum(a[0] for a in acc.values()), tot))
for k, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%6d x %9.2f us avg = %10.1f us (%5.1f %%)  %s" % (a[0], a[1] / a[0], a[1], 100 * a[1] / tot, k))
