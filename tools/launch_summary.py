"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): count, total and share per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[start + 2:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "")[:70]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print("%d launches, %.2f ms summed kernel time (serialised, cold cache)" % (sum(a[0] for a in agg.values()), tot / 1e6))
for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 45]:
    print("%6d %10.1f us %5.1f%%  avg %7.1f us  %s" % (a[0], a[1] / 1e3, a[1] / tot * 100, a[1] / 1e3 / a[0], n))
