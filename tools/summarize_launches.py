#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys
for f in sys.argv[1:]:
    lines = [l for l in open(f) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else (v * 1000 if unit in ("ms", "msecond") else v)
        a = agg.setdefault(name, [0, 0.0, 0.0, 1e18]); a[0] += 1; a[1] += v; a[2] = max(a[2], v); a[3] = min(a[3], v)
    tot = sum(a[1] for a in agg.values())
    print("%s: %d launches, %.1f us total" % (f, sum(a[0] for a in agg.values()), tot))
    print("| kernel | launches | total us | avg us | min | max | share |\n|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.2f | %.2f | %.2f | %.3f |" % (k[:60], a[0], a[1], a[1] / a[0], a[3], a[2], a[1] / tot))
