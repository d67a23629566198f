#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (SASS view).
usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_hot.py [N]"""
import csv, sys
n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
exe = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
print("instructions %d, samples %d, warp-instr executed %d" % (len(body), tot, exe))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))
for i in order[:n]:
    r = body[i]
    extra = ""
    for k in ("L1 Wavefronts Shared Excessive", "L2 Theoretical Sectors Global Excessive"):
        if k in ix and r[ix[k]] not in ("", "0"):
            extra += " %s=%s" % (k.split()[1] + k.split()[-1], r[ix[k]])
    print("%5d %5.1f%% #%-4d exec %-8s %s%s" % (int(r[ix["# Samples"]] or 0), 100.0 * int(r[ix["# Samples"]] or 0) / max(tot, 1), i, r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:90], extra))
