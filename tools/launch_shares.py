#!/usr/bin/env python
"""ncu launch list (gpu__time_duration.sum, --csv) of the profiled steady-state steps -> markdown share table."""
import csv
import re
import sys
from collections import OrderedDict


def main(src, steps=1):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    col = {h: i for i, h in enumerate(rows[0])}
    agg = OrderedDict()
    for r in rows[1:]:
        if r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", r[col["Kernel Name"]])
        name = re.sub(r"<unnamed>::", "", name)
        name = re.sub(r"\(.*\)$", "", name)[:60]
        v = float(r[col["Metric Value"]].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}[r[col["Metric Unit"]]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("kernel | launches/step | us/step | share\n --- | --- | --- | ---")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("`%s` | %g | %.1f | %.1f%%" % (name, n / steps, us / steps, 100 * us / tot))
    print("total | %g | %.1f |" % (sum(a[0] for a in agg.values()) / steps, tot / steps))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
