#!/usr/bin/env python
"""ncu --csv log of (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum) per conv launch of ONE step
-> profiles/conv_traffic.json, which bench.py reports as roofline.traffic (mean DRAM bytes per launch)."""
import csv
import json
import sys


def main(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[1:]:
        key = r[col["ID"]]
        d = per.setdefault(key, {"kernel": r[col["Kernel Name"]]})
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        name = r[col["Metric Name"]]
        if name.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        d[name] = v
    launches = list(per.values())
    tot = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in launches)
    out = {"launches": len(launches), "dram_bytes_per_step": tot, "bytes_per_launch": tot / max(1, len(launches)),
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_conv_tc over one 16-frame step "
                     "(tools/gpu_final.sh), cold-cache replay"}
    json.dump(out, open(dst, "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
